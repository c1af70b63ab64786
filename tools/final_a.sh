#!/bin/bash
# round-end evidence, part A (one GPU): headline bench with the reference arm, the other configs, launch list + full ncu capture
O=gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r02c_bench.json 2> $O/r02c_bench.err; tail -2 $O/r02c_bench.err
for w in C1 C3 C4 C5; do timeout 900 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu > $O/r02c_bench_$w.json 2> $O/r02c_bench_$w.err; tail -2 $O/r02c_bench_$w.err; done
bash tools/profile.sh $O > $O/profile.log 2>&1; tail -2 $O/profile.log
python - <<P
import json
for w in ["", "_C1", "_C3", "_C4", "_C5"]:
    try:
        j = json.load(open("gpurun_out/r02c_bench%s.json" % w))
        print(w or "C2", j["value"], j.get("encode_gbs"), j.get("decode_gbs"), "e2e", j["e2e"]["value"], j["roofline"]["kernel"], j["roofline"]["frac"], j.get("parity", {}).get("decoded", "")[:30])
    except Exception as e:
        print(w, "failed", e)
P
