#!/bin/bash
# round-end evidence, part C (one GPU): smoke, C1/C3 bench lines with the final bench.py, sanitizers incl. the wide-code kernels, whole GPU suite
O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for w in C1 C3; do timeout 900 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu > $O/r02c_bench_$w.json 2> $O/r02c_bench_$w.err; tail -2 $O/r02c_bench_$w.err; done
for tool in memcheck racecheck synccheck initcheck; do timeout 600 compute-sanitizer --tool $tool python tools/sanitize_small.py > $O/san_$tool.log 2>&1; echo "$tool rc=$?"; tail -3 $O/san_$tool.log; done
timeout 1500 python -m pytest tests -m gpu -x -q > $O/gputest6.log 2>&1; tail -3 $O/gputest6.log
python - <<P
import json
for w in ["_C1", "_C3"]:
    j = json.load(open("gpurun_out/r02c_bench%s.json" % w))
    print(w, j["value"], j.get("encode_gbs"), j.get("decode_gbs"), "e2e", j["e2e"]["value"], j["roofline"]["kernel"], j["roofline"]["frac"])
P
