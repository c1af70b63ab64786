timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err; tail -2 gpurun_out/r02d_bench.err
python - <<P
import json
j = json.load(open("gpurun_out/r02d_bench.json"))
print(j["value"], j.get("encode_gbs"), j.get("decode_gbs"), "e2e", j["e2e"]["value"], "files", j.get("e2e_files"), j["roofline"]["kernel"], j["roofline"]["frac"], j["roofline"]["traffic"], j["cpu_baseline"]["value"], j.get("parity"))
P
