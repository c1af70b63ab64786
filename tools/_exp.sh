timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_configs.py -x -q -m gpu > gpurun_out/t4.log 2>&1; tail -4 gpurun_out/t4.log
for w in C2 C1 C3 C4 C5; do
  timeout 900 python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/r02_bench_$w.json 2> gpurun_out/r02_bench_$w.err; tail -2 gpurun_out/r02_bench_$w.err
  python - <<P
import json
try:
    j=json.load(open("gpurun_out/r02_bench_$w.json"))
    print("$w", j["value"], "e2e", j["e2e"]["value"], {k:(v.get("ms_per_step"),v.get("frac_of_hbm")) for k,v in j["kernels"].items() if v.get("ms_per_step",0)>0.05}, j.get("parity"), j.get("e2e_files"), j.get("direction_hbm_frac"), j.get("decode_hbm_frac"))
except Exception as e: print("$w failed", e)
P
done
