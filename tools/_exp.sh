timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_configs.py -x -q -m gpu > gpurun_out/t21.log 2>&1; tail -3 gpurun_out/t21.log
bash tools/ab_bench.sh
for w in C3; do timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu --no-files > gpurun_out/b21_$w.json 2> gpurun_out/b21_$w.err; tail -3 gpurun_out/b21_$w.err; python - <<P
import json
try:
    j=json.load(open("gpurun_out/b21_$w.json"))
    print("$w", j["value"], j.get("encode_gbs"), j.get("decode_gbs"), {k:(v.get("ms_per_step")) for k,v in j["kernels"].items() if v.get("ms_per_step",0)>0.05})
except Exception as e: print("$w failed", e)
P
done
