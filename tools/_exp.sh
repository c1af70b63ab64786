timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_configs.py -x -q -m gpu > gpurun_out/t11.log 2>&1; tail -3 gpurun_out/t11.log
bash tools/ab_bench.sh
timeout 600 python bench.py --workload C4 --steps 3 --warmup 3 --no-cpu --no-files > gpurun_out/b11_c4.json 2> gpurun_out/b11_c4.err; python - <<P
import json
j=json.load(open("gpurun_out/b11_c4.json"))
print(j["value"], j["encode_gbs"], j["decode_gbs"], {k:(v.get("ms_per_step")) for k,v in j["kernels"].items() if v.get("ms_per_step",0)>0.05})
P
