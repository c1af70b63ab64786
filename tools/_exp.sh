timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_configs.py -x -q -m gpu > gpurun_out/t20.log 2>&1; tail -3 gpurun_out/t20.log
for v in default ns1 ns2; do
  if [ $v = default ]; then unset SHAFA_LIB; else export SHAFA_LIB=$PWD/shafa-cd_b200/build/libshafa_$v.so; fi
  timeout 600 python bench.py --workload C4 --steps 3 --warmup 3 --no-cpu --no-files > gpurun_out/b20_$v.json 2> gpurun_out/b20_$v.err; tail -3 gpurun_out/b20_$v.err; python - <<P
import json
try:
    j=json.load(open("gpurun_out/b20_$v.json"))
    print("$v", j["value"], j.get("encode_gbs"), j.get("decode_gbs"), {k:(v.get("ms_per_step")) for k,v in j["kernels"].items() if v.get("ms_per_step",0)>0.05})
except Exception as e: print("$v failed", e)
P
done
