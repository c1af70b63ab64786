timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_configs.py -x -q -m gpu > gpurun_out/t14.log 2>&1; tail -3 gpurun_out/t14.log
bash tools/ab_bench.sh
for w in C4 C5; do timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu --no-files > gpurun_out/b14_$w.json 2> gpurun_out/b14_$w.err; python - <<P
import json
try:
    j=json.load(open("gpurun_out/b14_$w.json"))
    print("$w", j["value"], j.get("encode_gbs"), j.get("decode_gbs"), {k:(v.get("ms_per_step")) for k,v in j["kernels"].items() if v.get("ms_per_step",0)>0.05})
except Exception as e: print("$w failed", e)
P
done
bash tools/profile.sh gpurun_out full-only > gpurun_out/profile.log 2>&1; tail -2 gpurun_out/profile.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_(rle_encode|rle_decode)' --launch-skip 6 --launch-count 2 -o gpurun_out/prof_c3 -f python bench.py --workload C3 --steps 1 --warmup 3 --no-cpu --no-files > gpurun_out/ncu_c3.log 2>&1; tail -2 gpurun_out/ncu_c3.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_dec_sync' --launch-skip 3 --launch-count 1 -o gpurun_out/prof_c4 -f python bench.py --workload C4 --steps 1 --warmup 3 --no-cpu --no-files > gpurun_out/ncu_c4.log 2>&1; tail -2 gpurun_out/ncu_c4.log
