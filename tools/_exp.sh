nvidia-smi topo -m | head -8
SHAFA_VERBOSE=1 timeout 600 python tools/multi_gpu_e2e.py C3 > gpurun_out/multi_c3_2.json 2> gpurun_out/multi_c3_2.err; cat gpurun_out/multi_c3_2.json; grep "pipelined encode" gpurun_out/multi_c3_2.err | tail -3
SHAFA_PIPE_CHUNK_MB=256 timeout 600 python tools/multi_gpu_e2e.py C3 > gpurun_out/multi_c3_2b.json 2> gpurun_out/multi_c3_2b.err; cat gpurun_out/multi_c3_2b.json
timeout 600 python tools/multi_gpu_e2e.py C2 > gpurun_out/multi_c2_2.json 2> gpurun_out/multi_c2_2.err; cat gpurun_out/multi_c2_2.json
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu > gpurun_out/t8.log 2>&1; tail -3 gpurun_out/t8.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-files > gpurun_out/b8.json 2> gpurun_out/b8.err; tail -2 gpurun_out/b8.err
python - <<P
import json
j=json.load(open("gpurun_out/b8.json"))
print(j["value"], j["encode_gbs"], j["decode_gbs"], "e2e", j["e2e"]["value"], {k:(v.get("ms_per_step"),v.get("frac_of_hbm")) for k,v in j["kernels"].items() if v.get("ms_per_step",0)>0.03})
P
