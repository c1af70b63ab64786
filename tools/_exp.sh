timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_multi.py tests/test_cli.py -x -q -m gpu > gpurun_out/t6.log 2>&1; tail -3 gpurun_out/t6.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/b6.json 2> gpurun_out/b6.err; tail -2 gpurun_out/b6.err
python - <<P
import json
j=json.load(open("gpurun_out/b6.json"))
print(j["value"], j["encode_gbs"], j["decode_gbs"], "e2e", j["e2e"]["value"], j.get("e2e_files"), {k:(v.get("ms_per_step"),v.get("frac_of_hbm")) for k,v in j["kernels"].items() if v.get("ms_per_step",0)>0.05})
P
timeout 600 python bench.py --workload C4 --steps 3 --warmup 3 --no-cpu --no-files > gpurun_out/b6_c4.json 2> gpurun_out/b6_c4.err; python - <<P
import json
j=json.load(open("gpurun_out/b6_c4.json"))
print(j["value"], j["encode_gbs"], j["decode_gbs"], {k:(v.get("ms_per_step"),v.get("frac_of_hbm")) for k,v in j["kernels"].items() if v.get("ms_per_step",0)>0.05})
P
timeout 900 python tools/cli_e2e.py > gpurun_out/cli_e2e2.json 2> gpurun_out/cli_e2e2.err; cat gpurun_out/cli_e2e2.json; tail -3 gpurun_out/cli_e2e2.err
