timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu --no-files > gpurun_out/r02c_bench_2gpu.json 2> gpurun_out/r02c_bench_2gpu.err; tail -3 gpurun_out/r02c_bench_2gpu.err; python - <<P
import json
j=json.loads(open("gpurun_out/r02c_bench_2gpu.json").read().strip().splitlines()[-1])
print(j["n_gpus"], j["value"], "e2e", j["e2e"]["value"], "one_file", json.dumps(j.get("one_file"))[:600])
P
timeout 600 python tools/multi_gpu_e2e.py C3 > gpurun_out/r02c_one_file_2gpu_C3.json 2> gpurun_out/r02c_one_file_2gpu_C3.err; tail -c 700 gpurun_out/r02c_one_file_2gpu_C3.json
