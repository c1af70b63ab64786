timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu --no-files > gpurun_out/r02c_bench_4gpu.json 2> gpurun_out/r02c_bench_4gpu.err; tail -2 gpurun_out/r02c_bench_4gpu.err; python - <<P
import json
j=json.loads(open("gpurun_out/r02c_bench_4gpu.json").read().strip().splitlines()[-1])
print(j["n_gpus"], j["value"], "e2e", j["e2e"]["value"], "one_file", json.dumps(j.get("one_file"))[:700])
P
