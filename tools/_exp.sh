# scratch: the command line of the last experiment sent to the GPU box with `gpurun -- 'bash tools/_exp.sh'`
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_configs.py -x -q -m gpu | tail -2
bash tools/ab_bench.sh
