timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_cli.py -x -q -m gpu > gpurun_out/t23.log 2>&1; tail -15 gpurun_out/t23.log
