timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu > gpurun_out/t22.log 2>&1; tail -5 gpurun_out/t22.log
