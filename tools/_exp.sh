SHAFA_FUSED_TRACE=1 timeout 200 python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/b_trace.json 2>gpurun_out/b_trace.err; grep -c trace gpurun_out/b_trace.err; tail -2 gpurun_out/b_trace.err
