timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/gputest3.log 2>&1; tail -10 gpurun_out/gputest3.log
bash tools/profile.sh gpurun_out > gpurun_out/profile.log 2>&1; tail -3 gpurun_out/profile.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; tail -2 gpurun_out/r02_bench.err
for tool in memcheck racecheck synccheck; do timeout 500 compute-sanitizer --tool $tool python tools/sanitize_small.py > gpurun_out/san_$tool.log 2>&1; echo "$tool rc=$?"; tail -4 gpurun_out/san_$tool.log; done
