#!/bin/bash
# Run on the GPU box (under gpurun): launch list + one full ncu capture of the hot kernels.
# Outputs land in gpurun_out/; tools/summarize_ncu.py turns them into profiles/<tag>_* here.
set -u
OUT=${1:-gpurun_out}
mkdir -p "$OUT"
if [ "${2:-all}" != "full-only" ]; then
# every launch with its device time (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches.csv" \
    python bench.py --steps 2 --warmup 3 --no-cpu > "$OUT/bench_under_ncu.log" 2>&1
fi
# hot kernels, full set, one step after warm-up (6 matching launches per step; three warm-up steps come first)
ncu --set full --clock-control none --import-source on -k regex:'^k_(rle_encode|sf_encode|dec_sync|dec_write|rle_decode)$' \
    --launch-skip 18 --launch-count 6 -o "$OUT/prof" -f python bench.py --steps 1 --warmup 3 --no-cpu > "$OUT/bench_under_ncu_full.log" 2>&1
ls -la "$OUT"
