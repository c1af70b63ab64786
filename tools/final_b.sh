#!/bin/bash
# round-end evidence, part B (one GPU): full ncu capture with the right kernel filter, sanitizers, CLI file-to-file, the whole GPU suite
O=gpurun_out
bash tools/profile.sh $O full-only > $O/profile.log 2>&1; tail -2 $O/profile.log
for tool in memcheck racecheck synccheck initcheck; do timeout 600 compute-sanitizer --tool $tool python tools/sanitize_small.py > $O/san_$tool.log 2>&1; echo "$tool rc=$?"; tail -3 $O/san_$tool.log; done
timeout 600 python tools/cli_e2e.py > $O/r02c_cli_e2e.json 2> $O/r02c_cli_e2e.err; tail -c 600 $O/r02c_cli_e2e.json
timeout 1500 python -m pytest tests -m gpu -x -q > $O/gputest5.log 2>&1; tail -3 $O/gputest5.log
