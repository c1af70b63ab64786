#!/bin/bash
# run on the GPU box: bench.py kernel times for the default library and every build/libshafa_<variant>.so given
for v in default "$@"; do
  if [ $v = default ]; then unset SHAFA_LIB; else export SHAFA_LIB=$PWD/shafa-cd_b200/build/libshafa_$v.so; fi
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-files > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err || tail -3 gpurun_out/ab_$v.err
  python - <<P
import json
try:
    j=json.load(open("gpurun_out/ab_$v.json"))
    print("$v", j["value"], {k:v.get("ms_per_step") for k,v in j["kernels"].items() if v.get("ms_per_step",0)>0.03})
except Exception as e: print("$v failed", e)
P
done
