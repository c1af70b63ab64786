#!/bin/bash
# What the GPU box looks like (host cores, memory, NUMA, GPU topology): written to gpurun_out/box.txt
OUT=${1:-gpurun_out}; mkdir -p "$OUT"
{ echo "== nproc"; nproc; echo "== free -g"; free -g; echo "== lscpu"; lscpu | head -25; echo "== numa"; ls /sys/devices/system/node/ 2>/dev/null;
  echo "== nvidia-smi -L"; nvidia-smi -L; echo "== topo"; nvidia-smi topo -m 2>&1 | head -30; echo "== df"; df -h /dev/shm /tmp | cat; } > "$OUT/box.txt" 2>&1
