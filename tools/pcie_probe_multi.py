#!/usr/bin/env python3
"""Aggregate host<->device copy ceiling of this box with 1, 2, 4, ... GPUs copying AT THE SAME TIME (pinned,
portable host memory; one host thread and two streams per GPU; both directions at once, as the e2e path does).

    python tools/pcie_probe_multi.py [BYTES_PER_GPU]

The e2e figure of bench.py at N GPUs is bounded by this aggregate (all GPUs of these boxes sit on one NUMA node),
not by the per-GPU link.  Prints one JSON line: per GPU count, GB/s up, down and each way with both busy.
"""
import json
import sys
import threading
import time

import torch


def run(ngpu: int, n: int, mode: str, reps: int = 4) -> float:
    """aggregate GB/s over all GPUs for `mode` in {'h2d', 'd2h', 'both'} (both: GB/s per direction)"""
    bufs = []
    for g in range(ngpu):
        dev = torch.device("cuda", g)
        bufs.append((dev, torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory(),
                     torch.empty(n, dtype=torch.uint8, device=dev), torch.empty(n, dtype=torch.uint8, device=dev),
                     torch.cuda.Stream(dev), torch.cuda.Stream(dev)))
    barrier = threading.Barrier(ngpu + 1)

    def worker(b):
        dev, h_a, h_b, d_a, d_b, s1, s2 = b
        torch.cuda.set_device(dev)

        def once():
            if mode in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    d_a.copy_(h_a, non_blocking=True)
            if mode in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    h_b.copy_(d_b, non_blocking=True)
        once()
        s1.synchronize(); s2.synchronize()
        barrier.wait()
        for _ in range(reps):
            once()
        s1.synchronize(); s2.synchronize()
        barrier.wait()

    threads = [threading.Thread(target=worker, args=(b,)) for b in bufs]
    for t in threads:
        t.start()
    barrier.wait()
    t0 = time.perf_counter()
    barrier.wait()
    dt = time.perf_counter() - t0
    for t in threads:
        t.join()
    return ngpu * n * reps / dt / 1e9


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 30
    total = torch.cuda.device_count()
    out = {"bytes_per_gpu": n, "gpus_visible": total, "by_gpu_count": {}}
    g = 1
    while g <= total:
        out["by_gpu_count"][str(g)] = {"h2d_gbs": round(run(g, n, "h2d"), 1), "d2h_gbs": round(run(g, n, "d2h"), 1),
                                       "bidir_each_way_gbs": round(run(g, n, "both"), 1)}
        g *= 2
    print(json.dumps(out))


if __name__ == "__main__":
    main()
