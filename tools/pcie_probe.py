#!/usr/bin/env python3
"""Host<->device copy rates of this box (pinned memory), alone and both directions at once.

The e2e figure of bench.py moves N + Z bytes up and R + Z + N bytes down per round trip; this probe
gives the copy-bound ceiling to read that number against.  Prints one JSON line.
"""
import json
import sys

import torch


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 30
    dev = torch.device("cuda", 0)
    h_a = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_b = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_a = torch.empty(n, dtype=torch.uint8, device=dev)
    d_b = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def timed(fn, reps=5):
        fn()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize(dev)
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / reps / 1e3

    def h2d():
        with torch.cuda.stream(s1):
            d_a.copy_(h_a, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            h_b.copy_(d_b, non_blocking=True)

    def both():
        h2d()
        d2h()

    def sync_wrap(f):
        def g():
            f()
            s1.synchronize()
            s2.synchronize()
        return g

    t_up = timed(sync_wrap(h2d))
    t_dn = timed(sync_wrap(d2h))
    t_bi = timed(sync_wrap(both))
    # chunked copies as the pipelined host path issues them (64 MiB pieces)
    ch = 64 << 20

    def h2d_chunks():
        with torch.cuda.stream(s1):
            for o in range(0, n, ch):
                d_a[o:o + ch].copy_(h_a[o:o + ch], non_blocking=True)

    t_upc = timed(sync_wrap(h2d_chunks))
    print(json.dumps({"bytes": n, "h2d_gbs": round(n / t_up / 1e9, 2), "d2h_gbs": round(n / t_dn / 1e9, 2),
                      "bidir_each_gbs": round(n / t_bi / 1e9, 2), "h2d_64MiB_chunks_gbs": round(n / t_upc / 1e9, 2)}))


if __name__ == "__main__":
    main()
