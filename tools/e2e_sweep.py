#!/usr/bin/env python
"""Sweep the host-pipeline tunables (worker threads x chunk size) for the e2e round trip. Run under gpurun."""
import ctypes as C, os, sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from __graft_entry__ import load_package
pkg = load_package()
n, bs = 1 << 30, pkg.datagen.BLOCK_m
data = pkg.datagen.textlike(n, seed=2024)
sh = pkg.Shafa(0)
lib = sh.lib
def pinned(nbytes):
    ptr = lib.shafa_cuda_host_alloc(int(nbytes))
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), (int(nbytes),))
pin_in = pinned(n); pin_in[:] = data
pin_rle = pinned(n + (64 << 20)); pin_pay = pinned(n); pin_out = pinned(n + 4096)
for workers in (1, 2, 3, 4):
    for chunk in (32, 64, 128, 256):
        os.environ["SHAFA_PIPE_WORKERS"] = str(workers); os.environ["SHAFA_PIPE_CHUNK_MB"] = str(chunk)
        ts, te, td = [], [], []
        for i in range(4):
            t0 = time.perf_counter()
            c = sh.encode_raw(pin_in, bs, rle_buf=pin_rle, payload_buf=pin_pay)
            t1 = time.perf_counter()
            B = int(c.n_blocks)
            arr = lambda p, k: np.ctypeslib.as_array(p, (k,))
            out, _ = sh.decode_tables(pin_pay[: int(c.payload_total)], arr(c.payload_offsets, B), arr(c.z, B), arr(c.sym_sizes, B),
                                      arr(c.code_len, B * 256), arr(c.code_bits, B * 256), bool(c.rle_mode), n, out=pin_out)
            t2 = time.perf_counter()
            sh.encoded_free(c)
            if i:
                ts.append(t2 - t0); te.append(t1 - t0); td.append(t2 - t1)
        assert np.array_equal(out[:n], data)
        print(f"workers {workers} chunk {chunk:4d} MiB: e2e {n / min(ts) / 1e9:6.2f} GB/s  (encode {min(te) * 1e3:6.1f} ms, decode {min(td) * 1e3:6.1f} ms)", flush=True)
