#!/usr/bin/env python3
"""ONE file over 1, 2, 4, ... GPUs from ONE host process (shafa_cuda_*_multi): strong scaling of the host-buffer path.

    python tools/multi_gpu_e2e.py [C3|C2] [BYTES]

For every power-of-two device count up to what is visible: encode (F, T, C) and decode (-m d) of the same pinned host
buffers, GB/s of input, artefacts checked against the reference's fingerprints when the size is the config's own.
Prints one JSON line (committed under profiles/).  Not a bench.py figure: bench.py reports the same path as `one_file`.
"""
import ctypes as C
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from __graft_entry__ import load_package  # noqa: E402


def main():
    pkg = load_package()
    lib = pkg.load_library()
    name = sys.argv[1] if len(sys.argv) > 1 else "C3"
    gen, bname, n, fflags, _, desc = bench.WORKLOADS[name]
    if len(sys.argv) > 2:
        n = int(sys.argv[2])
    bs = {"K": pkg.datagen.BLOCK_K, "m": pkg.datagen.BLOCK_m, "M": pkg.datagen.BLOCK_M}[bname]
    case = bench.golden_case(name, n)
    data = pkg.datagen.GENERATORS[gen](n, seed=bench.SEEDS[gen])

    def pinned(nbytes):
        ptr = lib.shafa_cuda_host_alloc(int(nbytes))
        if not ptr:
            raise MemoryError("pinned allocation failed")
        return ptr, np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), (int(nbytes),))

    _, a_in = pinned(n); a_in[:] = data
    _, a_rle = pinned(n + n // 16 + 64 * 2**20)
    _, a_pay = pinned(n + n // 8 + 64 * 2**20)
    _, a_out = pinned(n + 4096)
    total = lib.shafa_cuda_device_count()
    out = {"workload": desc.format(gib=f"{n / 2**30:g} GiB"), "bytes": n, "blocks": (n + bs - 1) // bs, "devices_visible": total, "by_device_count": {}}
    g = 1
    while g <= total:
        with pkg.ShafaMulti(list(range(g))) as m:
            enc_t, dec_t = [], []
            for it in range(4):
                t0 = time.perf_counter()
                c = m.encode_raw(a_in, bs, force_freq="f" in fflags, rle_buf=a_rle, payload_buf=a_pay)
                t1 = time.perf_counter()
                B = int(c.n_blocks)
                z = np.ctypeslib.as_array(c.z, (B,)); po = np.ctypeslib.as_array(c.payload_offsets, (B,)); ss = np.ctypeslib.as_array(c.sym_sizes, (B,))
                ln = np.ctypeslib.as_array(c.code_len, (B * 256,)); cd = np.ctypeslib.as_array(c.code_bits, (B * 256,))
                dec, _ = m.decode_tables(a_pay[: int(c.payload_total)], po, z, ss, ln, cd, bool(c.rle_mode), n, out=a_out)
                t2 = time.perf_counter()
                if it == 0:
                    assert np.array_equal(dec[:n], data), "round trip mismatch"
                    if case:
                        parts = [f"@{B}".encode()]
                        for b in range(B):
                            parts += [f"@{int(z[b])}@".encode(), a_pay[int(po[b]):int(po[b]) + int(z[b])]]
                        pre = ".rle" if c.rle_mode else ""
                        assert bench.sha_arrays(*parts) == case["artefacts"][pre + ".shaf"], ".shaf differs from the reference"
                        if c.rle_mode:
                            assert bench.sha_arrays(a_rle[: int(c.rle_total)]) == case["artefacts"][".rle"], ".rle differs from the reference"
                else:
                    enc_t.append(t1 - t0); dec_t.append(t2 - t1)
                m.encoded_free(c)
        e, d = min(enc_t), min(dec_t)
        out["by_device_count"][str(g)] = {"encode_gbs": round(n / e / 1e9, 2), "decode_gbs": round(n / d / 1e9, 2),
                                         "round_trip_gbs": round(n / (e + d) / 1e9, 2), "parity": "identical to the reference" if case else "round trip only"}
        g *= 2
    print(json.dumps(out))


if __name__ == "__main__":
    main()
