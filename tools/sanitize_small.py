#!/usr/bin/env python3
"""Small inputs through every kernel, meant to run under compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool racecheck python tools/sanitize_small.py

Covers what the kernel tests cover at sizes a sanitizer finishes in minutes: TMA-staged histogram, RLE encode/decode,
SF build/encode, the exact decoder incl. its all-starts path and the lean kernel (uniform bytes: codes of 7-9 bits), the bulk-store copy-out
of k_dec_write, and the fused single-pass decoder (SHAFA_DECODE_FUSED=1, second half).
"""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from __graft_entry__ import load_package  # noqa: E402


def main():
    pkg = load_package()
    # the 700 kB uniform case has one block above DEC_LEAN_MIN_Z (k_dec_lut16 + k_dec_sync_lean) and one below it (the all-starts
    # path inside k_dec_sync)
    cases = [("textlike", 300_000, 65536), ("uniform", 150_000, 65536), ("uniform", 700_000, 655360), ("runheavy", 400_000, 65536),
             ("zipf", 200_000, 65536)]
    with pkg.Shafa(device=0) as sh:
        for fused in ("0", "1"):
            os.environ["SHAFA_DECODE_FUSED"] = fused
            for gen, n, bs in cases:
                data = pkg.datagen.GENERATORS[gen](n, seed=3)
                enc = sh.encode(data, bs, force_freq=True)
                out = sh.decode(enc)
                assert np.array_equal(out, data), (gen, fused)
                if enc.rle_mode:
                    assert np.array_equal(sh.decode(enc, rle_expand=False), enc.rle), (gen, fused)
                print("ok", gen, "fused" if fused == "1" else "exact", "R" if enc.rle_mode else "N", flush=True)
        # code tables with codes of 65..255 bits (sf_wide.cu): the reference's own fixture
        g = ROOT / "tests" / "golden" / "wide_codes"
        data = np.frombuffer((g / "x").read_bytes(), dtype=np.uint8)
        _, sizes, ln, c4 = pkg.parse_cod_wide((g / "x.cod").read_bytes())
        offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
        payload, poff, z = sh.sf_encode_wide_host(data, offs, ln, c4)
        assert pkg.shaf_bytes(z, payload, poff) == (g / "x.shaf").read_bytes()
        assert np.array_equal(sh.sf_decode_wide_host(payload, poff, z, sizes, ln, c4), data)
        print("ok wide codes", flush=True)
    print("SANITIZE_SMALL_OK")


if __name__ == "__main__":
    main()
