#!/usr/bin/env python3
"""Golden fixture for code tables with codes LONGER THAN 64 BITS (tests/golden/wide_codes/).

Module T cannot produce such a table, but a hand-made .cod may carry it: the reference's packer keeps 32 bytes per code
(/root/reference/src/modules/c.c:18,115-160) and its decoder builds its tree from the strings (d.c:422-504).  This script
writes a small input and a .cod whose codes are '1' * d + '0' for depths d up to 254 (plus '1' * 255, which completes the
tree), runs the UNMODIFIED reference (oracle/_ref/shafa) on them -- `-m c`, then `-m d` on its own output -- and keeps
input, .cod and the reference's .shaf byte for byte.  Run here (the reference binary does not exist on the GPU box).
"""
import shutil
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import oracle_py as orc  # noqa: E402

OUT = Path(__file__).resolve().parent / "wide_codes"
DEPTHS = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 20, 40, 63, 64, 65, 66, 100, 127, 128, 129, 190, 200, 230, 254]


def make(n: int, blocks: int, seed: int):
    rng = np.random.default_rng(seed)
    syms = rng.permutation(256)[: len(DEPTHS) + 1]
    codes = [""] * 256
    for k, d in enumerate(DEPTHS):
        codes[syms[k]] = "1" * d + "0"
    codes[syms[-1]] = "1" * 255
    p = np.array([2.0 ** -min(d + 1, 12) for d in DEPTHS] + [2.0 ** -12])
    p /= p.sum()
    data = syms[rng.choice(len(p), size=n, p=p)].astype(np.uint8)
    cuts = np.linspace(0, n, blocks + 1).astype(int)
    for b in range(blocks):  # every symbol, the deepest ones included, occurs in every block
        data[cuts[b]: cuts[b] + len(syms)] = syms
    cod = f"@N@{blocks}".encode()
    for b in range(blocks):
        cod += f"@{cuts[b + 1] - cuts[b]}@".encode() + ";".join(codes).encode()
    cod += b"@0"
    return data, cod, codes, cuts


def main():
    if not orc.ref_available():
        sys.exit("oracle/_ref/shafa is not built (make -C oracle _ref)")
    OUT.mkdir(exist_ok=True)
    data, cod, codes, cuts = make(30000, 3, 5)
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        (td / "x").write_bytes(data.tobytes())
        (td / "x.cod").write_bytes(cod)
        r = orc.ref_run(["x", "-m", "c"], str(td))
        assert r.returncode == 0, r.stderr
        shaf = (td / "x.shaf").read_bytes()
        # the oracle's packer agrees with the reference on these codes
        want = b"@3" + b"".join(b"@%d@" % len(pl) + pl for pl in
                               (orc.sf_encode_block(data[cuts[b]: cuts[b + 1]], codes).tobytes() for b in range(3)))
        assert shaf == want, "oracle and reference disagree on the wide-code payload"
        (td / "x").unlink()
        r = orc.ref_run(["x.shaf", "-m", "d"], str(td))
        assert r.returncode == 0, r.stderr
        assert (td / "x").read_bytes() == data.tobytes(), "the reference does not decode its own wide-code .shaf"
        (OUT / "x").write_bytes(data.tobytes())
        (OUT / "x.cod").write_bytes(cod)
        shutil.copy(td / "x.shaf", OUT / "x.shaf")
    print("wrote", OUT, len(data), "bytes,", len(shaf), "byte .shaf, longest code 255 bits")


if __name__ == "__main__":
    main()
