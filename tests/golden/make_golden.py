#!/usr/bin/env python
"""Generate the golden fixtures in this directory by running the UNMODIFIED
reference binary (oracle/_ref/shafa, built from /root/reference/src by
`make -C oracle _ref`).  Run from the repo root inside the build container:

    python tests/golden/make_golden.py

Small cases keep every artefact byte for byte (tests/golden/<case>/...); larger
cases keep sha256 + sizes in big_cases.json and are regenerated from the seeded
generators in shafa-cd_b200/datagen.py at test time.  The reference has no
golden vectors of its own, so these are the pin.
"""
from __future__ import annotations

import hashlib
import json
import shutil
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from __graft_entry__ import load_package  # noqa: E402

datagen = load_package().datagen
REF = ROOT / "oracle" / "_ref" / "shafa"
OUT = Path(__file__).resolve().parent


def run(args, cwd):
    r = subprocess.run([str(REF)] + args, cwd=cwd, capture_output=True)
    if r.returncode != 0:
        raise RuntimeError(f"reference failed: {args}: {r.stderr.decode(errors='replace')}")
    return r


def edges_input() -> np.ndarray:
    """Hand-made block exercising every RLE rule (SURVEY.md section 4 list)."""
    parts = []
    sym = 65
    for run_len in (1, 2, 3, 4, 5, 254, 255, 256, 257, 258, 259, 509, 510, 511, 512, 765, 766, 1020, 1023):
        parts.append(np.full(run_len, sym, dtype=np.uint8))
        sym = 65 + (sym - 64) % 26
    for run_len in (1, 2, 3, 4, 255, 256, 258, 259, 600):
        parts.append(np.full(run_len, 0, dtype=np.uint8))
        parts.append(np.array([7], dtype=np.uint8))
    parts.append(np.array([0, 1, 0, 0, 2, 0, 0, 0, 3, 3, 3, 0, 4, 4, 4, 4, 0], dtype=np.uint8))
    parts.append(np.frombuffer(b"abcabcabcaabbccaaabbbcccaaaabbbbcccc", dtype=np.uint8))
    return np.concatenate(parts)


def block_edge_input() -> np.ndarray:
    """Runs that straddle the 64 KiB default block edge and a short last block."""
    rng = np.random.default_rng(11)
    a = rng.integers(1, 5, size=3 * 65536 + 777, dtype=np.uint8)
    a[65536 - 100:65536 + 300] = 9          # run cut by block edge 1
    a[2 * 65536 - 2:2 * 65536 + 2] = 0      # zeros cut by block edge 2
    a[3 * 65536 - 700:3 * 65536 + 5] = 200  # long run into the short last block
    a[-3:] = 0                               # file ends in zeros
    return a


SMALL = [
    # name, data, F flags
    ("zipf_200k", lambda: datagen.zipf(200_000, seed=1), []),
    ("text_150k", lambda: datagen.textlike(150_000, seed=2), []),
    ("text_150k_cf", lambda: datagen.textlike(150_000, seed=2), ["-c", "f"]),
    ("runs_300k", lambda: datagen.runheavy(300_000, seed=3), []),
    ("edges", edges_input, []),
    ("block_edges", block_edge_input, []),
    ("block_edges_cr", block_edge_input, ["-c", "r"]),
    ("uniform_70k_cf", lambda: datagen.uniform(70_000, seed=4), ["-c", "f"]),
    ("uniform_70k_cr", lambda: datagen.uniform(70_000, seed=4), ["-c", "r"]),
    ("twosym", lambda: np.tile(np.array([120, 121], dtype=np.uint8), 1000), []),
    ("min_1k", lambda: datagen.zipf(1024, seed=5), []),
]

BIG = [
    # name, generator, n, seed, -b flag, F flags
    ("zipf_4m_K", "zipf", 4 * 1024 * 1024, 12345, "K", []),
    ("text_20m_m", "textlike", 20 * 1024 * 1024, 2024, "m", []),
    ("runs_70m_M", "runheavy", 70 * 1024 * 1024, 99, "M", []),
    ("uniform_2m_K_cf", "uniform", 2 * 1024 * 1024, 7, "K", ["-c", "f"]),
    ("mixed_24m_m", "mixed4m", 24 * 1024 * 1024, 5, "m", []),
]


def one_symbol_tail() -> np.ndarray:
    """size % block == 1 in N mode: the last block is one byte."""
    return datagen.zipf(2 * 65536 + 1, seed=6)


def one_symbol_homogeneous_last_block() -> np.ndarray:
    """N mode, last block = 1000 copies of one byte (block 0 decides N for the whole file)."""
    a = datagen.uniform(2 * 65536 + 1000, seed=8)
    a[2 * 65536:] = 77
    return a


def one_symbol_rle_last_block() -> np.ndarray:
    """R mode (-c r), last block = 3 identical non-zero bytes: three literals of one symbol in the .rle block."""
    a = datagen.textlike(65536 + 3, seed=9)
    a[65536:] = 66
    return a


# Inputs with a block that holds ONE distinct symbol: T gives it the empty code (t.c:185), C writes "@0@" and
# the reference decoder then crashes (d.c:533-536), so these pin F, T and C only.
ENCODE_ONLY = [
    ("tail_1byte", one_symbol_tail, []),
    ("homog_last_block", one_symbol_homogeneous_last_block, []),
    ("rle_last_3same", one_symbol_rle_last_block, ["-c", "r"]),
]


def encode_only(workdir: Path, fflags):
    run(["input.bin", "-m", "f"] + fflags, workdir)
    rle = (workdir / "input.bin.rle").exists()
    base = "input.bin.rle" if rle else "input.bin"
    run([base + ".freq", "-m", "t"], workdir)
    run([base, "-m", "c"], workdir)
    return {p.name[len("input.bin"):]: p.read_bytes() for p in sorted(workdir.iterdir()) if p.name != "input.bin"}


def pipeline(workdir: Path, fflags, bflag):
    """F -> T -> C -> D with the reference; returns {suffix: bytes}."""
    b = ["-b", bflag] if bflag else []
    original = (workdir / "input.bin").read_bytes()
    run(["input.bin", "-m", "f"] + b + fflags, workdir)
    arte = {}
    rle = (workdir / "input.bin.rle").exists()
    base = "input.bin.rle" if rle else "input.bin"
    run([base + ".freq", "-m", "t"], workdir)
    run([base, "-m", "c"], workdir)
    if rle and (workdir / "input.bin.freq").exists():  # -c f in R mode: T over the N-header .freq too
        run(["input.bin.freq", "-m", "t"], workdir)
    for p in sorted(workdir.iterdir()):
        if p.name != "input.bin":
            arte[p.name[len("input.bin"):]] = p.read_bytes()
    # decode checks (the decoder overwrites input.bin / input.bin.rle)
    if rle:
        saved_rle = (workdir / "input.bin.rle").read_bytes()
        (workdir / "input.bin.rle").unlink()
        run(["input.bin.rle.shaf", "-m", "d", "-d", "s"], workdir)
        assert (workdir / "input.bin.rle").read_bytes() == saved_rle, "-d s mismatch"
        (workdir / "input.bin").unlink()
        run(["input.bin.rle", "-m", "d", "-d", "r"], workdir)
        assert (workdir / "input.bin").read_bytes() == original, "-d r mismatch"
        (workdir / "input.bin").unlink()
        run(["input.bin.rle.shaf", "-m", "d"], workdir)
        assert (workdir / "input.bin").read_bytes() == original, "-m d mismatch"
    else:
        (workdir / "input.bin").unlink()
        run(["input.bin.shaf", "-m", "d"], workdir)
        assert (workdir / "input.bin").read_bytes() == original, "-m d mismatch"
    return arte


def t_cases() -> bytes:
    """A hand-assembled N-header .freq holding many synthetic histograms for T."""
    rng = np.random.default_rng(77)
    hists = []

    def add(h):
        h = np.asarray(h, dtype=np.uint64)
        assert h.shape == (256,) and h.sum() > 0
        hists.append(h)

    for _ in range(120):  # random sparse
        h = np.zeros(256, dtype=np.uint64)
        k = int(rng.integers(1, 257))
        idx = rng.choice(256, size=k, replace=False)
        h[idx] = rng.integers(1, 10 ** int(rng.integers(1, 7)), size=k)
        add(h)
    for _ in range(60):  # heavy ties
        h = rng.integers(0, 4, size=256).astype(np.uint64)
        if h.sum() == 0:
            h[3] = 1
        add(h)
    for k in (1, 2, 3, 4, 5, 7, 8, 9, 16, 17, 31, 32, 33, 64, 100, 128, 200, 255, 256):  # equal counts
        h = np.zeros(256, dtype=np.uint64)
        h[:k] = 5
        add(h)
        h = np.zeros(256, dtype=np.uint64)
        h[256 - k:] = 1
        add(h)
    fib = [1, 1]
    while len(fib) < 38:
        fib.append(fib[-1] + fib[-2])
    h = np.zeros(256, dtype=np.uint64)
    h[10:10 + len(fib)] = fib  # deep chain (max code length 37)
    add(h)
    h = np.zeros(256, dtype=np.uint64)
    h[: len(fib)] = fib[::-1]
    add(h)
    for base in (2, 3):
        h = np.zeros(256, dtype=np.uint64)
        h[:18] = [base ** i for i in range(18)]
        add(h)
    for _ in range(60):  # zipf-ish with permutation
        p = 1.0 / np.arange(1, 257) ** float(rng.uniform(0.5, 2.5))
        p /= p.sum()
        h = rng.multinomial(int(rng.integers(1000, 5_000_000)), p).astype(np.uint64)
        add(h[rng.permutation(256)])
    for _ in range(40):  # near-uniform (codes 7-9 bits)
        add(rng.integers(2400, 2700, size=256).astype(np.uint64))
    out = [f"@N@{len(hists)}".encode()]
    for h in hists:
        fields = []
        for s in range(256):
            fields.append("" if s and h[s] == h[s - 1] else str(int(h[s])))
        out.append(f"@{int(h.sum())}@".encode() + ";".join(fields).encode())
    out.append(b"@0")
    return b"".join(out)


def main():
    if not REF.exists():
        subprocess.check_call(["make", "-C", str(ROOT / "oracle"), "_ref"])
    for name, gen, fflags in ([] if "--encode-only" in sys.argv else SMALL):
        d = OUT / name
        if d.exists():
            shutil.rmtree(d)
        d.mkdir(parents=True)
        data = np.ascontiguousarray(gen(), dtype=np.uint8)
        with tempfile.TemporaryDirectory() as td:
            td = Path(td)
            (td / "input.bin").write_bytes(data.tobytes())
            arte = pipeline(td, fflags, None)
        (d / "input.bin").write_bytes(data.tobytes())
        for suffix, blob in arte.items():
            (d / ("input.bin" + suffix)).write_bytes(blob)
        (d / "flags.json").write_text(json.dumps({"fflags": fflags, "block": "default"}))
        print(name, data.size, sorted(arte))

    for name, gen, fflags in ENCODE_ONLY:
        d = OUT / "encode_only" / name
        if d.exists():
            shutil.rmtree(d)
        d.mkdir(parents=True)
        data = np.ascontiguousarray(gen(), dtype=np.uint8)
        with tempfile.TemporaryDirectory() as td:
            td = Path(td)
            (td / "input.bin").write_bytes(data.tobytes())
            arte = encode_only(td, fflags)
        (d / "input.bin").write_bytes(data.tobytes())
        for suffix, blob in arte.items():
            (d / ("input.bin" + suffix)).write_bytes(blob)
        (d / "flags.json").write_text(json.dumps({"fflags": fflags, "block": "default"}))
        print("encode_only", name, data.size, sorted(arte))
    if "--encode-only" in sys.argv:
        return

    big = {}
    for name, gname, n, seed, bflag, fflags in BIG:
        data = datagen.GENERATORS[gname](n, seed=seed)
        with tempfile.TemporaryDirectory(dir="/dev/shm") as td:
            td = Path(td)
            (td / "input.bin").write_bytes(data.tobytes())
            arte = pipeline(td, fflags, bflag)
        big[name] = {
            "generator": gname, "n": n, "seed": seed, "block": bflag, "fflags": fflags,
            "input_sha256": hashlib.sha256(data.tobytes()).hexdigest(),
            "artefacts": {s: {"size": len(b), "sha256": hashlib.sha256(b).hexdigest()} for s, b in arte.items()},
        }
        print(name, {s: len(b) for s, b in arte.items()})
    (OUT / "big_cases.json").write_text(json.dumps(big, indent=1, sort_keys=True))

    freq = t_cases()
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        (td / "tcases.freq").write_bytes(freq)
        run(["tcases.freq", "-m", "t"], td)
        cod = (td / "tcases.cod").read_bytes()
    (OUT / "tcases.freq").write_bytes(freq)
    (OUT / "tcases.cod").write_bytes(cod)
    print("tcases", len(freq), len(cod))


if __name__ == "__main__":
    main()
