#!/usr/bin/env python
"""Full-size golden fingerprints of BASELINE.json's five configurations, produced by the
UNMODIFIED reference binary (oracle/_ref/shafa, built from /root/reference/src by
`make -C oracle _ref`).  Run from the repo root inside the build container:

    python tests/golden/make_full_golden.py [case ...]      # default: every case

Per case the seeded generator of shafa-cd_b200/datagen.py writes the input, the reference
runs F, T and C on it, and size + sha256 of every artefact (.rle .freq .rle.freq .cod
.rle.cod .shaf / .rle.shaf) go to tests/golden/full_cases.json.  The reference decoders are
then run on the reference's own outputs (-d s, -d r and the combined -m d) and their outputs
are checked against the .rle / input hashes, so the decode targets are pinned too.  The `-m
gpu` tests (tests/test_gpu_full_configs.py) and bench.py compare the CUDA path's hashes with
this file; /root/reference is not needed at test time.

Work files live on /dev/shm when they fit and on /tmp otherwise (the reference keeps the whole
file resident).
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import subprocess
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from __graft_entry__ import load_package  # noqa: E402

datagen = load_package().datagen
REF = ROOT / "oracle" / "_ref" / "shafa"
OUT = Path(__file__).resolve().parent / "full_cases.json"
GIB, MIB = 1 << 30, 1 << 20

# name -> generator, bytes, seed, -b flag, F flags, BASELINE.json config it stands for
CASES = {
    "C1_zipf_64m_K": ("zipf", 64 * MIB, 12345, "K", [], "configs[0]"),
    "C2_text_1g_m": ("textlike", 1 * GIB, 2024, "m", [], "configs[1]"),
    "C3_runs_4g_M": ("runheavy", 4 * GIB, 99, "M", [], "configs[2]"),
    "C4_uniform_4g_K_cf": ("uniform", 4 * GIB, 7, "K", ["-c", "f"], "configs[3]"),
    "C5_mixed_4g_m": ("mixed_tiled", 4 * GIB, 5, "m", [], "configs[4] at a quarter of its size (fallback when host memory is short)"),
    "C5_mixed_16g_m": ("mixed_tiled", 16 * GIB, 5, "m", [], "configs[4]"),
}


def sha_file(path: Path) -> dict:
    h = hashlib.sha256()
    n = 0
    with open(path, "rb") as f:
        while True:
            b = f.read(64 * MIB)
            if not b:
                break
            h.update(b)
            n += len(b)
    return {"size": n, "sha256": h.hexdigest()}


def run(args, cwd):
    t0 = time.time()
    r = subprocess.run([str(REF)] + args, cwd=cwd, capture_output=True)
    if r.returncode != 0:
        raise RuntimeError(f"reference failed: {args}: {r.stderr.decode(errors='replace')[:400]}")
    return round(time.time() - t0, 2)


def workdir_for(nbytes: int) -> Path:
    shm_free = shutil.disk_usage("/dev/shm").free if os.path.isdir("/dev/shm") else 0
    base = Path("/dev/shm") if 5 * nbytes < shm_free and nbytes <= 2 * GIB else Path("/tmp")
    d = base / "shafa_full_golden"
    if d.exists():
        shutil.rmtree(d)
    d.mkdir(parents=True)
    return d


def one(name: str) -> dict:
    gname, n, seed, bflag, fflags, what = CASES[name]
    wd = workdir_for(n)
    t0 = time.time()
    data = datagen.GENERATORS[gname](n, seed=seed)
    gen_s = round(time.time() - t0, 1)
    x = wd / "x"
    with open(x, "wb") as f:
        for lo in range(0, n, 256 * MIB):
            f.write(memoryview(data[lo:lo + 256 * MIB]))
    input_fp = {"size": n, "sha256": hashlib.sha256(data).hexdigest()}
    del data
    secs = {}
    secs["F"] = run(["x", "-m", "f", "-b", bflag] + fflags, wd)
    rle = (wd / "x.rle").exists()
    base = "x.rle" if rle else "x"
    secs["T"] = run([base + ".freq", "-m", "t"], wd)
    secs["C"] = run([base, "-m", "c"], wd)
    if rle and (wd / "x.freq").exists():
        run(["x.freq", "-m", "t"], wd)   # -c f in R mode: the N-header .freq has its own .cod
    arte = {p.name[1:]: sha_file(p) for p in sorted(wd.iterdir()) if p.name != "x"}
    entry = {"generator": gname, "n": n, "seed": seed, "block": bflag, "fflags": fflags, "config": what,
             "rle_mode": rle, "input": input_fp, "artefacts": arte}
    # decode targets, pinned by the reference's own decoders
    x.unlink()
    if rle:
        rle_fp = arte[".rle"]
        (wd / "x.rle").unlink()
        secs["D_s"] = run(["x.rle.shaf", "-m", "d", "-d", "s"], wd)
        assert sha_file(wd / "x.rle") == rle_fp, "-d s output differs from .rle"
        secs["D_r"] = run(["x.rle", "-m", "d", "-d", "r"], wd)
        assert sha_file(x) == input_fp, "-d r output differs from the input"
        x.unlink()
        secs["D"] = run(["x.rle.shaf", "-m", "d"], wd)
        assert sha_file(x) == input_fp, "-m d output differs from the input"
    else:
        secs["D"] = run(["x.shaf", "-m", "d"], wd)
        assert sha_file(x) == input_fp, "-m d output differs from the input"
    entry["reference_decoders_checked"] = True
    entry["reference_seconds_here"] = secs
    entry["generator_seconds_here"] = gen_s
    shutil.rmtree(wd)
    return entry


def main():
    if not REF.exists():
        subprocess.check_call(["make", "-C", str(ROOT / "oracle"), "_ref"])
    names = sys.argv[1:] or list(CASES)
    full = json.loads(OUT.read_text()) if OUT.exists() else {}
    for name in names:
        full[name] = one(name)
        print(name, {k: v["size"] for k, v in full[name]["artefacts"].items()}, full[name]["reference_seconds_here"], flush=True)
        OUT.write_text(json.dumps(full, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
