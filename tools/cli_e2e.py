#!/usr/bin/env python3
"""File-to-file wall times of the two `shafa` binaries on the same input (tmpfs), module by module.

    python tools/cli_e2e.py [BYTES]        # default 1 GiB of the C2 text-like generator, -b m

Ours: shafa-cd_b200/shafa (CUDA path; every invocation pays process start + CUDA context creation + file I/O).
Reference: oracle/_ref/shafa (the unmodified reference build), when present.  Outputs are compared byte for byte.
Prints one JSON line.  Not a bench.py figure: it complements `e2e` (host buffers) with the CLI's file I/O.
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def sha(p: Path) -> str:
    h = hashlib.sha256()
    with open(p, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 24), b""):
            h.update(chunk)
    return h.hexdigest()


def run_pipeline(binary: Path, workdir: Path, data: bytes) -> dict:
    x = workdir / "x"
    x.write_bytes(data)
    t = {}

    def run(args, key):
        t0 = time.perf_counter()
        r = subprocess.run([str(binary)] + args, cwd=workdir, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        t[key] = round(time.perf_counter() - t0, 4)
        if r.returncode != 0:
            raise RuntimeError(f"{binary} {args}: {r.stderr.decode(errors='replace')[:300]}")

    run(["x", "-m", "f", "-b", "m"], "F")
    base = "x.rle" if (workdir / "x.rle").exists() else "x"
    run([base + ".freq", "-m", "t"], "T")
    run([base, "-m", "c"], "C")
    digests = {n: sha(workdir / n) for n in sorted(os.listdir(workdir)) if n != "x"}
    x.unlink()
    run([base + ".shaf", "-m", "d"], "D")
    ok = x.read_bytes() == data
    for p in workdir.iterdir():
        p.unlink()
    return {"seconds": t, "round_trip_ok": ok, "digests": digests}


def run_combined(binary: Path, workdir: Path, data: bytes) -> dict:
    """The default invocation: all compression modules in ONE process (shafa.c:293-298), then -m d."""
    x = workdir / "x"
    x.write_bytes(data)
    t = {}
    t0 = time.perf_counter()
    r = subprocess.run([str(binary), "x", "-b", "m"], cwd=workdir, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    t["FTC"] = round(time.perf_counter() - t0, 4)
    if r.returncode != 0:
        raise RuntimeError(f"{binary} x -b m: {r.stderr.decode(errors='replace')[:300]}")
    base = "x.rle" if (workdir / "x.rle").exists() else "x"
    x.unlink()
    t0 = time.perf_counter()
    r = subprocess.run([str(binary), base + ".shaf", "-m", "d"], cwd=workdir, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    t["D"] = round(time.perf_counter() - t0, 4)
    ok = r.returncode == 0 and x.read_bytes() == data
    for p in workdir.iterdir():
        p.unlink()
    return {"seconds": t, "round_trip_ok": ok}


def main():
    from __graft_entry__ import load_package
    pkg = load_package()
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 30
    data = pkg.datagen.textlike(n, seed=2024).tobytes()
    tmp_root = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    out = {"bytes": n, "workload": "C2 text-like, -b m", "files_on": "tmpfs" if tmp_root else "tmp"}
    ours = ROOT / "shafa-cd_b200" / "shafa"
    ref = ROOT / "oracle" / "_ref" / "shafa"
    with tempfile.TemporaryDirectory(dir=tmp_root) as td:
        run_pipeline(ours, Path(td), data[: 1 << 24])  # warm the page cache / driver once
        o = run_pipeline(ours, Path(td), data)
        tot = sum(o["seconds"].values())
        out["ours"] = {"seconds": o["seconds"], "round_trip_ok": o["round_trip_ok"], "gbs": round(n / tot / 1e9, 3)}
        out["ours"]["one_process_encode"] = run_combined(ours, Path(td), data)
        if ref.exists():
            r = run_pipeline(ref, Path(td), data)
            tot = sum(r["seconds"].values())
            out["reference"] = {"seconds": r["seconds"], "round_trip_ok": r["round_trip_ok"], "gbs": round(n / tot / 1e9, 3),
                                "cores": os.cpu_count()}
            out["reference"]["one_process_encode"] = run_combined(ref, Path(td), data)
            out["files_identical"] = (o["digests"] == r["digests"])
            out["speedup"] = round(out["ours"]["gbs"] / out["reference"]["gbs"], 2)
            oo, rr = out["ours"]["one_process_encode"]["seconds"], out["reference"]["one_process_encode"]["seconds"]
            out["speedup_one_process_encode"] = round(rr["FTC"] / oo["FTC"], 2)     # `shafa x -b m` (F, T, C fused in one pass)
            out["speedup_decode"] = round(rr["D"] / oo["D"], 2)                      # `shafa x.rle.shaf -m d`
        # where the fused invocation spends its time (stderr of SHAFA_VERBOSE=1)
        (Path(td) / "x").write_bytes(data)
        r = subprocess.run([str(ours), "x", "-b", "m"], cwd=td, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env={**os.environ, "SHAFA_VERBOSE": "1"})
        out["ours"]["fused_breakdown"] = [l for l in r.stderr.decode(errors="replace").splitlines() if l.startswith("fused")]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
