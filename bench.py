#!/usr/bin/env python
"""bench.py -- F+C encode & D decode throughput of the Shafa block pipeline on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload C2|C1|C3|C4|C5] [--bytes B] [--no-cpu] [--no-files] [--no-one-file]

A "step" is one full F -> T -> C -> D round trip over one synthetic file of
BASELINE.json configs[1] (C2): 1 GiB low-entropy text-like bytes, m = 8 MiB blocks
(R mode).  For N > 1 (torchrun, one rank per GPU) every rank runs the same step on
its own 1 GiB shard of blocks -- blocks are independent, there is no collective on
the data path (weak scaling); timing is the max over ranks.

One JSON line on stdout (rank 0):
  value         device-resident (input already in HBM), whole job
  e2e           the host-buffer C-ABI calls (shafa_cuda_encode_host / decode_host) on pinned host
                buffers, host<->device copies inside the timed region
  e2e_files     (N = 1) the `shafa` CLI file to file on tmpfs: `shafa x -b m` then `shafa x.rle.shaf -m d`
                -- what the reference arm pays too (process start, CUDA start-up, file I/O)
  one_file      (N > 1) ONE 4 GiB run-heavy file (C3, -b M, 64 blocks) sharded over all N GPUs by ONE
                host process through shafa_cuda_*_multi (strong scaling); the other ranks idle meanwhile
  roofline      the dominant kernel, CUDA events on the launch stream; `traffic` = its DRAM bytes per
                launch from the newest committed ncu capture (profiles/*_traffic.json)
  parity        sha256 of the timed step's .rle / .freq / .cod / .shaf against the reference's fingerprints
                (tests/golden/full_cases.json), rank 0, outside the timed region
  cpu_baseline  the unmodified reference binary (oracle/_ref/shafa) on this box's host cores, full 1 GiB
`--impl reference` times that binary alone on the same 1 GiB config (rank 0 only under torchrun).
`--workload` times the other BASELINE configs at their own sizes with the same harness (C5 is
decode-only: -d s, -d r and the combined -m d); the metric is quoted on C2.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

GIB = 1 << 30
MIB = 1 << 20
METRIC = "F+C encode & D decode round-trip throughput, GB/s of input"
CPU_SAMPLE_BYTES = int(os.environ.get("SHAFA_BENCH_CPU_SAMPLE_MB", "1024")) * MIB   # the reference arm runs the full 1 GiB config

# generator, block size name, default bytes, F flags, golden case (full size only), description
WORKLOADS = {
    "C2": ("textlike", "m", 1 * GIB, (), "C2_text_1g_m", "C2: {gib} text-like synthetic (seed 2024), -b m (8 MiB blocks), F->T->C->D round trip"),
    "C1": ("zipf", "K", 64 * MIB, (), "C1_zipf_64m_K", "C1: {gib} zipf synthetic (seed 12345), -b K (640 KiB blocks), N mode, F->T->C->D round trip"),
    "C3": ("runheavy", "M", 4 * GIB, (), "C3_runs_4g_M", "C3: {gib} run-heavy synthetic (seed 99), -b M (64 MiB blocks), R mode, F->T->C->D round trip"),
    "C4": ("uniform", "K", 4 * GIB, ("f",), "C4_uniform_4g_K_cf", "C4: {gib} uniform synthetic (seed 7), -c f -b K (640 KiB blocks), N mode, F->T->C->D round trip"),
    "C5": ("mixed_tiled", "m", 16 * GIB, (), "C5_mixed_16g_m", "C5: {gib} mixed synthetic (seed 5), -b m, decode only: -d s + -d r (= -m d staged)"),
}
SEEDS = {"textlike": 2024, "zipf": 12345, "runheavy": 99, "uniform": 7, "mixed_tiled": 5}


def workload_config(pkg, name: str, n: int, world: int) -> dict:
    """Input properties only: identical for our arm and the reference arm on the same workload."""
    gen, bname, _, fflags, _, desc = WORKLOADS[name]
    bs = {"K": pkg.datagen.BLOCK_K, "m": pkg.datagen.BLOCK_m, "M": pkg.datagen.BLOCK_M}[bname]
    return {"workload": desc.format(gib=f"{n / GIB:g} GiB"), "bytes_per_gpu": n, "block_size": bs, "blocks_per_gpu": (n + bs - 1) // bs,
            "flags": ["-c " + f for f in fflags], "l2": "inputs larger than the 126 MB L2 (one pass over the whole shard per kernel)",
            "sharding": f"{world} x independent block shards, no collective"}


def golden_case(name: str, n: int):
    _, _, default_n, _, case, _ = WORKLOADS[name]
    p = ROOT / "tests" / "golden" / "full_cases.json"
    if n != default_n or not p.exists():
        return None
    return json.loads(p.read_text()).get(case)


def measured_peak_hbm():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def profiled_traffic(kernel: str):
    """DRAM bytes per launch of `kernel` from the newest committed ncu capture (profiles/*_traffic.json, written by
    tools/summarize_ncu.py from one `ncu --set full` run of this bench on the C2 workload); (None, why) if there is none."""
    files = sorted((ROOT / "profiles").glob("*_traffic.json"))
    if not files:
        return None, "no profiles/*_traffic.json"
    try:
        j = json.loads(files[-1].read_text())
        return int(j["bytes_per_launch"][kernel]), f"profiles/{files[-1].name}"
    except Exception as e:  # noqa: BLE001
        return None, f"profiles/{files[-1].name}: {type(e).__name__}"


def sha_arrays(*chunks) -> dict:
    h = hashlib.sha256()
    n = 0
    for c in chunks:
        if isinstance(c, np.ndarray):
            for lo in range(0, c.size, 256 * MIB):
                h.update(memoryview(c[lo:lo + 256 * MIB]))
            n += c.size
        else:
            h.update(c)
            n += len(c)
    return {"size": n, "sha256": h.hexdigest()}


# ----------------------------------------------------------------------------------
# reference arm / cpu baseline: the unmodified reference binary on host cores
# ----------------------------------------------------------------------------------

def reference_round_trip(sample: np.ndarray, bflag: str, workdir: Path, extra=(), fflags=()) -> dict:
    """F, T, C, D with oracle/_ref/shafa on `sample`; wall seconds per module.  `extra`: more CLI flags for every call."""
    from oracle import oracle_py as orc

    x = workdir / "x"
    with open(x, "wb") as f:
        f.write(memoryview(sample))
    t = {}

    def run(args, key):
        t0 = time.perf_counter()
        r = orc.ref_run(list(args) + list(extra), str(workdir), timeout=3600)
        t[key] = time.perf_counter() - t0
        if r.returncode != 0:
            raise RuntimeError(f"reference failed {args}: {r.stderr.decode(errors='replace')[:300]}")

    run(["x", "-m", "f", "-b", bflag] + [a for f in fflags for a in ("-c", f)], "F")
    rle = (workdir / "x.rle").exists()
    base = "x.rle" if rle else "x"
    run([base + ".freq", "-m", "t"], "T")
    run([base, "-m", "c"], "C")
    sizes = {p.name[1:]: p.stat().st_size for p in workdir.iterdir() if p.name != "x"}
    x.unlink()
    run([base + ".shaf", "-m", "d"], "D")
    ok = x.stat().st_size == sample.size and np.array_equal(np.fromfile(x, dtype=np.uint8), sample)
    for p in workdir.iterdir():
        p.unlink()
    return {"seconds": t, "ok": ok, "rle_mode": rle, "sizes": sizes}


def port_round_trip(sample: np.ndarray, block: int) -> dict:
    """Fallback when the reference binary is absent: the oracle restatement (1 thread)."""
    from oracle import oracle_py as orc

    t = {}
    t0 = time.perf_counter(); f = orc.module_f(sample, block); t["F"] = time.perf_counter() - t0
    sym = f["rle"] if f["rle_mode"] else sample.tobytes()
    freq = f["rle_freq"] if f["rle_mode"] else f["freq"]
    t0 = time.perf_counter(); cod = orc.module_t(freq); t["T"] = time.perf_counter() - t0
    t0 = time.perf_counter(); shaf = orc.module_c(sym, cod); t["C"] = time.perf_counter() - t0
    t0 = time.perf_counter(); out = orc.module_d_shaf(shaf, cod, f["rle_mode"]); t["D"] = time.perf_counter() - t0
    return {"seconds": t, "ok": out == sample.tobytes(), "rle_mode": f["rle_mode"], "sizes": {}}


def cpu_arm(pkg, steps: int, warmup: int, per_core: bool = False, sample=None):
    """Returns (per-step seconds list, description dict, bytes).  per_core: also one --no-multithread pass."""
    from oracle import oracle_py as orc

    if sample is None:
        sample = pkg.datagen.textlike(CPU_SAMPLE_BYTES, seed=2024)
    nbytes = sample.size
    use_ref = orc.ref_available()
    tmp_root = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    times, last = [], None
    with tempfile.TemporaryDirectory(dir=tmp_root) as td:
        for i in range(warmup + steps):
            r = reference_round_trip(sample, "m", Path(td)) if use_ref else port_round_trip(sample, pkg.datagen.BLOCK_m)
            if not r["ok"]:
                raise RuntimeError("CPU arm round trip mismatch")
            if i >= warmup:
                times.append(sum(r["seconds"].values()))
                last = r
        single = reference_round_trip(sample, "m", Path(td), extra=("--no-multithread",)) if (use_ref and per_core) else None
    desc = {
        "kind": "reference" if use_ref else "port",
        "cores": os.cpu_count() if use_ref else 1,
        "sample": f"{nbytes / MIB:g} MiB of the C2 text-like generator (seed 2024), -b m ({nbytes // pkg.datagen.BLOCK_m} blocks)"
                  + (" = the whole config" if nbytes == GIB else "")
                  + f", files on {'tmpfs' if tmp_root else 'tmp'}; F,T single-threaded, C,D one pthread per block (reference design); "
                    "seconds = the four module invocations (process start + file I/O included)",
        "module_seconds": {k: round(v, 4) for k, v in last["seconds"].items()},
        "artefact_sizes": last["sizes"],
    }
    if single:
        desc["single_thread"] = {"value": round(nbytes / sum(single["seconds"].values()) / 1e9, 4), "unit": "GB/s",
                                 "flag": "--no-multithread", "module_seconds": {k: round(v, 4) for k, v in single["seconds"].items()}}
    return times, desc, nbytes


# ----------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------

class ClockSampler:
    """nvidia-smi in the background (the recipe's clocks line plus a timestamp); only the samples taken
    inside the marked load windows count.  nvidia-smi needs a moment to start, so it is launched before
    the warm-up and `wait_ready` blocks until its first line is there."""
    FIELDS = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.path = tempfile.NamedTemporaryFile(prefix="clocks_", suffix=".csv", delete=False).name
        self.proc = None
        self.windows = []
        if shutil.which("nvidia-smi"):
            try:
                self.proc = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                              "-lms", "50"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
            except Exception:
                self.proc = None

    def wait_ready(self, timeout_s: float = 10.0):
        t_end = time.time() + timeout_s
        while self.proc and time.time() < t_end:
            if os.path.getsize(self.path) > 0:
                return
            time.sleep(0.02)

    def window(self, t0: float, t1: float):
        self.windows.append((t0, t1))

    @staticmethod
    def _ts(text: str):
        import datetime
        try:
            return datetime.datetime.strptime(text.strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
        except ValueError:
            return None

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in Path(self.path).read_text().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                rows.append((self._ts(f[0]), float(f[1]), float(f[2]), [n for n, v in zip(names, f[5:9]) if v.lower().startswith("active")]))
            except ValueError:
                continue
        os.unlink(self.path)
        # a sample describes the ~50 ms before its timestamp: accept it up to one period after a window
        inside = [r for r in rows if r[0] is not None and any(t0 <= r[0] <= t1 + 0.06 for t0, t1 in self.windows)]
        scope = "load windows"
        if not inside:
            inside, scope = rows, "whole run (no sample fell inside a load window)"
        sm = [r[1] for r in inside]
        reasons = sorted({n for r in inside for n in r[3]})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max((r[2] for r in rows), default=None),
                "samples": len(sm), "scope": scope, "reasons": reasons}


# ----------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------

ALGO_BYTES_DOC = {
    "rle_encode": "N + R", "rle_encode_sizes": "n0 (block-0 probe)", "histogram": "S", "sf_encode": "S + Z",
    "dec_fused": "Z + S", "dec_sync": "Z", "dec_sync_lean": "Z", "dec_sync (both kernels)": "Z", "dec_write": "Z + S", "rle_decode": "R + N",
}


def cli_files_round_trip(data: np.ndarray, bflag: str, fflags=()) -> dict:
    """`shafa x -b ..` then `shafa x[.rle].shaf -m d` with OUR binary, file to file on tmpfs (what the reference arm pays)."""
    binary = ROOT / "shafa-cd_b200" / "shafa"
    tmp_root = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    with tempfile.TemporaryDirectory(dir=tmp_root) as td:
        x = Path(td) / "x"
        with open(x, "wb") as f:
            f.write(memoryview(data))
        t = {}
        t0 = time.perf_counter()
        r = subprocess.run([str(binary), "x", "-b", bflag] + [a for f in fflags for a in ("-c", f)], cwd=td, capture_output=True)
        t["FTC"] = time.perf_counter() - t0
        if r.returncode != 0:
            raise RuntimeError(f"shafa x -b {bflag}: {r.stderr.decode(errors='replace')[:300]}")
        shaf = "x.rle.shaf" if (Path(td) / "x.rle.shaf").exists() else "x.shaf"
        x.unlink()
        t0 = time.perf_counter()
        r = subprocess.run([str(binary), shaf, "-m", "d"], cwd=td, capture_output=True)
        t["D"] = time.perf_counter() - t0
        ok = r.returncode == 0 and x.stat().st_size == data.size and np.array_equal(np.fromfile(x, dtype=np.uint8), data)
    if not ok:
        raise RuntimeError("CLI file round trip mismatch")
    tot = sum(t.values())
    return {"value": round(data.size / tot / 1e9, 4), "unit": "GB/s", "seconds": {k: round(v, 3) for k, v in t.items()},
            "api": "shafa-cd_b200/shafa: `shafa x -b m` (F,T,C fused) then `shafa x.rle.shaf -m d`, files on tmpfs; process start and CUDA start-up included"}


def one_file_multi(pkg, world: int, lib) -> dict:
    """ONE file over all N GPUs from this one process (shafa_cuda_*_multi): C3, 4 GiB, -b M = 64 blocks, strong scaling.
    Host buffers pinned (portable); copies inside the timed region; artefacts checked against the reference's fingerprints."""
    import ctypes as C

    name = "C3"
    gen, _, n, _, _, desc = WORKLOADS[name]
    n = int(os.environ.get("SHAFA_BENCH_ONE_FILE_BYTES", n))
    bs = pkg.datagen.BLOCK_M
    case = golden_case(name, n)
    data = pkg.datagen.GENERATORS[gen](n, seed=SEEDS[gen])

    def pinned(nbytes):
        ptr = lib.shafa_cuda_host_alloc(int(nbytes))
        if not ptr:
            raise MemoryError("pinned allocation failed")
        return ptr, np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), (int(nbytes),))

    p_in, a_in = pinned(n); a_in[:] = data
    p_rle, a_rle = pinned(n // 8 + 64 * MIB)
    p_pay, a_pay = pinned(n // 8 + 64 * MIB)
    p_out, a_out = pinned(n + 4096)
    res = {"workload": desc.format(gib=f"{n / GIB:g} GiB"), "devices": world, "blocks": (n + bs - 1) // bs, "scaling": "strong",
           "api": "shafa_cuda_encode_host_multi + shafa_cuda_decode_host_multi, one host process, pinned host buffers, no collective"}
    with pkg.ShafaMulti(list(range(world))) as m:
        times = []
        for it in range(3):
            t0 = time.perf_counter()
            c = m.encode_raw(a_in, bs, rle_buf=a_rle, payload_buf=a_pay)
            t1 = time.perf_counter()
            B = int(c.n_blocks)
            z = np.ctypeslib.as_array(c.z, (B,)); po = np.ctypeslib.as_array(c.payload_offsets, (B,)); ss = np.ctypeslib.as_array(c.sym_sizes, (B,))
            ln = np.ctypeslib.as_array(c.code_len, (B * 256,)); cd = np.ctypeslib.as_array(c.code_bits, (B * 256,))
            out, _ = m.decode_tables(a_pay[: int(c.payload_total)], po, z, ss, ln, cd, bool(c.rle_mode), n, out=a_out)
            t2 = time.perf_counter()
            if it == 0:
                if not np.array_equal(out[:n], data):
                    raise RuntimeError("one-file multi-GPU round trip mismatch")
                if case:
                    parts = [f"@{B}".encode()]
                    for b in range(B):
                        parts += [f"@{int(z[b])}@".encode(), a_pay[int(po[b]):int(po[b]) + int(z[b])]]
                    got = {".rle": sha_arrays(a_rle[: int(c.rle_total)]), ".rle.shaf": sha_arrays(*parts)}
                    res["parity"] = {k: ("identical to the reference" if got[k] == case["artefacts"][k] else "MISMATCH") for k in got}
                    if any(v == "MISMATCH" for v in res["parity"].values()):
                        raise RuntimeError(f"one-file multi-GPU artefacts differ from the reference: {res['parity']}")
            else:
                times.append((t1 - t0, t2 - t1))
            m.encoded_free(c)
        res["launches"] = m.launch_count()
    enc_s = sum(t[0] for t in times) / len(times)
    dec_s = sum(t[1] for t in times) / len(times)
    res.update(value=round(n / (enc_s + dec_s) / 1e9, 3), unit="GB/s", encode_gbs=round(n / enc_s / 1e9, 3), decode_gbs=round(n / dec_s / 1e9, 3))
    for ptr in (p_in, p_rle, p_pay, p_out):
        lib.shafa_cuda_host_free(ptr)
    return res


def run_ours(args):
    import torch
    import torch.distributed as dist
    from __graft_entry__ import load_package

    pkg = load_package()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        cpu_group = dist.new_group(backend="gloo")   # waits that must not keep a GPU busy (the one-file leg)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    gen_name, _, default_n, fflags, _, _ = WORKLOADS[args.workload]
    n = args.bytes or default_n
    cfg = workload_config(pkg, args.workload, n, world)
    bs = cfg["block_size"]
    force_freq = "f" in fflags
    decode_only = args.workload == "C5"
    data = pkg.datagen.GENERATORS[gen_name](n, seed=SEEDS[gen_name] + rank)      # this rank's shard of blocks
    case = golden_case(args.workload, n) if rank == 0 else None
    B = (n + bs - 1) // bs

    clocks = ClockSampler(local_rank)   # started early: nvidia-smi takes a while to deliver its first line
    sh = pkg.Shafa(device=local_rank)
    stream = torch.cuda.current_stream(dev)
    sh.set_stream(stream.cuda_stream)

    # ---- device-resident buffers (torch is only the allocator / stream / event provider) ----
    def dbuf(nbytes):
        return torch.empty(int(nbytes + 256), dtype=torch.uint8, device=dev)

    h_in = torch.from_numpy(data)
    d_in = dbuf(n); d_in[:n].copy_(h_in)
    # 1.25 n covers every BASELINE config in R mode (C2 shrinks by 10 %, C3 40x); the worst case is 2n and the call reports it
    rle_cap = n + n // 4 + 3 * B + 64
    d_rle = dbuf(rle_cap)
    d_hist = dbuf(B * 256 * 8); d_len = dbuf(B * 256); d_code = dbuf(B * 256 * 8)
    d_hist_orig = dbuf(B * 256 * 8) if force_freq else None
    pay_cap = n + n // 8 + 32 * B + 4096
    d_payload = dbuf(pay_cap)
    d_sym2 = dbuf(rle_cap)
    d_out = dbuf(n)
    torch.cuda.synchronize(dev)
    info = {}
    p = lambda t: t.data_ptr()
    block_offs = np.minimum(np.arange(0, B + 1, dtype=np.uint64) * np.uint64(bs), np.uint64(n))

    def encode(ev_mid=None):
        first = min(n, bs)
        rle0 = int(sh.rle_encode_blocks(p(d_in), first, bs)[0])              # F: block-0 probe (f.c:246-258)
        rle_mode = sh.rle_gate(first, rle0)
        if rle_mode:
            sym_sizes = sh.rle_encode_blocks(p(d_in), n, bs, p(d_rle), rle_cap, p(d_hist))   # F: RLE + fused histogram
            d_sym = d_rle
            if force_freq:
                sh.histogram_blocks(p(d_in), block_offs, p(d_hist_orig))                      # F: -c f, the original's .freq too
        else:
            sym_sizes = np.diff(block_offs)
            d_sym = d_in
        offs = np.concatenate([[0], np.cumsum(sym_sizes)]).astype(np.uint64)
        if not rle_mode:
            sh.histogram_blocks(p(d_in), offs, p(d_hist))                                      # F: histogram
        sh.sf_build_codes(p(d_hist), B, p(d_len), p(d_code))                                   # T
        poff, z = sh.sf_encode_blocks(p(d_sym), offs, p(d_len), p(d_code), p(d_hist), p(d_payload), pay_cap)   # C
        if ev_mid is not None:
            ev_mid.record(stream)
        info.update(rle_mode=rle_mode, S=int(offs[-1]), Z=int(z.sum()), R=int(offs[-1]) if rle_mode else 0,
                    sym_sizes=sym_sizes, offs=offs, poff=poff, z=z)

    def decode(ev_mid=None):
        sh.sf_decode_blocks(p(d_payload), info["poff"], info["z"], info["sym_sizes"], p(d_len), p(d_code), p(d_sym2), int(info["offs"][-1]) + 64)  # D (SF)
        if ev_mid is not None:
            ev_mid.record(stream)
        if info["rle_mode"]:
            fin = sh.rle_decode_blocks(p(d_sym2), info["offs"], p(d_out), n)                    # D (RLE)
            assert int(fin.sum()) == n

    def step(ev_mid=None):
        if decode_only:
            decode(ev_mid)        # ev_mid splits -d s from -d r
        else:
            encode(ev_mid)        # ev_mid splits encode from decode
            decode()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    if decode_only:
        encode()                  # the .rle.shaf to decode is produced once, outside every timed region
    for _ in range(args.warmup):
        step()
    barrier()
    # ---- verification, outside the timed region: round trip, and the artefacts against the reference's fingerprints ----
    result = (d_out if info["rle_mode"] else d_sym2)[:n].cpu().numpy()
    if not np.array_equal(result, data):
        raise RuntimeError("round trip mismatch on the GPU path")
    del result
    parity = None
    if case is not None:
        art = case["artefacts"]
        pre = ".rle" if info["rle_mode"] else ""
        payload_host = d_payload[: int(info["poff"][-1]) + int(info["z"][-1])].cpu().numpy()
        parts = [f"@{B}".encode()]
        for b in range(B):
            o, zz = int(info["poff"][b]), int(info["z"][b])
            parts += [f"@{zz}@".encode(), payload_host[o:o + zz]]
        got = {pre + ".shaf": sha_arrays(*parts)}
        if info["rle_mode"]:
            got[".rle"] = sha_arrays(d_rle[: info["R"]].cpu().numpy())
        hist = d_hist[: B * 256 * 8].cpu().numpy().view(np.uint64).reshape(B, 256)
        ln = d_len[: B * 256].cpu().numpy().reshape(B, 256)
        cd = d_code[: B * 256 * 8].cpu().numpy().view(np.uint64).reshape(B, 256)
        mode = "R" if info["rle_mode"] else "N"
        got[pre + ".freq"] = sha_arrays(pkg.freq_text(mode, info["sym_sizes"], hist))
        got[pre + ".cod"] = sha_arrays(pkg.cod_text(mode, info["sym_sizes"], ln, cd))
        parity = {k: ("identical to the reference" if v == art[k] else "MISMATCH") for k, v in got.items()}
        parity["decoded"] = "identical to the input (whose sha256 " + ("matches" if sha_arrays(data) == case["input"] else "DIFFERS from") + " the fingerprinted input)"
        parity["source"] = f"tests/golden/full_cases.json[{WORKLOADS[args.workload][4]}] (oracle/_ref/shafa, the unmodified reference)"
        del payload_host
        if any(v == "MISMATCH" for v in parity.values()):
            raise RuntimeError(f"artefacts of the timed step differ from the reference: {parity}")

    sh.profile(True); sh.profile_reset()
    clocks.wait_ready()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    mids = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    launches0 = sh.launch_count()
    barrier()
    w0 = time.time()
    ev0.record(stream)
    for i in range(args.steps):
        starts[i].record(stream)
        step(mids[i])
    ev1.record(stream)
    barrier()
    clocks.window(w0, time.time())
    total_ms = ev0.elapsed_time(ev1)
    launches = sh.launch_count() - launches0
    kstats = sh.profile_read()
    sh.profile(False)
    first_ms = sum(starts[i].elapsed_time(mids[i]) for i in range(args.steps)) / args.steps     # encode (or -d s for C5)
    second_ms = total_ms / args.steps - first_ms                                                 # decode (or -d r for C5)

    # ---- e2e: host buffers through the C-ABI, copies inside the timed region ----
    e2e_steps = max(1, min(args.steps, 3))
    lib = sh.lib
    import ctypes as C

    def pinned(nbytes):
        ptr = lib.shafa_cuda_host_alloc(int(nbytes))
        if not ptr:
            raise MemoryError("pinned allocation failed")
        return ptr, np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), (int(nbytes),))

    sh.set_stream(None)
    S, Z = info["S"], info["Z"]
    pin_in_p, pin_in = pinned(n); pin_in[:] = data
    pin_rle_p, pin_rle = pinned(S + 4096)
    pin_pay_p, pin_pay = pinned(Z + 32 * B + 4096)
    pin_out_p, pin_out = pinned(n + 4096)
    e2e_times, h2d, d2h = [], 0, 0
    c_keep = None
    for i in range(1 + e2e_steps):
        barrier()
        w0 = time.time()
        t0 = time.perf_counter()
        if not decode_only or c_keep is None:
            c = sh.encode_raw(pin_in, bs, force_freq=force_freq, rle_buf=pin_rle, payload_buf=pin_pay)
            if decode_only:
                c_keep = c
                t0 = time.perf_counter()
        else:
            c = c_keep
        Bc = int(c.n_blocks)
        z_arr = np.ctypeslib.as_array(c.z, (Bc,)); po_arr = np.ctypeslib.as_array(c.payload_offsets, (Bc,))
        ss_arr = np.ctypeslib.as_array(c.sym_sizes, (Bc,))
        ln_arr = np.ctypeslib.as_array(c.code_len, (Bc * 256,)); cd_arr = np.ctypeslib.as_array(c.code_bits, (Bc * 256,))
        out, _ = sh.decode_tables(pin_pay[: int(c.payload_total)], po_arr, z_arr, ss_arr, ln_arr, cd_arr, bool(c.rle_mode), n, out=pin_out)
        t1 = time.perf_counter()
        if i == 0:
            if not np.array_equal(out[:n], data):
                raise RuntimeError("e2e round trip mismatch")
        else:
            e2e_times.append(t1 - t0)
            clocks.window(w0, time.time())
        if decode_only:
            h2d = int(c.payload_total) + Bc * 256 * 9
            d2h = n
        else:
            h2d = n + int(c.payload_total) + Bc * 256 * 9
            d2h = int(c.rle_total) + int(c.payload_total) + Bc * 256 * (8 + 9) + n
            sh.encoded_free(c)
    if c_keep is not None:
        sh.encoded_free(c_keep)
    e2e_s = sum(e2e_times) / len(e2e_times)
    clk = clocks.stop()

    # ---- reduce over ranks (max time) ----
    t = torch.tensor([total_ms, e2e_s * 1e3, first_ms, second_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, first_ms, second_ms = [float(x) for x in t.tolist()]

    for ptr in (pin_in_p, pin_rle_p, pin_pay_p, pin_out_p):
        lib.shafa_cuda_host_free(ptr)
    # free this rank's device buffers before the one-file leg puts its own shards on every GPU
    del d_in, d_rle, d_payload, d_sym2, d_out
    torch.cuda.empty_cache()

    one_file = None
    if world > 1 and not args.no_one_file and args.workload == "C2":
        # rank 0 drives all N GPUs; the others wait on a CPU (gloo) barrier -- an NCCL barrier would spin a kernel on
        # their GPU for the whole leg and halve what rank 0 gets out of it (measured)
        barrier()
        dist.barrier(group=cpu_group)
        if rank == 0:
            try:
                one_file = one_file_multi(pkg, world, lib)
            except Exception as e:  # noqa: BLE001 -- reported; the headline line must still be printed
                one_file = {"error": f"{type(e).__name__}: {e}"}
        dist.barrier(group=cpu_group)

    if rank == 0:
        peak, peak_src = measured_peak_hbm()
        R = info["R"]
        algo = {"rle_encode": n + R, "rle_encode_sizes": min(n, bs), "histogram": S, "sf_encode": S + Z, "dec_fused": Z + S,
                "dec_sync": Z, "dec_write": Z + S, "rle_decode": R + n}
        # The synchronisation pass is two kernels over disjoint sets of blocks (k_dec_sync, and k_dec_sync_lean for complete
        # codes of nearly one length): a kernel that ran alone is credited with Z, a mixed file only with their sum.
        sync_ms = {k: kstats[k]["ms"] for k in ("dec_sync", "dec_sync_lean") if k in kstats and kstats[k]["launches"]}
        if len(sync_ms) == 2:
            small, big = sorted(sync_ms, key=sync_ms.get)
            if sync_ms[small] < 0.1 * sync_ms[big]:    # the other one only looked at its (empty) class
                algo["dec_sync_lean" if big == "dec_sync_lean" else "dec_sync"] = Z
                algo.pop(small, None)
            else:
                algo.pop("dec_sync", None)
                algo["dec_sync (both kernels)"] = Z
                kstats = dict(kstats)
                kstats["dec_sync (both kernels)"] = {"ms": sum(sync_ms.values()), "launches": kstats["dec_sync"]["launches"]}
        elif "dec_sync_lean" in sync_ms:
            algo["dec_sync_lean"] = Z
        kernels = {}
        for name, st in kstats.items():
            per_step_ms = st["ms"] / args.steps
            ent = {"launches_per_step": st["launches"] / args.steps, "ms_per_step": round(per_step_ms, 4)}
            if name in algo and st["launches"]:
                avg_ms = st["ms"] / st["launches"]
                per_launch_bytes = algo[name]   # one launch per step for these kernels
                ent["algorithmic_bytes"] = ALGO_BYTES_DOC[name]
                ent["gbs"] = round(per_launch_bytes / (avg_ms * 1e-3) / 1e9, 1) if avg_ms > 0 else None
                ent["frac_of_hbm"] = round(ent["gbs"] / peak, 4) if ent["gbs"] else None
            kernels[name] = ent
        dom = max((k for k in kernels if k in algo), key=lambda k: kernels[k]["ms_per_step"])
        ms_per_step = total_ms / args.steps
        value = world * n / (ms_per_step * 1e-3) / 1e9
        # pipeline-level roofline: algorithmic bytes of the staged pipeline: F (N+R | N) + C (S+Z) + D (Z+S [+ R+N])
        enc_bytes = (n + R) + (S + Z)
        dec_bytes = (Z + S) + ((R + n) if info["rle_mode"] else 0)
        pipe_bytes = dec_bytes if decode_only else enc_bytes + dec_bytes
        full_c2 = args.workload == "C2" and n == GIB
        traffic, traffic_src = (profiled_traffic(dom) if full_c2 else (None, "captured on the C2 1 GiB workload only"))
        cfg.update(mode="R" if info["rle_mode"] else "N", S_symbol_stream_bytes=S, Z_payload_bytes=Z)
        gbs = lambda nbytes, ms: round(world * nbytes / (ms * 1e-3) / 1e9, 3) if ms > 0 else None
        line = {
            "metric": METRIC, "value": round(value, 3), "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic", "config": cfg,
            "pipeline_hbm": {"algorithmic_bytes": pipe_bytes, "formula": "(Z+S) + (R+N)" if decode_only else "(N+R) + (S+Z) + (Z+S) + (R+N)",
                             "achieved_gbs": round(pipe_bytes / (ms_per_step * 1e-3) / 1e9, 1),
                             "frac": round(pipe_bytes / (ms_per_step * 1e-3) / 1e9 / peak, 4)},
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["gbs"], "peak": peak, "unit": "GB/s",
                         "frac": kernels[dom]["frac_of_hbm"], "frac_of_nominal_8000": round(kernels[dom]["gbs"] / 8000.0, 4), "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "algorithmic_bytes": ALGO_BYTES_DOC[dom], "algorithmic_bytes_per_launch": algo[dom]},
            "kernels": kernels,
            "e2e": {"value": round(world * n / (e2e_ms * 1e-3) / 1e9, 3), "unit": "GB/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "api": ("shafa_cuda_decode_host" if decode_only else "shafa_cuda_encode_host + shafa_cuda_decode_host") + " on pinned host buffers"},
            "gpu_launches": int(launches), "clocks": clk,
        }
        if decode_only:
            line.update(decode_sf_gbs_out=gbs(S, first_ms), decode_sf_gbs_compressed_in=gbs(Z, first_ms), decode_rle_gbs_out=gbs(n, second_ms),
                        decode_gbs=gbs(n, ms_per_step), decode_gbs_compressed_in=gbs(Z, ms_per_step),
                        decode_hbm_frac={"-d s (Z+S)": round((Z + S) / (first_ms * 1e-3) / 1e9 / peak, 4),
                                         "-d r (R+N)": round((R + n) / (second_ms * 1e-3) / 1e9 / peak, 4) if second_ms > 0 else None})
        else:
            line.update(encode_gbs=gbs(n, first_ms), decode_gbs=gbs(n, second_ms), decode_gbs_compressed_in=gbs(Z, second_ms),
                        direction_hbm_frac={"encode (N+R)+(S+Z)": round(enc_bytes / (first_ms * 1e-3) / 1e9 / peak, 4),
                                            "decode (Z+S)+(R+N)": round(dec_bytes / (second_ms * 1e-3) / 1e9 / peak, 4)})
        if parity is not None:
            line["parity"] = parity
        if one_file is not None:
            line["one_file"] = one_file
        if world == 1 and not args.no_files and not decode_only:
            try:
                line["e2e_files"] = cli_files_round_trip(data, WORKLOADS[args.workload][1], fflags)
            except Exception as e:  # noqa: BLE001
                line["e2e_files"] = {"error": f"{type(e).__name__}: {e}"}
        if world == 1 and not args.no_cpu and full_c2:
            times, desc, nbytes = cpu_arm(pkg, 1, 0, per_core=True, sample=data)
            desc["value"] = round(nbytes / times[0] / 1e9, 4)
            desc["unit"] = "GB/s"
            line["cpu_baseline"] = desc
        print(json.dumps(line), flush=True)
    sh.close()
    if world > 1:
        dist.destroy_process_group()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from __graft_entry__ import load_package

    pkg = load_package()
    times, desc, nbytes = cpu_arm(pkg, args.steps, args.warmup)
    sec = sum(times) / len(times)
    v = round(nbytes / sec / 1e9, 4)
    desc.update(value=v, unit="GB/s")
    cfg = workload_config(pkg, "C2", nbytes, args.gpus)
    sizes = desc.get("artefact_sizes") or {}
    if ".rle" in sizes:
        cfg.update(mode="R", S_symbol_stream_bytes=sizes[".rle"])
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(sec * 1e3, 2), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": cfg,
        "cpu_baseline": desc,
        "e2e": {"value": v, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--bytes", type=int, default=0, help="bytes per GPU (default: the size BASELINE.json names for the workload)")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="C2", help="input shape (default C2, the one the metric is quoted on)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-files", action="store_true", help="skip the e2e_files leg (CLI, file to file)")
    ap.add_argument("--no-one-file", action="store_true", help="N > 1: skip the one-file-over-N-GPUs leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_ours(args)


if __name__ == "__main__":
    main()
