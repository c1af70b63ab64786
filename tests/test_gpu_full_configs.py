"""BASELINE.json's five configurations at their FULL sizes, compared with fingerprints of the
unmodified reference (tests/golden/full_cases.json, written by tests/golden/make_full_golden.py
from oracle/_ref/shafa): size + sha256 of .rle, .freq / .rle.freq, .cod / .rle.cod, .shaf /
.rle.shaf, and the decode targets (-d s -> the .rle bytes, -d r and -m d -> the input).
The CUDA path is compared with the reference's output, never with itself.

C5 (16 GiB) needs ~60 GiB of host memory for its buffers; on a smaller host the same generator
runs at 4 GiB (its own reference fingerprints) and the test says which size ran.
"""
from __future__ import annotations

import hashlib
import json
import os
import threading
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = Path(__file__).resolve().parent / "golden" / "full_cases.json"
GIB, MIB = 1 << 30, 1 << 20
BLOCKS = {"K": 655360, "m": 8388608, "M": 67108864}


def full_cases():
    return json.loads(GOLDEN.read_text())


def mem_available_gib() -> float:
    try:
        for line in Path("/proc/meminfo").read_text().splitlines():
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) / (1 << 20)
    except OSError:
        pass
    return 0.0


def sha(*chunks) -> dict:
    h = hashlib.sha256()
    n = 0
    for c in chunks:
        if isinstance(c, np.ndarray):
            for lo in range(0, c.size, 256 * MIB):      # bounded slices (hashlib releases the GIL per update)
                h.update(memoryview(c[lo:lo + 256 * MIB]))
            n += c.size
        else:
            h.update(c)
            n += len(c)
    return {"size": n, "sha256": h.hexdigest()}


class EncView:
    """Zero-copy numpy views of a shafa_cuda_encoded struct (binding.Encoded copies everything)."""

    def __init__(self, c):
        B = int(c.n_blocks)
        self.c, self.B = c, B
        self.rle_mode = bool(c.rle_mode)
        as_arr = np.ctypeslib.as_array
        self.block_sizes = as_arr(c.block_sizes, (B,))
        self.sym_sizes = as_arr(c.sym_sizes, (B,))
        self.hist = as_arr(c.hist, (B, 256))
        self.hist_orig = as_arr(c.hist_orig, (B, 256)) if c.hist_orig else None
        self.code_len = as_arr(c.code_len, (B, 256))
        self.code_bits = as_arr(c.code_bits, (B, 256))
        self.z = as_arr(c.z, (B,))
        self.payload_offsets = as_arr(c.payload_offsets, (B,))
        self.payload = as_arr(c.payload, (int(c.payload_total),))
        self.rle = as_arr(c.rle, (int(c.rle_total),)) if self.rle_mode else None

    def shaf_sha(self) -> dict:
        """sha256 of the .shaf file (c.c:351,256-258) without assembling it."""
        parts = [f"@{self.B}".encode()]
        for b in range(self.B):
            o, z = int(self.payload_offsets[b]), int(self.z[b])
            parts.append(f"@{z}@".encode())
            parts.append(self.payload[o:o + z])
        return sha(*parts)


def check_artefacts(pkg, name, case, v: EncView):
    """Every artefact of F, T and C against the reference's fingerprints."""
    art = case["artefacts"]
    mode_r = case["rle_mode"]
    force_freq = "f" in case["fflags"]
    assert v.rle_mode == mode_r, f"{name}: R/N decision differs from the reference (f.c:246-258)"
    pre = ".rle" if mode_r else ""
    results, threads = {}, []
    if mode_r:
        threads.append(threading.Thread(target=lambda: results.__setitem__(".rle", sha(v.rle))))
    threads.append(threading.Thread(target=lambda: results.__setitem__(pre + ".shaf", v.shaf_sha())))
    for t in threads:
        t.start()
    mode = "R" if mode_r else "N"
    results[pre + ".freq"] = sha(pkg.freq_text(mode, v.sym_sizes, v.hist))
    results[pre + ".cod"] = sha(pkg.cod_text(mode, v.sym_sizes, v.code_len, v.code_bits))
    if mode_r and force_freq:
        results[".freq"] = sha(pkg.freq_text("N", v.block_sizes, v.hist_orig))
    for t in threads:
        t.join()
    for suffix, fp in results.items():
        assert fp == art[suffix], f"{name}: {suffix} differs from the reference ({fp} vs {art[suffix]})"
    missing = set(art) - set(results) - ({".cod"} if (mode_r and force_freq) else set())
    assert not missing, f"{name}: artefacts not compared: {missing}"


def encode_and_check(shafa, pkg, name, case, data):
    """F -> T -> C on the GPU through shafa_cuda_encode_host (pipelined path); every artefact vs the reference's."""
    n, bs = data.size, BLOCKS[case["block"]]
    art = case["artefacts"]
    got_input = {}
    t_in = threading.Thread(target=lambda: got_input.__setitem__("fp", sha(data)))
    t_in.start()
    mode_r = case["rle_mode"]
    rle_buf = np.empty((art[".rle"]["size"] + 4096) if mode_r else 4096, dtype=np.uint8)
    pay_buf = np.empty(art[".rle.shaf" if mode_r else ".shaf"]["size"] + 64 * (n // bs + 2) + 4096, dtype=np.uint8)
    c = shafa.encode_raw(data, bs, force_rle="r" in case["fflags"], force_freq="f" in case["fflags"], rle_buf=rle_buf, payload_buf=pay_buf)
    v = EncView(c)
    v.keep = (rle_buf, pay_buf)
    check_artefacts(pkg, name, case, v)
    t_in.join()
    assert got_input["fp"] == case["input"], f"{name}: generator output differs from the fingerprinted input"
    return c, v


def decode_all(shafa, name, v: EncView, data):
    """-d s, -d r and the combined -m d on the GPU, against the reference-checked targets."""
    n = data.size
    out = np.empty(max(n, int(v.sym_sizes.sum())) + 4096, dtype=np.uint8)
    offs = np.concatenate([[0], np.cumsum(v.sym_sizes)]).astype(np.uint64)
    if v.rle_mode:
        sym, fin = shafa.decode_tables(v.payload, v.payload_offsets, v.z, v.sym_sizes, v.code_len, v.code_bits, False, out.size, out=out)
        assert sym.size == v.rle.size and np.array_equal(sym, v.rle), f"{name}: -d s output differs from the reference .rle"
        dec, fin = shafa.rle_decode_host(v.rle, offs, n + 64)
        assert np.array_equal(fin, v.block_sizes) and dec.size == n and np.array_equal(dec, data), f"{name}: -d r"
        del dec
    dec, fin = shafa.decode_tables(v.payload, v.payload_offsets, v.z, v.sym_sizes, v.code_len, v.code_bits, v.rle_mode, out.size, out=out)
    assert dec.size == n and np.array_equal(fin, v.block_sizes) and np.array_equal(dec, data), f"{name}: -m d"


def run_case(shafa, pkg, name):
    case = full_cases()[name]
    data = pkg.datagen.GENERATORS[case["generator"]](case["n"], seed=case["seed"])
    c, v = encode_and_check(shafa, pkg, name, case, data)
    try:
        decode_all(shafa, name, v, data)
    finally:
        shafa.encoded_free(c)


def test_full_c1_zipf_64mib_K(shafa, pkg):
    run_case(shafa, pkg, "C1_zipf_64m_K")


def test_full_c2_text_1gib_m(shafa, pkg):
    run_case(shafa, pkg, "C2_text_1g_m")


def test_full_c3_runs_4gib_M(shafa, pkg):
    run_case(shafa, pkg, "C3_runs_4g_M")


def test_full_c4_uniform_4gib_K_force_freq(shafa, pkg):
    """6,554 blocks of 640 KiB: the per-block loop of c.c:360-421 as one launch."""
    run_case(shafa, pkg, "C4_uniform_4g_K_cf")


def c5_case_name() -> str:
    big = mem_available_gib() >= float(os.environ.get("SHAFA_TEST_C5_MIN_GIB", "90"))
    name = "C5_mixed_16g_m" if big else "C5_mixed_4g_m"
    return name if name in full_cases() else "C5_mixed_4g_m"


def test_full_c5_mixed_decode(shafa, pkg, record_property):
    """configs[4]: 16 GiB mixed, -d s / -d r / -m d (shafa.c:206-257 routing); 4 GiB when the host is small."""
    name = c5_case_name()
    record_property("c5_case", name)
    print(f"[C5] running {name} (MemAvailable {mem_available_gib():.0f} GiB)")
    run_case(shafa, pkg, name)
