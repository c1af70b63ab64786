"""CPU-side checks of the boundary: the C-ABI library loads and exports every
symbol include/shafa_cuda.h declares; host-only helpers; emulator of the kernels'
tile decomposition (tests/cpu/emu.cpp) against the oracle."""
from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def test_library_exports_declared_symbols(pkg):
    syms = pkg.declared_symbols()
    assert len(syms) >= 25
    lib = pkg.load_library()   # raises if any declared symbol is missing
    for s in syms:
        assert hasattr(lib, s), s


def test_no_cpu_fallback(pkg):
    lib = pkg.load_library()
    if lib.shafa_cuda_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.ShafaError):
        pkg.Shafa(device=0)
    ctx = ctypes.c_void_p()
    assert lib.shafa_cuda_create(0, ctypes.byref(ctx)) != 0


def test_rle_gate_matches_oracle(pkg, orc):
    lib = pkg.load_library()
    rng = np.random.default_rng(0)
    for n0 in (1024, 65536, 655360, 8388608, 67108864):
        for rle0 in list(rng.integers(int(n0 * 0.9), int(n0 * 1.1), size=300)) + [n0, int(n0 * 0.95), int(n0 * 0.95) + 1]:
            for force in (0, 1):
                assert bool(lib.shafa_cuda_rle_gate(n0, int(rle0), force)) == orc.rle_gate(n0, int(rle0), bool(force))


def test_text_helpers_match_oracle(pkg, orc):
    rng = np.random.default_rng(1)
    data = pkg.datagen.textlike(200_000, seed=9)
    B, bs = 4, 65536
    sizes = [min(bs, data.size - b * bs) for b in range(B)]
    hist = np.stack([orc.histogram(data[b * bs:b * bs + sizes[b]]) for b in range(B)])
    # .freq text against the oracle's module F in forced N-ish mode
    f = orc.module_f(data[: sum(sizes)], bs, force_freq=True)
    assert pkg.freq_text("N", sizes, hist) == f["freq"]
    ln = np.zeros((B, 256), np.uint8)
    cd = np.zeros((B, 256), np.uint64)
    for b in range(B):
        ln[b], cd[b], _ = orc.sf_codes_packed(hist[b])
    cod = pkg.cod_text("N", sizes, ln, cd)
    assert cod == orc.module_t(f["freq"])
    mode, sz2, ln2, cd2 = pkg.parse_cod(cod)
    assert mode == "N" and list(sz2) == sizes and np.array_equal(ln, ln2) and np.array_equal(cd, cd2)
    shaf = orc.module_c(data[: sum(sizes)], cod)
    z, payload, offs = pkg.parse_shaf(shaf)
    assert pkg.shaf_bytes(z, payload, offs) == shaf
    _ = rng


def test_kernel_tile_emulator(orc, tmp_path):
    """tests/cpu/emu.cpp drives csrc/core.h through the kernels' tile decomposition."""
    exe = tmp_path / "emu"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wno-unknown-pragmas", "-o", str(exe), str(ROOT / "tests/cpu/emu.cpp"),
                           "-I", str(ROOT), "-L", str(ROOT / "oracle"), "-loracle", f"-Wl,-rpath,{ROOT / 'oracle'}"])
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "EMU OK" in r.stdout, r.stdout[-2000:]
