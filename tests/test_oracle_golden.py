"""Pin the oracle (CPU restatement) to the fixtures produced by the unmodified
reference binary (tests/golden/make_golden.py).  CPU only."""
from __future__ import annotations

import hashlib
import json

import numpy as np
import pytest

from conftest import GOLDEN, encode_only_cases, small_cases

BLOCKS = {"default": 65536, "K": 655360, "m": 8388608, "M": 67108864}


def _check_pipeline(orc, data, block_size, fflags, artefact):
    """artefact(suffix) -> expected bytes or None."""
    f = orc.module_f(data, block_size, force_rle="r" in fflags[1::2], force_freq="f" in fflags[1::2])
    want_rle = artefact(".rle")
    assert f["rle_mode"] == (want_rle is not None)
    if f["rle_mode"]:
        assert f["rle"] == want_rle
        assert f["rle_freq"] == artefact(".rle.freq")
        sym, freq, base = f["rle"], f["rle_freq"], ".rle"
    else:
        assert f["freq"] == artefact(".freq")
        sym, freq, base = data.tobytes(), f["freq"], ""
    if f["rle_mode"] and artefact(".freq") is not None:  # -c f
        assert f["freq"] == artefact(".freq")
        assert orc.module_t(f["freq"]) == artefact(".cod")
    cod = orc.module_t(freq)
    assert cod == artefact(base + ".cod")
    shaf = orc.module_c(sym, cod)
    assert shaf == artefact(base + ".shaf")
    # decoders (d.c) on the reference's own artefacts
    assert orc.module_d_shaf(shaf, cod, rle=False) == sym
    if f["rle_mode"]:
        assert orc.module_d_rle(sym, freq) == data.tobytes()
        assert orc.module_d_shaf(shaf, cod, rle=True) == data.tobytes()


@pytest.mark.parametrize("case", small_cases())
def test_small_case(orc, case):
    d = GOLDEN / case
    flags = json.loads((d / "flags.json").read_text())
    data = np.frombuffer((d / "input.bin").read_bytes(), dtype=np.uint8)

    def artefact(suffix):
        p = d / ("input.bin" + suffix)
        return p.read_bytes() if p.exists() else None

    _check_pipeline(orc, data, BLOCKS[flags["block"]], flags["fflags"], artefact)


def test_big_cases(orc, pkg):
    big = json.loads((GOLDEN / "big_cases.json").read_text())
    for name, spec in big.items():
        data = pkg.datagen.GENERATORS[spec["generator"]](spec["n"], seed=spec["seed"])
        assert hashlib.sha256(data.tobytes()).hexdigest() == spec["input_sha256"], name
        seen = {}

        def artefact(suffix, spec=spec, seen=seen):
            a = spec["artefacts"].get(suffix)
            if a is None:
                return None

            class _H:  # compares equal to bytes with the recorded size+hash
                def __eq__(self, other):
                    seen[suffix] = True
                    return len(other) == a["size"] and hashlib.sha256(other).hexdigest() == a["sha256"]

                __hash__ = None

            return _H()

        _check_pipeline(orc, data, BLOCKS[spec["block"]], spec["fflags"], artefact)
        assert set(seen) == set(spec["artefacts"]), name


def test_t_cases(orc):
    freq = (GOLDEN / "tcases.freq").read_bytes()
    assert orc.module_t(freq) == (GOLDEN / "tcases.cod").read_bytes()


def test_rle_gate_float32(orc):
    # f.c:253-257: float32 quotient vs double 0.05 -- 5 % exactly is NOT < 0.05 in float->double
    assert orc.rle_gate(1000, 951) is False      # 4.9 %
    assert orc.rle_gate(1000, 949) is True       # 5.1 %
    assert orc.rle_gate(1000, 951, force_rle=True) is True
    assert orc.rle_gate(655360, 655360 + 5000) is False  # negative saving
    n0 = 655360
    for rle0 in range(n0 - 33000, n0 - 32500):
        want = not (np.float32(np.float32(n0 - rle0) / np.float32(n0)) < 0.05)
        assert orc.rle_gate(n0, rle0) == want


def test_block_primitives_roundtrip(orc, pkg):
    rng = np.random.default_rng(3)
    for gen in ("zipf", "textlike", "runheavy", "uniform"):
        data = pkg.datagen.GENERATORS[gen](50_000, seed=int(rng.integers(1 << 30)))
        rle = orc.rle_encode_block(data)
        assert np.array_equal(orc.rle_decode_block(rle), data)
        codes = orc.sf_codes(orc.histogram(data))
        payload = orc.sf_encode_block(data, codes)
        assert np.array_equal(orc.sf_decode_block(payload, data.size, codes), data)
        ln, cd, mx = orc.sf_codes_packed(orc.histogram(data))
        for s, c in enumerate(codes):
            assert ln[s] == len(c) and (not c or int(c, 2) == int(cd[s]))
        assert mx == max(len(c) for c in codes)


@pytest.mark.parametrize("case", encode_only_cases())
def test_one_symbol_block_encode(orc, case):
    """A block with one distinct symbol: empty code (t.c:185), "@0@" in the .shaf (c.c:65-80)."""
    d = GOLDEN / "encode_only" / case
    flags = json.loads((d / "flags.json").read_text())
    data = np.frombuffer((d / "input.bin").read_bytes(), dtype=np.uint8)
    f = orc.module_f(data, 65536, force_rle="r" in flags["fflags"][1::2], force_freq=False)
    base = "input.bin.rle" if f["rle_mode"] else "input.bin"
    assert f["rle_mode"] == (d / "input.bin.rle").exists()
    if f["rle_mode"]:
        assert f["rle"] == (d / "input.bin.rle").read_bytes()
    freq = f["rle_freq"] if f["rle_mode"] else f["freq"]
    assert freq == (d / (base + ".freq")).read_bytes()
    cod = orc.module_t(freq)
    assert cod == (d / (base + ".cod")).read_bytes()
    shaf = orc.module_c(f["rle"] if f["rle_mode"] else data.tobytes(), cod)
    assert shaf == (d / (base + ".shaf")).read_bytes()
    assert b"@0@" in shaf[-8:]


def test_wide_codes_fixture(orc, pkg):
    """tests/golden/wide_codes (make_wide_golden.py): a hand-made .cod with codes of up to 255 bits and the unmodified
    reference's .shaf for it.  The oracle packs and decodes it like the reference; the host-side parser of the wide
    form reads the same table."""
    d = GOLDEN / "wide_codes"
    data, cod, shaf = (d / "x").read_bytes(), (d / "x.cod").read_bytes(), (d / "x.shaf").read_bytes()
    assert orc.module_c(data, cod) == shaf
    assert orc.module_d_shaf(shaf, cod, rle=False) == data
    mode, sizes, ln, c4 = pkg.parse_cod_wide(cod)
    assert mode == "N" and int(sizes.sum()) == len(data) and int(ln.max()) == 255
    fields = cod.split(b"@")[4].split(b";")   # "", N, 3, size, fields, ...
    for s in range(256):
        v = sum(int(c4[0, s, k]) << (64 * k) for k in range(4))
        assert ln[0, s] == len(fields[s]) and v == (int(fields[s], 2) if fields[s] else 0)
    with pytest.raises(pkg.ShafaError):
        pkg.parse_cod(cod)   # the 64-bit form refuses it
