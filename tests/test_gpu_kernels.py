"""Parity tests proper: every CUDA entry point of the C-ABI against the oracle on
the same inputs (bit-exact: this is byte/integer work).  Run with -m gpu on a B200."""
from __future__ import annotations

import json

import numpy as np
import pytest

from conftest import GOLDEN, encode_only_cases, small_cases

pytestmark = pytest.mark.gpu

BLOCKS = {"default": 65536, "K": 655360, "m": 8388608, "M": 67108864}


def assert_same(got, want, what):
    got = np.frombuffer(got, dtype=np.uint8) if isinstance(got, (bytes, bytearray)) else np.asarray(got)
    want = np.frombuffer(want, dtype=np.uint8) if isinstance(want, (bytes, bytearray)) else np.asarray(want)
    if got.shape != want.shape:
        n = min(got.size, want.size)
        bad = np.flatnonzero(got.ravel()[:n] != want.ravel()[:n])
        first = int(bad[0]) if bad.size else n
        raise AssertionError(f"{what}: size {got.shape} != {want.shape}; first difference at {first}")
    if not np.array_equal(got, want):
        bad = np.flatnonzero(got.ravel() != want.ravel())
        i = int(bad[0])
        lo = max(0, i - 8)
        raise AssertionError(
            f"{what}: {bad.size} differing elements, first at {i}: got {got.ravel()[lo:i + 8].tolist()} want {want.ravel()[lo:i + 8].tolist()}")


def ragged_offsets(rng, total, nblocks):
    cuts = np.sort(rng.choice(np.arange(1, total), size=nblocks - 1, replace=False)) if nblocks > 1 else np.array([], dtype=np.int64)
    return np.concatenate([[0], cuts, [total]]).astype(np.uint64)


# ------------------------------------------------------------------ histogram (f.c:63-79)
@pytest.mark.parametrize("n,nblocks,lead", [(1, 1, 0), (5000, 1, 0), (100_000, 7, 0), (1_000_003, 13, 37), (3 * 16384, 3, 16384 - 5)])
def test_histogram_blocks(shafa, n, nblocks, lead):
    rng = np.random.default_rng(n)
    data = rng.integers(0, 256, size=lead + n, dtype=np.uint8)
    data[lead: lead + n // 2] = rng.choice([0, 32, 101], size=n // 2).astype(np.uint8)  # skewed half
    offs = ragged_offsets(rng, n, nblocks) + np.uint64(lead)
    d_in = shafa.to_device(data)
    d_hist = shafa.alloc(nblocks * 256 * 8)
    shafa.histogram_blocks(d_in, offs, d_hist)
    got = d_hist.download(np.uint64).reshape(nblocks, 256)
    for b in range(nblocks):
        want = np.bincount(data[int(offs[b]):int(offs[b + 1])], minlength=256).astype(np.uint64)
        assert_same(got[b], want, f"hist block {b}")


# ------------------------------------------------------------------ module T (t.c:74-210)
def test_sf_build_codes_tcases(shafa, orc, pkg):
    freq = (GOLDEN / "tcases.freq").read_bytes()
    blocks = freq.split(b"@")[3:-1]          # '', 'N', 'B', n0, fields0, n1, fields1, ..., '0'
    sizes = [int(x) for x in blocks[0::2]]
    hists = np.stack([orc.freq_parse(f) for f in blocks[1::2]])
    B = len(sizes)
    d_hist = shafa.to_device(hists)
    d_len = shafa.alloc(B * 256)
    d_code = shafa.alloc(B * 256 * 8)
    shafa.sf_build_codes(d_hist, B, d_len, d_code)
    ln = d_len.download(np.uint8).reshape(B, 256)
    cd = d_code.download(np.uint64).reshape(B, 256)
    for b in range(B):
        wl, wc, _ = orc.sf_codes_packed(hists[b])
        assert_same(ln[b], wl, f"code lengths block {b}")
        assert_same(cd[b], wc, f"code bits block {b}")
    # and the container text against the reference's own .cod
    assert pkg.cod_text("N", sizes, ln, cd) == (GOLDEN / "tcases.cod").read_bytes()


def test_sf_build_rejects_int_overflow(shafa, pkg):
    h = np.zeros((1, 256), dtype=np.uint64)
    h[0, :4] = 1 << 30
    d_hist = shafa.to_device(h)
    with pytest.raises(pkg.ShafaError) as e:
        shafa.sf_build_codes(d_hist, 1, shafa.alloc(256), shafa.alloc(2048))
    assert e.value.code == 4


# ------------------------------------------------------------------ RLE encode (f.c:29-55) + fused histogram (f.c:310)
def _check_rle_encode(shafa, orc, data, block_size):
    n = data.size
    B = (n + block_size - 1) // block_size
    d_in = shafa.to_device(data)
    cap = 2 * n + 3 * B + 64
    d_rle = shafa.alloc(cap)
    d_hist = shafa.alloc(B * 256 * 8)
    sizes = shafa.rle_encode_blocks(d_in, n, block_size, d_rle, cap, d_hist)
    want_blocks = [orc.rle_encode_block(data[b * block_size:(b + 1) * block_size]) for b in range(B)]
    assert_same(sizes, np.array([w.size for w in want_blocks], dtype=np.uint64), "rle sizes")
    got = d_rle.download(np.uint8, int(sizes.sum()))
    assert_same(got, np.concatenate(want_blocks), "rle bytes")
    hist = d_hist.download(np.uint64).reshape(B, 256)
    for b in range(B):
        assert_same(hist[b], np.bincount(want_blocks[b], minlength=256).astype(np.uint64), f"fused hist block {b}")
    # sizes-only variant (the block-0 probe)
    assert_same(shafa.rle_encode_blocks(d_in, n, block_size), sizes, "rle sizes (probe)")


@pytest.mark.parametrize("case", small_cases())
def test_rle_encode_golden_inputs(shafa, orc, case):
    data = np.frombuffer((GOLDEN / case / "input.bin").read_bytes(), dtype=np.uint8)
    _check_rle_encode(shafa, orc, data, 65536)


@pytest.mark.parametrize("gen,n,bs", [("runheavy", 3_000_000, 655360), ("textlike", 2_500_000, 655360), ("uniform", 700_001, 65536),
                                      ("zipf", 1_000_000, 4096), ("runheavy", 9_000_000, 8388608), ("mixed4m", 9_000_000, 655360)])
def test_rle_encode_generators(shafa, orc, pkg, gen, n, bs):
    _check_rle_encode(shafa, orc, pkg.datagen.GENERATORS[gen](n, seed=n % 97), bs)


def test_rle_encode_long_runs_across_tiles(shafa, orc):
    data = np.full(1_500_000, 7, dtype=np.uint8)       # one run over hundreds of tiles, cut by block edges
    data[1_000_000:1_000_003] = 0
    _check_rle_encode(shafa, orc, data, 655360)
    z = np.zeros(300_000, dtype=np.uint8)              # all zeros
    z[123_456] = 1
    _check_rle_encode(shafa, orc, z, 65536)


# ------------------------------------------------------------------ module C (c.c:52-83)
def _codes_for(orc, stream, offs):
    B = len(offs) - 1
    ln = np.zeros((B, 256), np.uint8)
    cd = np.zeros((B, 256), np.uint64)
    strs = []
    for b in range(B):
        h = orc.histogram(stream[int(offs[b]):int(offs[b + 1])])
        ln[b], cd[b], mx = orc.sf_codes_packed(h)
        assert mx >= 0
        strs.append(orc.sf_codes(h))
    return ln, cd, strs


def _check_sf_encode_decode(shafa, orc, stream, offs, with_hist):
    B = len(offs) - 1
    ln, cd, strs = _codes_for(orc, stream, offs)
    d_sym = shafa.to_device(stream)
    d_len, d_code = shafa.to_device(ln), shafa.to_device(cd)
    d_hist = None
    if with_hist:
        d_hist = shafa.alloc(B * 256 * 8)
        shafa.histogram_blocks(d_sym, offs, d_hist)
    total = int(offs[-1] - offs[0])
    cap = total * 2 + 64 * B + 4096
    d_payload = shafa.alloc(cap)
    poff, z = shafa.sf_encode_blocks(d_sym, offs, d_len, d_code, d_hist, d_payload, cap)
    payload = d_payload.download(np.uint8)
    for b in range(B):
        want = orc.sf_encode_block(stream[int(offs[b]):int(offs[b + 1])], strs[b])
        assert int(z[b]) == want.size, f"z block {b}: {int(z[b])} != {want.size}"
        assert int(poff[b]) % 16 == 0
        assert_same(payload[int(poff[b]):int(poff[b]) + want.size], want, f"payload block {b}")
    # decode it back on the GPU
    n = (offs[1:] - offs[:-1]).astype(np.uint64)
    d_out = shafa.alloc(total + 64)
    shafa.sf_decode_blocks(d_payload, poff, z, n, d_len, d_code, d_out, total + 64)
    assert_same(d_out.download(np.uint8, total), stream[int(offs[0]):int(offs[-1])], "sf decode")


@pytest.mark.parametrize("gen,n,nblocks,lead,with_hist", [
    ("zipf", 200_000, 3, 0, True), ("textlike", 1_000_000, 5, 4099, False), ("uniform", 300_000, 2, 13, True),
    ("runheavy", 800_000, 4, 0, False), ("zipf", 5_000_000, 9, 777, True), ("textlike", 40, 1, 5, False)])
def test_sf_encode_decode_ragged(shafa, orc, pkg, gen, n, nblocks, lead, with_hist):
    rng = np.random.default_rng(n + nblocks)
    body = pkg.datagen.GENERATORS[gen](n, seed=3)
    if gen == "runheavy":   # a realistic C input in R mode is the RLE stream
        body = orc.rle_encode_block(body)
        n = body.size
    stream = np.concatenate([np.full(lead, 0xAA, np.uint8), body])
    offs = ragged_offsets(rng, n, nblocks) + np.uint64(lead)
    _check_sf_encode_decode(shafa, orc, stream, offs, with_hist)


def test_sf_encode_long_codes(shafa, orc, pkg):
    """Fibonacci counts give codes > 26 bits: the generic encoder and the decoder's sorted
    fallback, next to an ordinary block in the same call."""
    fib = [1, 1]
    while len(fib) < 31:
        fib.append(fib[-1] + fib[-2])
    rng = np.random.default_rng(5)
    deep = np.repeat(np.arange(31, dtype=np.uint8) * 5, fib)
    rng.shuffle(deep)
    plain = pkg.datagen.zipf(300_000, seed=8)
    stream = np.concatenate([deep, plain])
    offs = np.array([0, deep.size, stream.size], dtype=np.uint64)
    ln, _, _ = _codes_for(orc, stream, offs)
    assert ln[0].max() > 26 and ln[1].max() <= 26
    _check_sf_encode_decode(shafa, orc, stream, offs, True)


def test_sf_encode_tile_with_long_codes_everywhere(shafa, orc):
    """A 16 KiB tile made only of rare symbols (>= 13-bit codes) needs more than one staging window."""
    rng = np.random.default_rng(3)
    common = np.repeat(np.arange(6, dtype=np.uint8), [1_500_000, 750_000, 375_000, 187_000, 94_000, 47_000])
    rng.shuffle(common)
    rare = np.arange(10, 250, dtype=np.uint8)
    body = np.concatenate([common[:1_000_003], np.tile(rare, 170), common[1_000_003:]])
    offs = np.array([0, body.size], dtype=np.uint64)
    ln, _, _ = _codes_for(orc, body, offs)
    assert 12 < ln[0][rare].min() and ln[0].max() <= 26
    _check_sf_encode_decode(shafa, orc, body, offs, False)


def test_sf_encode_codes_between_17_and_26_bits(shafa, orc):
    """Halving counts give a deep code (max 17..26 bits): the accumulator path that takes one code at a time,
    with rare long codes sprinkled through otherwise ordinary tiles and ragged block edges."""
    rng = np.random.default_rng(17)
    counts = [max(1, 600_000 >> i) for i in range(24)]
    body = np.repeat((np.arange(24, dtype=np.uint8) * 7 + 3), counts)
    rng.shuffle(body)
    offs = np.array([0, 5, 400_001, body.size], dtype=np.uint64)
    ln, _, _ = _codes_for(orc, body, offs)
    assert 16 < ln[1].max() <= 26 and 16 < ln[2].max() <= 26
    _check_sf_encode_decode(shafa, orc, body, offs, True)


def test_sf_encode_symbol_without_code(shafa, orc, pkg):
    data = pkg.datagen.zipf(50_000, seed=1)
    offs = np.array([0, data.size], dtype=np.uint64)
    ln, cd, _ = _codes_for(orc, data, offs)
    present = np.flatnonzero(ln[0])
    ln[0, present[0]] = 0
    with pytest.raises(pkg.ShafaError) as e:
        shafa.sf_encode_blocks(shafa.to_device(data), offs, shafa.to_device(ln), shafa.to_device(cd), None, shafa.alloc(200_000), 200_000)
    assert e.value.code == 4


# ------------------------------------------------------------------ module D
def test_sf_decode_reference_payloads(shafa, orc, pkg):
    """Decode the reference's own .shaf fixtures with the reference's .cod tables."""
    for case in small_cases():
        d = GOLDEN / case
        base = "input.bin.rle" if (d / "input.bin.rle").exists() else "input.bin"
        mode, sizes, ln, cd = pkg.parse_cod((d / (base + ".cod")).read_bytes())
        z, payload, offs = pkg.parse_shaf((d / (base + ".shaf")).read_bytes())
        sym, _ = shafa.decode_tables(payload, offs, z, sizes, ln, cd, False, int(sizes.sum()) + 64)
        assert_same(sym, (d / base).read_bytes(), f"{case}: -d s")
        if mode == "R":
            orig = (d / "input.bin").read_bytes()
            out, fin = shafa.decode_tables(payload, offs, z, sizes, ln, cd, True, len(orig) + 64)
            assert_same(out, orig, f"{case}: -m d")


@pytest.mark.parametrize("alphabet,n", [(2, 6_000_000), (4, 3_000_000), (6, 2_500_000), (24, 1_600_000), (200, 900_000), (256, 700_000)])
def test_sf_decode_codes_of_nearly_one_length(shafa, orc, pkg, alphabet, n):
    """Uniformly drawn symbols give complete codes of nearly one length (1-bit codes for two symbols, 7..9 bits for 256):
    k_dec_classes / k_dec_lut16 / k_dec_sync_lean for the block large enough to be "lean" (>= 256 KiB of payload), the
    all-starts path inside k_dec_sync for the small one behind it, both in one call; against the oracle's decode."""
    rng = np.random.default_rng(alphabet)
    syms = rng.permutation(256)[:alphabet].astype(np.uint8)
    big = syms[rng.integers(0, alphabet, size=n)]
    small = syms[rng.integers(0, alphabet, size=40_000)]
    lens, codes, payloads = [], [], []
    for blk in (big, small):
        h = orc.histogram(blk)
        ln, cd, _ = orc.sf_codes_packed(h)
        lens.append(ln); codes.append(cd)
        payloads.append(orc.sf_encode_block(blk, orc.sf_codes(h)))
    assert payloads[0].size >= 4 * 65536 > payloads[1].size
    used = lens[0][lens[0] > 0]
    assert 2 * int(used.min()) > int(used.max()) and int(used.max()) <= 12, "not the code shape this test is about"
    z = np.array([p.size for p in payloads], dtype=np.uint64)
    slot0 = (int(z[0]) + 15) & ~15
    offs = np.array([0, slot0], dtype=np.uint64)
    buf = np.zeros(slot0 + ((int(z[1]) + 15) & ~15) + 32, dtype=np.uint8)
    buf[: int(z[0])] = payloads[0]
    buf[slot0: slot0 + int(z[1])] = payloads[1]
    sizes = np.array([big.size, small.size], dtype=np.uint64)
    shafa.profile(True); shafa.profile_reset()
    sym, _ = shafa.decode_tables(buf, offs, z, sizes, np.stack(lens), np.stack(codes), False, int(sizes.sum()) + 64)
    stats = shafa.profile_read()
    shafa.profile(False)
    assert_same(sym, np.concatenate([big, small]), f"{alphabet} symbols")
    assert "dec_sync_lean" in stats and "dec_lut16" in stats, stats.keys()


def _wide_table(rng, nsym, depths):
    """codes '1' * d + '0' for the given depths plus '1' * max(depths + 1): complete, prefix-free, up to 255 bits"""
    syms = rng.permutation(256)[:nsym]
    codes = [""] * 256
    for k, dpt in enumerate(depths):
        codes[syms[k]] = "1" * dpt + "0"
    codes[syms[len(depths)]] = "1" * (max(depths) + 1)
    return syms, codes


def _codes_to_wide(codes):
    ln = np.array([len(c) for c in codes], dtype=np.uint8)
    c4 = np.zeros((256, 4), dtype=np.uint64)
    for s, c in enumerate(codes):
        v = int(c, 2) if c else 0
        for k in range(4):
            c4[s, k] = (v >> (64 * k)) & 0xFFFFFFFFFFFFFFFF
    return ln, c4


def test_sf_wide_codes_fixture(shafa, pkg):
    """Codes of 65..255 bits (sf_wide.cu) against the reference's own .shaf for a hand-made .cod (tests/golden/wide_codes)."""
    d = GOLDEN / "wide_codes"
    data = np.frombuffer((d / "x").read_bytes(), dtype=np.uint8)
    mode, sizes, ln, c4 = pkg.parse_cod_wide((d / "x.cod").read_bytes())
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    payload, poff, z = shafa.sf_encode_wide_host(data, offs, ln, c4)
    want = (d / "x.shaf").read_bytes()
    assert pkg.shaf_bytes(z, payload, poff) == want
    z2, payload2, poff2 = pkg.parse_shaf(want)
    assert_same(shafa.sf_decode_wide_host(payload2, poff2, z2, sizes, ln, c4), data, "wide decode of the reference's .shaf")


@pytest.mark.parametrize("seed,n,nblocks", [(1, 5_000, 1), (2, 200_000, 4), (3, 64, 2)])
def test_sf_wide_codes_random_tables(shafa, orc, pkg, seed, n, nblocks):
    rng = np.random.default_rng(seed)
    depths = sorted(set([0, 1, 2, 3] + list(rng.integers(4, 254, size=30))))
    cuts = np.linspace(0, n, nblocks + 1).astype(int)
    blocks, lens, c4s, want = [], [], [], []
    for b in range(nblocks):
        syms, codes = _wide_table(rng, len(depths) + 1, depths)
        p = np.array([2.0 ** -min(dp + 1, 10) for dp in depths] + [2.0 ** -10])
        blk = syms[rng.choice(len(p), size=cuts[b + 1] - cuts[b], p=p / p.sum())].astype(np.uint8)
        blk[: min(len(syms), blk.size)] = syms[: min(len(syms), blk.size)]
        ln, c4 = _codes_to_wide(codes)
        blocks.append(blk); lens.append(ln); c4s.append(c4)
        want.append(orc.sf_encode_block(blk, codes))
    data = np.concatenate(blocks)
    offs = np.array(cuts, dtype=np.uint64)
    payload, poff, z = shafa.sf_encode_wide_host(data, offs, np.stack(lens), np.stack(c4s))
    for b in range(nblocks):
        assert int(z[b]) == want[b].size
        assert_same(payload[int(poff[b]): int(poff[b]) + int(z[b])], want[b], f"wide payload of block {b}")
    sizes = np.diff(cuts).astype(np.uint64)
    assert_same(shafa.sf_decode_wide_host(payload, poff, z, sizes, np.stack(lens), np.stack(c4s)), data, "wide round trip")
    # a payload cut short, and a table that lacks a symbol of the data: errors, not garbage
    with pytest.raises(pkg.ShafaError) as e:
        shafa.sf_decode_wide_host(payload, poff, np.maximum(z // 2, 1), sizes, np.stack(lens), np.stack(c4s))
    assert e.value.code == 4
    ln_bad = np.stack(lens).copy()
    ln_bad[0, int(data[0])] = 0
    with pytest.raises(pkg.ShafaError) as e:
        shafa.sf_encode_wide_host(data, offs, ln_bad, np.stack(c4s))
    assert e.value.code == 4


def test_sf_decode_rejects_short_payload(shafa, orc, pkg):
    data = pkg.datagen.zipf(100_000, seed=2)
    codes = orc.sf_codes(orc.histogram(data))
    ln, cd, _ = orc.sf_codes_packed(orc.histogram(data))
    payload = orc.sf_encode_block(data, codes)
    cut = payload[: payload.size // 2].copy()
    with pytest.raises(pkg.ShafaError) as e:
        shafa.decode_tables(np.concatenate([cut, np.zeros(32, np.uint8)]), [0], [cut.size], [data.size], ln[None], cd[None], False, data.size + 64)
    assert e.value.code == 4


def _check_rle_decode(shafa, orc, blocks, lead=0):
    enc = [orc.rle_encode_block(b) for b in blocks]
    stream = np.concatenate([np.full(lead, 0xEE, np.uint8)] + enc)
    offs = np.cumsum([lead] + [e.size for e in enc]).astype(np.uint64)
    want = np.concatenate(blocks)
    d_out = shafa.alloc(want.size + 64)
    fin = shafa.rle_decode_blocks(shafa.to_device(stream), offs, d_out, want.size + 64)
    assert_same(fin, np.array([b.size for b in blocks], dtype=np.uint64), "final sizes")
    assert_same(d_out.download(np.uint8, want.size), want, "rle decode")


@pytest.mark.parametrize("gen,n,nblocks,lead", [("runheavy", 2_000_000, 3, 0), ("textlike", 1_000_000, 4, 4093), ("uniform", 100_000, 2, 1),
                                                ("mixed4m", 9_000_000, 3, 0), ("runheavy", 20_000_000, 2, 16)])
def test_rle_decode_generators(shafa, orc, pkg, gen, n, nblocks, lead):
    data = pkg.datagen.GENERATORS[gen](n, seed=11)
    cuts = np.linspace(0, n, nblocks + 1).astype(int)
    _check_rle_decode(shafa, orc, [data[cuts[i]:cuts[i + 1]] for i in range(nblocks)], lead)


def test_rle_decode_arbitrary_streams(shafa, orc, pkg):
    """Streams no encoder makes: count 0 (one copy, d.c:179-184), zero symbols, escapes back to back."""
    rng = np.random.default_rng(8)
    for rep in range(12):
        s = np.where(rng.random(20_000 + rep) < 0.3, 0, rng.integers(0, 4, size=20_000 + rep)).astype(np.uint8)
        try:
            want = orc.rle_decode_block(s)
        except orc.OracleError:
            with pytest.raises(pkg.ShafaError):
                shafa.rle_decode_blocks(shafa.to_device(np.concatenate([s, np.zeros(16, np.uint8)])), [0, s.size], shafa.alloc(6_000_000), 6_000_000)
            continue
        d_out = shafa.alloc(want.size + 64)
        fin = shafa.rle_decode_blocks(shafa.to_device(np.concatenate([s, np.zeros(16, np.uint8)])), [0, s.size], d_out, want.size + 64)
        assert int(fin[0]) == want.size
        assert_same(d_out.download(np.uint8, want.size), want, f"arbitrary stream {rep}")


def test_rle_decode_capacity_and_cap(shafa, orc, pkg):
    data = np.full(1_000_000, 9, dtype=np.uint8)
    enc = orc.rle_encode_block(data)
    d = shafa.to_device(np.concatenate([enc, np.zeros(16, np.uint8)]))
    with pytest.raises(pkg.ShafaError) as e:
        shafa.rle_decode_blocks(d, [0, enc.size], shafa.alloc(1000), 1000)
    assert e.value.code == 2
    # one block that expands past 64 MiB + 1 KiB is refused like the reference does (d.c:165-168)
    big = np.tile(np.array([0, 5, 255], dtype=np.uint8), 64 * 1024 * 1024 // 255 + 100)
    with pytest.raises(pkg.ShafaError) as e:
        shafa.rle_decode_blocks(shafa.to_device(np.concatenate([big, np.zeros(16, np.uint8)])), [0, big.size], shafa.alloc(70 * 1024 * 1024), 70 * 1024 * 1024)
    assert e.value.code == 4


# ------------------------------------------------------------------ whole pipeline through the host-buffer calls
def _check_pipeline(shafa, orc, data, block_size, fflags, artefact):
    force_rle, force_freq = "r" in fflags[1::2], "f" in fflags[1::2]
    enc = shafa.encode(data, block_size, force_rle, force_freq)
    base = ".rle" if enc.rle_mode else ""
    assert enc.rle_mode == (artefact(".rle") is not None)
    if enc.rle_mode:
        assert_same(enc.rle_bytes(), artefact(".rle"), ".rle")
    assert enc.freq_text() == artefact(base + ".freq"), base + ".freq"
    if enc.rle_mode and force_freq:
        assert enc.freq_orig_text() == artefact(".freq"), ".freq (-c f)"
    assert enc.cod_text() == artefact(base + ".cod"), base + ".cod"
    assert_same(enc.shaf_bytes(), artefact(base + ".shaf"), base + ".shaf")
    assert_same(shafa.decode(enc), data, "round trip")
    if enc.rle_mode:
        assert_same(shafa.decode(enc, rle_expand=False), artefact(".rle"), "-d s")
        offs = np.concatenate([[0], np.cumsum(enc.sym_sizes)]).astype(np.uint64)
        out, fin = shafa.rle_decode_host(enc.rle, offs, data.size + 64)
        assert_same(out, data, "-d r")
        assert_same(fin, enc.block_sizes, "-d r sizes")


@pytest.mark.parametrize("case", small_cases())
def test_pipeline_small_golden(shafa, orc, case):
    d = GOLDEN / case
    flags = json.loads((d / "flags.json").read_text())
    data = np.frombuffer((d / "input.bin").read_bytes(), dtype=np.uint8)

    def artefact(suffix):
        p = d / ("input.bin" + suffix)
        return p.read_bytes() if p.exists() else None

    _check_pipeline(shafa, orc, data, BLOCKS[flags["block"]], flags["fflags"], artefact)


def test_pipeline_big_golden(shafa, orc, pkg):
    """Larger inputs with K/m/M blocks: compared with the oracle (itself pinned to the
    reference's hashes for exactly these inputs in test_oracle_golden.py)."""
    big = json.loads((GOLDEN / "big_cases.json").read_text())
    for name, spec in big.items():
        data = pkg.datagen.GENERATORS[spec["generator"]](spec["n"], seed=spec["seed"])
        fflags = spec["fflags"]
        f = orc.module_f(data, BLOCKS[spec["block"]], force_rle="r" in fflags[1::2], force_freq="f" in fflags[1::2])
        sym = f["rle"] if f["rle_mode"] else data.tobytes()
        freq = f["rle_freq"] if f["rle_mode"] else f["freq"]
        cod = orc.module_t(freq)
        arte = {".rle": f["rle"] if f["rle_mode"] else None}
        base = ".rle" if f["rle_mode"] else ""
        arte[base + ".freq"] = freq
        arte[base + ".cod"] = cod
        arte[base + ".shaf"] = orc.module_c(sym, cod)
        if f["rle_mode"] and "f" in fflags[1::2]:
            arte[".freq"] = f["freq"]
        _check_pipeline(shafa, orc, data, BLOCKS[spec["block"]], fflags, lambda s, arte=arte: arte.get(s))


@pytest.mark.parametrize("case", encode_only_cases())
def test_one_symbol_block_encode(shafa, pkg, case):
    """One distinct symbol in a block: z = 0 and an empty payload, byte-identical to the reference's files;
    decoding such a block is refused (the symbol is not in the files; the reference decoder crashes, d.c:533-536)."""
    d = GOLDEN / "encode_only" / case
    flags = json.loads((d / "flags.json").read_text())
    data = np.frombuffer((d / "input.bin").read_bytes(), dtype=np.uint8)
    enc = shafa.encode(data, 65536, force_rle="r" in flags["fflags"][1::2])
    base = "input.bin.rle" if enc.rle_mode else "input.bin"
    assert enc.rle_mode == (d / "input.bin.rle").exists()
    if enc.rle_mode:
        assert_same(enc.rle_bytes(), (d / "input.bin.rle").read_bytes(), ".rle")
    assert enc.freq_text() == (d / (base + ".freq")).read_bytes()
    assert enc.cod_text() == (d / (base + ".cod")).read_bytes()
    assert_same(enc.shaf_bytes(), (d / (base + ".shaf")).read_bytes(), ".shaf")
    assert int(enc.z[-1]) == 0 and int(enc.code_len[-1].max()) == 0
    with pytest.raises(pkg.ShafaError) as e:
        shafa.decode(enc)
    assert e.value.code == 4


# ------------------------------------------------------------------ the fused single-pass decoder (opt-in: SHAFA_DECODE_FUSED=1)
@pytest.fixture
def fused_decoder(monkeypatch):
    monkeypatch.setenv("SHAFA_DECODE_FUSED", "1")      # read by launch_sf_decode at every call
    yield


@pytest.mark.parametrize("gen,n,bs", [("textlike", 30_000_000, 8388608), ("zipf", 9_000_000, 655360), ("runheavy", 40_000_000, 8388608),
                                      ("mixed4m", 26_000_000, 655360), ("uniform", 3_000_000, 655360), ("textlike", 70_001, 65536)])
def test_fused_decoder_round_trips(shafa, orc, pkg, fused_decoder, gen, n, bs):
    """k_dec_fused (speculative records + splice, chain warp, bulk async copies) against the oracle's symbol stream;
    blocks with codes of nearly one length (uniform) are routed to the exact kernels inside the same call."""
    data = pkg.datagen.GENERATORS[gen](n, seed=23)
    enc = shafa.encode(data, bs)
    before = shafa.profile_read() if False else None
    shafa.profile(True); shafa.profile_reset()
    sym = shafa.decode(enc, rle_expand=False)
    stats = shafa.profile_read()
    shafa.profile(False)
    f = orc.module_f(data, bs)
    assert_same(sym, f["rle"] if f["rle_mode"] else data.tobytes(), f"{gen}: fused -d s")
    assert_same(shafa.decode(enc), data, f"{gen}: fused -m d")
    if gen != "uniform":
        assert "dec_fused" in stats, stats.keys()
    _ = before


def test_fused_decoder_golden_payloads(shafa, pkg, fused_decoder):
    for case in small_cases():
        d = GOLDEN / case
        base = "input.bin.rle" if (d / "input.bin.rle").exists() else "input.bin"
        mode, sizes, ln, cd = pkg.parse_cod((d / (base + ".cod")).read_bytes())
        z, payload, offs = pkg.parse_shaf((d / (base + ".shaf")).read_bytes())
        sym, _ = shafa.decode_tables(payload, offs, z, sizes, ln, cd, False, int(sizes.sum()) + 64)
        assert_same(sym, (d / base).read_bytes(), f"{case}: fused -d s")


def test_fused_decoder_speculation_failure_falls_back(shafa, orc, pkg, fused_decoder):
    """A three-symbol alphabet with codes 0 / 10 / 11 and data that is all '10 10 10 ...' re-synchronises only by luck:
    whatever the fused kernel guesses is verified, and a wrong guess is decoded again by the exact kernels."""
    rng = np.random.default_rng(4)
    data = np.full(3_000_000, 66, dtype=np.uint8)
    data[rng.integers(0, data.size, size=40)] = 65           # a few 'A' (code 0) and 'C'
    data[rng.integers(0, data.size, size=40)] = 67
    data[:3] = [65, 66, 67]
    hist = orc.histogram(data)
    ln, cd, _ = orc.sf_codes_packed(hist)
    payload = orc.sf_encode_block(data, orc.sf_codes(hist))
    sym, _ = shafa.decode_tables(np.concatenate([payload, np.zeros(32, np.uint8)]), [0], [payload.size], [data.size], ln[None], cd[None], False,
                                 data.size + 64)
    assert_same(sym, data, "low-synchronisation code")
