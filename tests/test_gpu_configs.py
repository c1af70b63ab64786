"""BASELINE.json configurations at (or near) their full sizes: size fingerprints recorded from the
reference during the survey plus size-independent properties
(encode -> decode round trip, per-block size sums, histogram totals)."""
from __future__ import annotations

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

MIB = 1 << 20


def _properties(shafa, data, enc, block_size):
    B = (data.size + block_size - 1) // block_size
    assert enc.n_blocks == B and int(enc.block_sizes.sum()) == data.size
    assert int(enc.hist.sum()) == int(enc.sym_sizes.sum())                      # every symbol counted once
    assert np.array_equal(enc.hist.sum(axis=1), enc.sym_sizes)                  # per block
    bits = (enc.hist * enc.code_len.astype(np.uint64)).sum(axis=1)
    assert np.array_equal(enc.z, (bits + 7) // 8)                               # z_b = ceil(sum(len)/8) (c.c:80)
    assert ((enc.hist > 0) == (enc.code_len > 0)).all()                         # a code for exactly the present symbols
    kraft = np.where(enc.code_len > 0, 2.0 ** -enc.code_len.astype(np.float64), 0).sum(axis=1)
    assert np.allclose(kraft, 1.0)                                              # full prefix tree per block
    out = shafa.decode(enc)
    assert out.size == data.size and np.array_equal(out, data)


def test_c1_zipf_64mib_K(shafa, pkg):
    data = pkg.datagen.zipf(64 * MIB, seed=12345)
    enc = shafa.encode(data, pkg.datagen.BLOCK_K)
    assert not enc.rle_mode and enc.n_blocks == 103
    assert len(enc.freq_text()) == 114_438 and len(enc.cod_text()) == 287_509 and len(enc.shaf_bytes()) == 44_613_363
    assert int(enc.code_len.max()) == 12
    _properties(shafa, data, enc, pkg.datagen.BLOCK_K)


def test_c3_runheavy_256mib_M(shafa, pkg):
    data = pkg.datagen.runheavy(256 * MIB, seed=99)
    enc = shafa.encode(data, pkg.datagen.BLOCK_M)
    assert enc.rle_mode and enc.n_blocks == 4
    assert enc.rle.size == 6_205_660 and len(enc.shaf_bytes()) == 3_641_256
    _properties(shafa, data, enc, pkg.datagen.BLOCK_M)
    offs = np.concatenate([[0], np.cumsum(enc.sym_sizes)]).astype(np.uint64)
    out, fin = shafa.rle_decode_host(enc.rle, offs, data.size + 64)             # -d r
    assert np.array_equal(fin, enc.block_sizes) and np.array_equal(out, data)


def test_c2_textlike_256mib_m(shafa, pkg):
    data = pkg.datagen.textlike(256 * MIB, seed=2024)
    enc = shafa.encode(data, pkg.datagen.BLOCK_m)
    assert enc.rle_mode and enc.n_blocks == 32
    _properties(shafa, data, enc, pkg.datagen.BLOCK_m)
    assert np.array_equal(shafa.decode(enc, rle_expand=False), enc.rle)         # -d s


def test_c4_uniform_64mib_K_force_freq(shafa, pkg):
    data = pkg.datagen.uniform(64 * MIB, seed=7)
    enc = shafa.encode(data, pkg.datagen.BLOCK_K, force_freq=True)
    assert not enc.rle_mode and enc.n_blocks == 103
    assert 7 <= int(enc.code_len[enc.code_len > 0].min()) and int(enc.code_len.max()) <= 9
    assert 1.0 < enc.z.sum() / data.size < 1.002                                # incompressible: ~0.09 % expansion
    _properties(shafa, data, enc, pkg.datagen.BLOCK_K)


def test_c5_mixed_decode_only(shafa, pkg):
    data = pkg.datagen.mixed(192 * MIB, seed=5)                                  # run-heavy / zipf / text segments of 64 MiB
    enc = shafa.encode(data, pkg.datagen.BLOCK_m)
    assert enc.rle_mode
    sym = shafa.decode(enc, rle_expand=False)                                    # -d s
    assert np.array_equal(sym, enc.rle)
    offs = np.concatenate([[0], np.cumsum(enc.sym_sizes)]).astype(np.uint64)
    out, _ = shafa.rle_decode_host(sym, offs, data.size + 64)                    # -d r
    assert np.array_equal(out, data)
    assert np.array_equal(shafa.decode(enc), data)                               # -m d


def test_pipelined_host_calls_match_serial(shafa, pkg):
    """With caller-provided output buffers the host-buffer calls run chunked on worker streams
    (csrc/host_pipeline.cu); the result must be identical to the serial path."""
    for gen, n, bs in (("textlike", 200 * MIB + 12345, pkg.datagen.BLOCK_m), ("zipf", 96 * MIB, pkg.datagen.BLOCK_K)):
        data = pkg.datagen.GENERATORS[gen](n, seed=21)
        serial = shafa.encode(data, bs)                                          # library-owned buffers: serial path
        rle_buf = np.empty(2 * n + 4096, dtype=np.uint8)
        pay_buf = np.empty(n + n // 4 + 65536, dtype=np.uint8)
        c = shafa.encode_raw(data, bs, rle_buf=rle_buf, payload_buf=pay_buf)     # pipelined path
        try:
            piped = pkg.Encoded(c)
        finally:
            shafa.encoded_free(c)
        assert piped.rle_mode == serial.rle_mode and piped.n_blocks == serial.n_blocks
        assert piped.rle_bytes() == serial.rle_bytes()
        assert piped.freq_text() == serial.freq_text() and piped.cod_text() == serial.cod_text()
        assert piped.shaf_bytes() == serial.shaf_bytes()
        assert np.array_equal(shafa.decode(piped), data)                         # decode_host is pipelined for >= 8 blocks
        if piped.rle_mode:
            assert np.array_equal(shafa.decode(piped, rle_expand=False), piped.rle)
        # too small an output buffer: the needed size is still reported (error 2), nothing is written past the end
        with pytest.raises(pkg.ShafaError) as e:
            shafa.decode_tables(piped.payload, piped.payload_offsets, piped.z, piped.sym_sizes, piped.code_len, piped.code_bits,
                                piped.rle_mode, 1 << 20)
        assert e.value.code == 2
