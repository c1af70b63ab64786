"""ONE file sharded over N GPUs by one host process (shafa_cuda_*_multi): outputs must equal the
single-device outputs and the reference's fingerprints for any device list.  On a one-GPU box the same device is
listed N times (N contexts, the same sharding logic); with >= 2 devices visible the shards really run on two GPUs."""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np
import pytest

import test_gpu_full_configs as full

pytestmark = pytest.mark.gpu

MIB = 1 << 20


def device_lists(pkg):
    n = pkg.load_library().shafa_cuda_device_count()
    lists = [[0, 0], [0, 0, 0, 0], [0] * 8]            # 2 / 4 / 8 shards' worth of contexts on one GPU
    if n >= 2:
        lists.append(list(range(min(n, 8))))            # real devices
    return lists


@pytest.mark.parametrize("gen,n,bs,flags", [("mixed4m", 40 * 655360 + 77, 655360, {}), ("zipf", 200 * 65536 + 5, 65536, {}),
                                            ("textlike", 9 * 8388608 + 123, 8388608, {"force_freq": True}),
                                            ("uniform", 3 * 8388608, 8388608, {"force_rle": True})])
def test_multi_equals_single_device(shafa, pkg, gen, n, bs, flags):
    data = pkg.datagen.GENERATORS[gen](n, seed=17)
    whole = shafa.encode(data, bs, **flags)
    for devs in device_lists(pkg):
        with pkg.ShafaMulti(devs) as m:
            enc = m.encode(data, bs, **flags)
            assert enc.rle_mode == whole.rle_mode and enc.n_blocks == whole.n_blocks, devs
            assert enc.rle_bytes() == whole.rle_bytes(), devs
            assert enc.freq_text() == whole.freq_text() and enc.cod_text() == whole.cod_text(), devs
            assert enc.freq_orig_text() == whole.freq_orig_text(), devs
            assert enc.shaf_bytes() == whole.shaf_bytes(), devs
            assert np.array_equal(m.decode(enc), data), devs                                   # -m d
            if enc.rle_mode:
                assert np.array_equal(m.decode(enc, rle_expand=False), enc.rle), devs        # -d s
                offs = np.concatenate([[0], np.cumsum(enc.sym_sizes)]).astype(np.uint64)
                out, fin = m.rle_decode_host(enc.rle, offs, n + 64)                            # -d r
                assert np.array_equal(out, data) and np.array_equal(fin, enc.block_sizes), devs
            assert m.launch_count() > 0


def test_multi_errors_are_reported(pkg):
    with pytest.raises(pkg.ShafaError):
        pkg.ShafaMulti([999])
    with pkg.ShafaMulti([0, 0]) as m:
        with pytest.raises(pkg.ShafaError) as e:
            m.encode(np.zeros(100, np.uint8), 65536)                                           # f.c:220: file too small
        assert e.value.code == 6
        data = pkg.datagen.textlike(3 * MIB, seed=3)
        enc = m.encode(data, 65536)
        with pytest.raises(pkg.ShafaError) as e:
            m.decode_tables(enc.payload, enc.payload_offsets, enc.z, enc.sym_sizes, enc.code_len, enc.code_bits, True, 1000)
        assert e.value.code == 2


@pytest.mark.parametrize("name", ["C3_runs_4g_M", "C2_text_1g_m"])
def test_multi_full_config_vs_reference(pkg, name):
    """Full-size configs sharded 8 ways (real devices when there are 8, else 8 contexts on what there is):
    artefact hashes against the reference's, then the three decode routes."""
    case = full.full_cases()[name]
    data = pkg.datagen.GENERATORS[case["generator"]](case["n"], seed=case["seed"])
    ndev = pkg.load_library().shafa_cuda_device_count()
    devs = [i % ndev for i in range(8)]
    with pkg.ShafaMulti(devs) as m:
        c, v = full.encode_and_check(m, pkg, name + f" on devices {devs}", case, data)
        try:
            full.decode_all(m, name, v, data)
        finally:
            m.encoded_free(c)
