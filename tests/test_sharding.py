"""Multi-GPU path: block sharding, the one file-wide value (block 0's R/N decision) and ordered merge.
CPU: world_size-2 gloo processes with the oracle as the per-shard engine.  GPU: shards run one after the
other on one device and must merge to exactly the whole-file result."""
from __future__ import annotations

import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def test_block_range_partitions(pkg):
    br = pkg.sharding.block_range
    for n_blocks in (1, 2, 7, 64, 103, 128, 6554):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                first, count = br(n_blocks, world, r)
                seen.extend(range(first, first + count))
            assert seen == list(range(n_blocks))
            counts = [br(n_blocks, world, r)[1] for r in range(world)]
            assert max(counts) - min(counts) <= 1


def _gloo_worker(rank: int, world: int, port: int, out_dir: str):
    sys.path.insert(0, str(ROOT))
    import torch
    import torch.distributed as dist
    from __graft_entry__ import load_package
    from oracle import oracle_py as orc

    pkg = load_package()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    bs = 65536
    data = pkg.datagen.mixed4m(9 * 65536 + 1234, seed=3)          # 10 blocks, R mode decided by block 0
    lo, hi = pkg.sharding.shard_bytes(data.size, bs, world, rank)
    # block 0's owner decides R/N (f.c:246-258) and tells the others: one byte, no data-path collective
    mode = torch.zeros(1, dtype=torch.uint8)
    if rank == 0:
        rle0 = orc.rle_encode_block(data[:bs]).size
        mode[0] = int(orc.rle_gate(bs, rle0))
    dist.broadcast(mode, src=0)
    rle_mode = bool(mode.item())
    # per-shard engine on the CPU = the oracle, block by block
    sym_blocks, hists = [], []
    for b0 in range(lo, hi, bs):
        blk = data[b0:min(hi, b0 + bs)]
        sym = orc.rle_encode_block(blk) if rle_mode else blk
        sym_blocks.append(sym)
        hists.append(orc.histogram(sym))
    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, hi, [s.tobytes() for s in sym_blocks], [h.tolist() for h in hists]))
    if rank == 0:
        gathered.sort(key=lambda g: g[0])
        assert gathered[0][0] == 0 and gathered[-1][1] == data.size
        assert all(a[1] == b[0] for a, b in zip(gathered, gathered[1:]))
        syms = [s for g in gathered for s in g[2]]
        hist = np.array([h for g in gathered for h in g[3]], dtype=np.uint64)
        sizes = [len(s) for s in syms]
        f = orc.module_f(data, bs)
        assert f["rle_mode"] == rle_mode
        assert b"".join(syms) == (f["rle"] if rle_mode else data.tobytes())
        assert pkg.freq_text("R" if rle_mode else "N", sizes, hist) == (f["rle_freq"] if rle_mode else f["freq"])
        Path(out_dir, "ok").write_text("ok")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_shards(tmp_path):
    import torch.multiprocessing as mp

    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").exists()


@pytest.mark.gpu
@pytest.mark.parametrize("gen,n,bs,world", [("mixed4m", 20 * 655360 + 77, 655360, 4), ("zipf", 7 * 65536 + 5, 65536, 2),
                                            ("textlike", 5 * 8388608 + 123, 8388608, 8)])
def test_shards_merge_to_whole_file(shafa, pkg, gen, n, bs, world):
    data = pkg.datagen.GENERATORS[gen](n, seed=13)
    whole = shafa.encode(data, bs)
    mode = None
    parts = []
    for rank in range(world):
        part, mode = pkg.sharding.encode_sharded(shafa, data, bs, world, rank, bcast=(lambda m, mode=mode: m if m is not None else mode))
        parts.append(part)
    merged = pkg.sharding.MergedEncoded(parts, bs)
    assert merged.rle_mode == whole.rle_mode and merged.n_blocks == whole.n_blocks
    assert merged.rle_bytes() == whole.rle_bytes()
    assert merged.freq_text() == whole.freq_text()
    assert merged.cod_text() == whole.cod_text()
    assert merged.shaf_bytes() == whole.shaf_bytes()
    assert np.array_equal(shafa.decode(merged), data)
