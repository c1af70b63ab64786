"""Block sharding across GPUs.

Blocks are independent in every module (f.c:38, c.c:80, d.c:569), so a file is
split into contiguous block ranges, one per GPU, and the per-shard results are
concatenated in block order.  The only file-wide value is block 0's R/N decision
(f.c:246-258): its owner decides and everybody else is told -- a one-byte
control-plane message, not a data-path collective.
"""
from __future__ import annotations

import numpy as np


def block_range(n_blocks: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced block range [first, first+count) of `rank` (earlier ranks take the remainder)."""
    base, extra = divmod(n_blocks, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def shard_bytes(n: int, block_size: int, world: int, rank: int) -> tuple[int, int]:
    """Byte range [lo, hi) of the file that `rank` owns."""
    n_blocks = (n + block_size - 1) // block_size
    first, count = block_range(n_blocks, world, rank)
    return min(n, first * block_size), min(n, (first + count) * block_size)


class MergedEncoded:
    """Per-shard results concatenated in block order; same text/container methods as `Encoded`."""

    def __init__(self, parts, block_size: int):
        from . import binding

        self._b = binding
        parts = [p for p in parts if p is not None and p.n_blocks]
        self.rle_mode = parts[0].rle_mode
        assert all(p.rle_mode == self.rle_mode for p in parts), "shards disagree on the R/N mode"
        self.block_size = block_size
        self.n = sum(p.n for p in parts)
        self.n_blocks = sum(p.n_blocks for p in parts)
        cat = np.concatenate
        self.block_sizes = cat([p.block_sizes for p in parts])
        self.sym_sizes = cat([p.sym_sizes for p in parts])
        self.hist = cat([p.hist for p in parts])
        self.hist_orig = cat([p.hist_orig for p in parts]) if parts[0].hist_orig is not None else None
        self.code_len = cat([p.code_len for p in parts])
        self.code_bits = cat([p.code_bits for p in parts])
        self.z = cat([p.z for p in parts])
        self.rle = cat([p.rle for p in parts])
        offs, payloads, base = [], [], 0
        for p in parts:
            offs.append(p.payload_offsets + np.uint64(base))
            pad = (-len(p.payload)) % 16
            payloads.append(np.concatenate([p.payload, np.zeros(pad, np.uint8)]))
            base += len(p.payload) + pad
        self.payload_offsets = cat(offs)
        self.payload = cat(payloads)

    @property
    def mode(self) -> str:
        return "R" if self.rle_mode else "N"

    def freq_text(self) -> bytes:
        return self._b.freq_text(self.mode, self.sym_sizes, self.hist)

    def freq_orig_text(self):
        return None if self.hist_orig is None else self._b.freq_text("N", self.block_sizes, self.hist_orig)

    def cod_text(self) -> bytes:
        return self._b.cod_text(self.mode, self.sym_sizes, self.code_len, self.code_bits)

    def shaf_bytes(self) -> bytes:
        return self._b.shaf_bytes(self.z, self.payload, self.payload_offsets)

    def rle_bytes(self) -> bytes:
        return self.rle.tobytes()


def encode_sharded(shafa, data: np.ndarray, block_size: int, world: int, rank: int, force_rle=False, force_freq=False,
                   bcast=None):
    """Encode this rank's block range on its GPU.  `bcast(value_or_None)` must return rank 0's value on every
    rank (torch.distributed broadcast of one byte); with bcast=None the caller passes mode through `rle_mode`."""
    lo, hi = shard_bytes(data.size, block_size, world, rank)
    part = None
    mode = None
    if rank == 0:
        part = shafa.encode(data[lo:hi], block_size, force_rle, force_freq)
        mode = part.rle_mode
    if bcast is not None:
        mode = bcast(mode)
    if rank != 0 and hi > lo:
        part = shafa.encode_shard(data[lo:hi], block_size, bool(mode), force_freq)
    return part, bool(mode)
