#!/usr/bin/env python3
"""Design tool (CPU only): statistics of self-synchronising sub-sequence decoding on the bench's data shapes.

For a sample of a BASELINE config's data it builds the reference's Shannon-Fano code (oracle), encodes one block and
measures, per 128-bit sub-sequence:
  * lookups of a greedy multi-code walk with a W-bit window (how many table lookups one full pass costs),
  * how many LOOKUPS a walk from the true start needs until it lands on a code boundary of the speculative walk from
    bit 0 ("merge"), i.e. what the second pass costs when the first pass keeps its code boundaries as a bit mask,
  * how many sub-sequences do not merge inside their own 128 bits (they need a further round).
usage: tools/sim_sync.py [textlike|zipf|uniform|runheavy] [MiB] [window bits]
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import importlib

datagen = importlib.import_module("shafa-cd_b200.datagen")
from oracle import oracle_py as orc


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "textlike"
    mib = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    W = int(sys.argv[3]) if len(sys.argv) > 3 else 13
    n = int(mib * 1024 * 1024)
    data = datagen.GENERATORS[kind](n)
    sym = orc.rle_encode_block(data) if kind in ("textlike", "runheavy") else data
    hist = orc.histogram(sym)
    codes = orc.sf_codes(hist)
    payload = orc.sf_encode_block(sym, codes)
    lens = np.array([len(c) for c in codes])
    print(f"{kind}: {sym.size} symbols, {payload.size} bytes, {8 * payload.size / sym.size:.2f} bits/symbol, code lengths {lens[lens > 0].min()}..{lens.max()}")
    bits = np.unpackbits(payload)
    nb = bits.size
    maxlen = int(lens.max())
    # code length matched at every bit position (0 = no code): table over maxlen-bit windows
    T = max(maxlen, W)
    tab = np.zeros(1 << T, dtype=np.uint8)
    for s, c in enumerate(codes):
        if c:
            v = int(c, 2) << (T - len(c))
            tab[v:v + (1 << (T - len(c)))] = len(c)
    padded = np.concatenate([bits, np.zeros(T + 64, dtype=np.uint8)]).astype(np.uint32)
    win = np.zeros(nb + 64, dtype=np.uint32)
    for k in range(T):
        win = (win << 1) | padded[k:k + nb + 64]
    clen = tab[win].astype(np.int64)  # length of the code starting at each bit
    clen[clen == 0] = 1
    # greedy W-bit window step from every position: take whole codes while they fit
    step = np.zeros(nb + 64, dtype=np.int64)
    pos = np.arange(nb + 64)
    cur = pos.copy()
    ncodes = np.zeros(nb + 64, dtype=np.int64)
    for _ in range(W):
        l = clen[np.minimum(cur, nb + 63)]
        ok = (cur + l - pos) <= W
        ok &= (cur - pos) == step  # still chained
        step = np.where(ok, step + l, step)
        ncodes += ok
        cur = np.where(ok, cur + l, cur)
    step[step == 0] = clen[step == 0]  # code longer than the window: one code
    print(f"window {W}: {step[:nb].mean():.2f} bits and {ncodes[:nb].mean():.2f} codes per lookup at a random position")
    # the true path
    true = np.zeros(nb + 200, dtype=bool)
    q = 0
    while q < nb:
        true[q] = True
        q += int(clen[q])
    SUB = 128
    nsub = nb // SUB - 1
    look_full, look_merge, nomerge, look_true = [], [], 0, []
    merged_codes = []
    for i in range(1, nsub):
        lo = i * SUB
        # speculative walk from bit 0 of the sub-sequence with full windows while q < 128: code boundaries
        q, nl = lo, 0
        spec = set()
        while q < lo + SUB:
            # all code boundaries inside this lookup
            e = q + int(step[q])
            c = q
            while c < e:
                spec.add(c)
                c += int(clen[c])
            q = e
            nl += 1
        spec.add(q)
        look_full.append(nl)
        # true start: first true code boundary >= lo
        s = lo
        while not true[s]:
            s += 1
        q, nl = s, 0
        while q not in spec and q < lo + SUB:
            q += int(step[q])
            nl += 1
        if q >= lo + SUB and q not in spec:
            nomerge += 1
        look_merge.append(nl)
        # code-level merge distance for comparison
        q, nc = s, 0
        while q not in spec and q < lo + SUB:
            q += int(clen[q])
            nc += 1
        merged_codes.append(nc)
        q, nl = s, 0
        while q < lo + SUB:
            q += int(step[q])
            nl += 1
        look_true.append(nl)
    look_full, look_merge, look_true, merged_codes = map(np.array, (look_full, look_merge, look_true, merged_codes))
    print(f"lookups per full pass from bit 0: mean {look_full.mean():.2f} max {look_full.max()}  (warp max of 32: {np.mean([look_full[k:k+32].max() for k in range(0, len(look_full) - 32, 32)]):.2f})")
    print(f"lookups per full pass from the true start: mean {look_true.mean():.2f}")
    print(f"lookups until merge (lookup walk onto old code boundaries): mean {look_merge.mean():.2f}  p90 {np.percentile(look_merge, 90):.0f} p99 {np.percentile(look_merge, 99):.0f} max {look_merge.max()}  (warp max of 32: {np.mean([look_merge[k:k+32].max() for k in range(0, len(look_merge) - 32, 32)]):.2f})")
    print(f"codes until merge (single-code walk): mean {merged_codes.mean():.2f} p99 {np.percentile(merged_codes, 99):.0f}")
    print(f"sub-sequences that do not merge inside their 128 bits: {nomerge} of {len(look_full)} = {100.0 * nomerge / len(look_full):.3f}%  ({256.0 * nomerge / len(look_full):.2f} per 4 KiB sequence)")


if __name__ == "__main__":
    main()
