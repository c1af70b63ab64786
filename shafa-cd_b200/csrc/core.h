// core.h -- per-thread arithmetic of the Shafa kernels, as host+device inline
// functions.  The CUDA kernels call these from their tiles; tests/cpu/emu.cpp
// calls the very same functions from a sequential tile emulator, so the rules
// (RLE emission, RLE parse automaton, Shannon-Fano split, bit packing, prefix
// decode) are checked against the oracle on the CPU before any GPU time is spent.
//
// Reference rules restated here (paths under /root/reference/src/modules):
//   RLE encode   f.c:29-55      RLE decode  d.c:144-186
//   SF split     t.c:130-152    bit packing c.c:52-83     prefix decode d.c:514-551
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define SHAFA_HD __host__ __device__ __forceinline__
#else
#define SHAFA_HD inline
#endif

namespace shafa {

SHAFA_HD int ctz32(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return __ffs((int)x) - 1;
#else
    return __builtin_ctz(x);
#endif
}
SHAFA_HD int clz32(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return __clz((int)x);
#else
    return x ? __builtin_clz(x) : 32;
#endif
}
SHAFA_HD uint32_t bswap32(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return __byte_perm(x, 0, 0x0123);
#else
    return __builtin_bswap32(x);
#endif
}

// upper 32 bits of (hi:lo) << (sh & 31)
SHAFA_HD uint32_t funnel_l(uint32_t lo, uint32_t hi, uint32_t sh)
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(lo, hi, sh);
#else
    sh &= 31u;
    return sh ? ((hi << sh) | (lo >> (32u - sh))) : hi;
#endif
}

// =============================================================================
// RLE encode (f.c:29-55) as a per-position rule.
//
// Let k be the offset of position p inside its maximal equal-byte run (runs are
// cut at block edges), c = k mod 255 and rem the run length remaining from p
// (capped at 255).  The greedy loop of the reference emits, at p:
//   c == 0            : {0,sym,rem} if sym == 0 or rem >= 4, else the literal sym
//   1 <= c <= 3       : the literal sym if sym != 0 and c + rem < 4
//   otherwise         : nothing
// "Heads" are the first positions of runs; `heads` is a bitmask over tile
// positions (bit p of word p>>5) that also has every position at or past the
// block end set, and must be readable up to bit p+255 for any queried p.
// =============================================================================

constexpr int RLE_ITEMS = 16;           // bytes per 16-byte group (one vector load)
constexpr int RLE_VECS = 4;             // groups per thread of the encoder: 64 contiguous bytes
constexpr int RLE_THREADS = 256;
constexpr int RLE_TILE = RLE_THREADS * RLE_ITEMS * RLE_VECS;  // 16 KiB per tile
constexpr int RLE_HALO = 288;           // look-ahead positions kept after the tile
constexpr int RLE_HEAD_WORDS = (RLE_TILE + RLE_HALO) / 32;

// distance from p to the next head strictly after p, capped at 255 (= min(rem,255))
SHAFA_HD uint32_t rle_run_remaining(const uint32_t *heads, uint32_t p)
{
    uint32_t q = p + 1;
    const uint32_t end = p + 256;  // inspect positions p+1 .. p+255
    while (q < end) {
        const uint32_t w = heads[q >> 5] >> (q & 31);
        if (w) {
            const uint32_t d = q + (uint32_t)ctz32(w) - p;
            return d < 255 ? d : 255;
        }
        q = (q | 31u) + 1;
    }
    return 255;
}

struct RleEmit {
    uint32_t lit;     // bit j: position j emits a literal
    uint32_t tri;     // bit j: position j emits a triple
    uint32_t cnt[4];  // count byte of the triple at position j = (cnt[j>>2] >> 8*(j&3)) & 255
    uint32_t bytes;   // bytes emitted by this thread
};

// Decide what the RLE_ITEMS positions starting at tile position p0 emit.
// `sym(j)` = byte at p0+j; nvalid = number of positions of this thread inside
// the block (0..16); c0 = (k mod 255) at p0, used when p0 is not a head.
template <typename SymFn>
SHAFA_HD RleEmit rle_emit_thread(const uint32_t *heads, uint32_t p0, int nvalid, uint32_t c0, SymFn sym)
{
    RleEmit e;
    e.lit = e.tri = e.bytes = 0;
    e.cnt[0] = e.cnt[1] = e.cnt[2] = e.cnt[3] = 0;
    uint32_t c = c0;
#pragma unroll
    for (int j = 0; j < RLE_ITEMS; j++) {
        if (j < nvalid) {
            const uint32_t p = p0 + (uint32_t)j;
            if ((heads[p >> 5] >> (p & 31)) & 1u) c = 0;
            if (c <= 3) {
                const uint32_t s = sym(j);
                const uint32_t rem = rle_run_remaining(heads, p);
                if (c == 0) {
                    if (s == 0 || rem >= 4) {
                        e.tri |= 1u << j;
                        e.cnt[j >> 2] |= rem << (8 * (j & 3));
                        e.bytes += 3;
                    } else {
                        e.lit |= 1u << j;
                        e.bytes += 1;
                    }
                } else if (s != 0 && c + rem < 4) {
                    e.lit |= 1u << j;
                    e.bytes += 1;
                }
            }
            c = (c == 254) ? 0 : c + 1;
        }
    }
    return e;
}

// Bit-parallel form of the same rule for the common case k < 255 at all 16
// positions of a thread (k = offset inside the run).  E[p] = 1 iff byte p equals
// byte p-1 ("continuation"; 0 at the block start and at/after the block end).
//   ewin   bit i = E[p0 - 8 + i]   (needs to be exact for positions p0-2 .. p0+18)
//   z16    bit j = byte p0+j is 0x00
//   valid  low nvalid bits set
// A run has length >= 4 iff it has >= 3 consecutive continuation bits:
//   T  = E & E>>1 & E>>2           first of three consecutive continuations
//   Lc = T | T<<1 | T<<2           every continuation position of a long run
//   LH = ~E & Lc>>1                heads of long runs
// triple at heads of long runs and at heads of zero runs; literal at every
// position of a short non-zero run; nothing elsewhere.
struct RleMasks {
    uint32_t lit, tri;
};
SHAFA_HD RleMasks rle_fast_masks(uint32_t ewin, uint32_t z16, uint32_t valid16)
{
    const uint32_t T = ewin & (ewin >> 1) & (ewin >> 2);
    const uint32_t Lc = T | (T << 1) | (T << 2);
    const uint32_t H = ~ewin;
    const uint32_t LH = H & (Lc >> 1);
    const uint32_t lc16 = (Lc >> 8) & 0xffffu, lh16 = (LH >> 8) & 0xffffu, h16 = (H >> 8) & 0xffffu;
    RleMasks m;
    m.tri = h16 & (lh16 | z16) & valid16;
    m.lit = ~z16 & ~lc16 & ~lh16 & valid16 & 0xffffu;
    return m;
}

// The same rule on 64 positions at once (one thread of the 16 KiB-tile kernel).
//   E      bit p = E[p0 + p]
//   before bit 0 = E[p0 - 2], bit 1 = E[p0 - 1]
//   after  bit 0 = E[p0 + 64], bit 1 = E[p0 + 65], bit 2 = E[p0 + 66]
struct RleMasks64 {
    uint64_t lit, tri;
};
SHAFA_HD RleMasks64 rle_fast_masks64(uint64_t E, uint32_t before, uint32_t after, uint64_t Z, uint64_t valid)
{
    const uint64_t a0 = after & 1u, a1 = (after >> 1) & 1u, a2 = (after >> 2) & 1u;
    const uint64_t m2 = before & 1u, m1 = (before >> 1) & 1u;
    const uint64_t E1 = (E >> 1) | (a0 << 63);               // E[p+1]
    const uint64_t E2 = (E >> 2) | (a0 << 62) | (a1 << 63);  // E[p+2]
    const uint64_t T = E & E1 & E2;                          // first of three consecutive continuations
    const uint64_t Tm1 = m1 & E & (E >> 1) & 1ull;           // T[-1]
    const uint64_t Tm2 = m2 & m1 & E & 1ull;                 // T[-2]
    const uint64_t T64 = a0 & a1 & a2;                       // T[64]
    const uint64_t Lc = T | (T << 1) | Tm1 | (T << 2) | (Tm1 << 1) | Tm2;
    const uint64_t Lc64 = T64 | (T >> 63) | ((T >> 62) & 1ull);
    const uint64_t H = ~E;
    const uint64_t LH = H & ((Lc >> 1) | (Lc64 << 63));
    RleMasks64 m;
    m.tri = (LH | (H & Z)) & valid;
    m.lit = ~Z & ~Lc & ~LH & valid;
    return m;
}

// Write this thread's emitted bytes to dst (dst points at the thread's first
// output byte).  Returns bytes written (== e.bytes).
template <typename SymFn>
SHAFA_HD uint32_t rle_write_thread(const RleEmit &e, uint8_t *dst, SymFn sym)
{
    uint32_t o = 0;
#pragma unroll
    for (int j = 0; j < RLE_ITEMS; j++) {
        if ((e.tri >> j) & 1u) {
            dst[o] = 0;
            dst[o + 1] = (uint8_t)sym(j);
            dst[o + 2] = (uint8_t)((e.cnt[j >> 2] >> (8 * (j & 3))) & 255u);
            o += 3;
        } else if ((e.lit >> j) & 1u) {
            dst[o++] = (uint8_t)sym(j);
        }
    }
    return o;
}

// =============================================================================
// RLE decode (d.c:144-186) as a 3-state automaton over the RLE bytes:
//   state 0 (start)      : byte 0x00 -> state 1, else literal (emit 1) -> state 0
//   state 1 (expect sym) : any byte -> state 2
//   state 2 (expect cnt) : emit cnt copies of the previous byte (1 if cnt == 0) -> state 0
// A thread summarises its bytes [j0, j1) as a map entry-state -> exit-state,
// packed 2 bits per entry state; maps compose associatively.
// =============================================================================

constexpr uint32_t FSM_IDENTITY = 0u | (1u << 2) | (2u << 4);

SHAFA_HD uint32_t fsm_apply(uint32_t map, uint32_t state) { return (map >> (2 * state)) & 3u; }
// first a, then b
SHAFA_HD uint32_t fsm_compose(uint32_t a, uint32_t b)
{
    return fsm_apply(b, fsm_apply(a, 0)) | (fsm_apply(b, fsm_apply(a, 1)) << 2) | (fsm_apply(b, fsm_apply(a, 2)) << 4);
}
SHAFA_HD bool fsm_is_const(uint32_t m) { return fsm_apply(m, 0) == fsm_apply(m, 1) && fsm_apply(m, 1) == fsm_apply(m, 2); }

// zmask bit q = (valid byte q == 0); the thread's valid bytes are [0, n), n <= 16
// (callers shift the mask so that the first valid byte is bit 0).
// ex[q] = exit state when the automaton is in state 0 at position q; past the
// end an overshoot of 0/1/2 positions means state 0 / 2 (count pending) / 1 (sym
// pending).  ex[q] = zero(q) ? ex[q+3] : ex[q+1], evaluated right to left in
// three rolling registers so that it unrolls without local memory.
SHAFA_HD uint32_t fsm_thread_map(uint32_t zmask, int n)
{
    uint32_t a = 0, b = 2, c = 1;  // ex[q+1], ex[q+2], ex[q+3] for q = n-1
#pragma unroll
    for (int q = RLE_ITEMS - 1; q >= 0; q--) {
        if (q < n) {
            const uint32_t v = ((zmask >> q) & 1u) ? c : a;
            c = b;
            b = a;
            a = v;
        }
    }
    // entry 0: state 0 at position 0 -> ex[0]; entry 1: sym at 0, count at 1 -> ex[2];
    // entry 2: count at 0 -> ex[1]
    return a | (c << 2) | (b << 4);
}

// Shortcut used by the kernel: >= 2 valid bytes and none of them zero means every
// entry state is back in state 0 within two bytes and stays there (constant map 0).
SHAFA_HD uint32_t fsm_thread_map_fast(uint32_t zmask, int n)
{
    if (zmask == 0 && n >= 2) return 0u;
    return fsm_thread_map(zmask, n);
}

// 64-byte form used by the 16 KiB-tile decoder.  Only the escape bytes matter: from a
// position in state 0 every non-zero byte is a literal and the next zero byte is an escape
// that swallows the two bytes after it.  Z bit q = (byte q == 0), bits >= n clear.
SHAFA_HD int ctz64(uint64_t x)
{
#if defined(__CUDA_ARCH__)
    return __ffsll((long long)x) - 1;
#else
    return __builtin_ctzll(x);
#endif
}

// Runs longer than 255: the reference counts to 255, emits, and starts over (f.c:38-49; the counter c of
// rle_emit_thread above), so the first position of every further 255-byte chunk behaves exactly like the
// head of a new run -- triple if the rest of the run has >= 4 bytes or the byte is 0x00, literals otherwise.
// Given a thread whose first position continues a run (E bit 0 set) at distance k0 from that run's head,
// clear the continuation bit of such chunk heads inside the window the bit-parallel rule looks at
// (positions p0-2 .. p0+66).  At most one falls into it (255 > 69).
SHAFA_HD void rle_chunk_heads(uint64_t k0, uint64_t *E, uint32_t *before, uint32_t *after)
{
    const uint64_t e = *E;
    const uint32_t lead = (~e) ? (uint32_t)ctz64(~e) : 64u;  // positions 0 .. lead-1 continue the incoming run
    const uint32_t r = (uint32_t)(k0 % 255u);
    const uint32_t j = r ? 255u - r : 0u;  // first position >= p0 at a multiple of 255 from the head
    if (j < lead) {
        *E = e & ~(1ull << j);
    } else if (lead == 64 && j < 67) {  // p0+64 .. p0+66 continue the run only through set bits
        const uint32_t i = j - 64, through = (1u << (i + 1)) - 1u;
        if ((*after & through) == through) *after &= ~(1u << i);
    }
    if (k0 > 1 && (k0 - 1) % 255u == 0) *before &= ~2u;  // p0-1
    if (k0 > 2 && (k0 - 2) % 255u == 0) *before &= ~1u;  // p0-2
}

struct FsmWalk {
    uint64_t esc;   // positions of the escape bytes met in state 0
    uint32_t exit;  // state after the n bytes
};
SHAFA_HD FsmWalk fsm_walk64(uint64_t Z, int n, uint32_t entry)
{
    FsmWalk w;
    w.esc = 0;
    const int cur = entry == 0 ? 0 : (entry == 1 ? 2 : 1);  // first position in state 0
    if (cur > n) {  // the pending triple is not even finished inside these n bytes
        w.exit = (entry == 1 && n == 0) ? 1u : 2u;
        return w;
    }
    w.exit = 0;
    uint64_t zz = (cur >= 64) ? 0ull : ((Z >> cur) << cur);
    while (zz) {
        const int zp = ctz64(zz);
        w.esc |= 1ull << zp;
        const int nxt = zp + 3;
        if (nxt > n) {
            w.exit = (nxt - n == 2) ? 1u : 2u;  // escape was the last byte: sym pending; second to last: count pending
            break;
        }
        zz = (nxt >= 64) ? 0ull : ((zz >> nxt) << nxt);
    }
    return w;
}
SHAFA_HD uint32_t fsm_map64(uint64_t Z, int n)
{
    if (Z == 0 && n >= 2) return 0u;  // no escape at all: every entry state is back in state 0 and stays there
    // The walks from entry states 1 and 2 start in state 0 at positions 2 and 1.  If the walk from entry state 0 is in
    // state 0 there as well (no escape at positions 0/1 resp. 0 swallowed them) the walks coincide from then on: one
    // walk instead of three for all but the threads whose first two bytes hold a zero.
    const FsmWalk w0 = fsm_walk64(Z, n, 0);
    const uint32_t e1 = (n >= 3 && (w0.esc & 3ull) == 0) ? w0.exit : fsm_walk64(Z, n, 1).exit;
    const uint32_t e2 = (n >= 3 && (w0.esc & 1ull) == 0) ? w0.exit : fsm_walk64(Z, n, 2).exit;
    return w0.exit | (e1 << 2) | (e2 << 4);
}
// positions that emit: lit = literal bytes, run = count bytes of triples (the run's byte is the one before)
struct RleDecMasks {
    uint64_t lit, run;
};
SHAFA_HD RleDecMasks rle_dec_masks64(uint64_t esc, int n, uint32_t entry)
{
    const uint64_t maskn = (n >= 64) ? ~0ull : ((1ull << n) - 1ull);
    const uint64_t pre = entry == 1 ? 3ull : (entry == 2 ? 1ull : 0ull);      // bytes owed to a triple opened earlier
    const uint64_t prerun = entry == 1 ? 2ull : (entry == 2 ? 1ull : 0ull);   // where that triple's count byte sits
    RleDecMasks m;
    m.run = ((esc << 2) | prerun) & maskn;
    m.lit = ~(esc | (esc << 1) | (esc << 2) | pre) & maskn;
    return m;
}

// Walk the thread's n valid bytes from entry state `state`; `byte(q)` returns
// byte q, prev = the byte just before byte 0 (needed when entering in state 2).
// emit(len, sym) is called once per record, in order.  Returns the exit state.
template <typename ByteFn, typename EmitFn>
SHAFA_HD uint32_t rle_parse_thread(uint32_t state, int n, uint32_t prev, ByteFn byte, EmitFn emit)
{
    uint32_t s = state, last = prev;
#pragma unroll
    for (int q = 0; q < RLE_ITEMS; q++) {
        if (q < n) {
            const uint32_t v = byte(q);
            if (s == 0) {
                if (v == 0)
                    s = 1;
                else
                    emit(1u, v);
            } else if (s == 1) {
                s = 2;
            } else {
                emit(v ? v : 1u, last);  // d.c:179-184: count 0 -> one copy
                s = 0;
            }
            last = v;
        }
    }
    return s;
}

// =============================================================================
// Shannon-Fano split (t.c:130-152): among d in [lo, hi-1] the SMALLEST d that
// minimises |2*(P[d+1]-P[lo]) - (P[hi+1]-P[lo])|.  P is the exclusive prefix of
// the rank-sorted counts (P[0] = 0, P[r+1] = sum of ranks 0..r); every count in
// [lo, hi] is > 0, so the sequence is strictly V-shaped and a binary search for
// the crossing point plus one comparison reproduces the reference's
// first-local-minimum scan (strict '<' keeps the smaller d on ties).
// =============================================================================
SHAFA_HD int sf_split(const uint32_t *P, int lo, int hi)
{
    const uint64_t base = P[lo];
    const uint64_t T = (uint64_t)P[hi + 1] - base;
    int a = lo, c = hi;
    while (a < c) {  // smallest d with 2*left(d) >= T
        const int m = (a + c) >> 1;
        if (2 * ((uint64_t)P[m + 1] - base) >= T)
            c = m;
        else
            a = m + 1;
    }
    const int dstar = a;
    if (dstar == lo) return lo;
    const uint64_t dif1 = T - 2 * ((uint64_t)P[dstar] - base);  // at dstar-1, positive
    if (dstar == hi) return hi - 1;
    const uint64_t dif2 = 2 * ((uint64_t)P[dstar + 1] - base) - T;  // at dstar, >= 0
    return (dif1 <= dif2) ? dstar - 1 : dstar;
}

// =============================================================================
// Bit packing (c.c:52-83): codes are concatenated MSB-first.  Output is handled
// as big-endian 32-bit words (bit 31 of word w = stream bit 32*w) that are
// byte-swapped when stored.
// =============================================================================

constexpr int ENC_SHORT_MAX = 26;                          // fast path: code + 5-bit length in one word
constexpr int ENC_THREADS = 256;
constexpr int ENC_VECS = 4;                                // 16-byte vectors per thread, contiguous
constexpr int ENC_ITEMS = 16 * ENC_VECS;                   // 64 symbols per thread
constexpr int ENC_TILE = ENC_THREADS * ENC_ITEMS;          // 16 KiB of symbols per tile
#ifndef SHAFA_ENC_STAGE_WORDS
#define SHAFA_ENC_STAGE_WORDS 5120
#endif
constexpr int ENC_STAGE_WORDS = SHAFA_ENC_STAGE_WORDS;      // staging window: 20 KiB of the tile's bit range (10 bits per symbol)
// Encoder LUT entry of a code of at most ENC_SHORT_MAX (< 28) bits: length in the top five bits -- so that
// "bits += length" is one shift-add -- and the right-aligned code below.
SHAFA_HD uint32_t enc_entry(uint64_t code, uint32_t len) { return (len << 27) | (uint32_t)code; }
SHAFA_HD uint32_t enc_len(uint32_t entry) { return entry >> 27; }
SHAFA_HD uint32_t enc_code(uint32_t entry) { return entry & 0x07FFFFFFu; }

// 64-bit window starting at bit `q` of a big-endian word array
SHAFA_HD uint64_t bit_window(const uint32_t *words, uint32_t q)
{
    const uint32_t i = q >> 5, sh = q & 31u;
    const uint64_t hi = ((uint64_t)words[i] << 32) | words[i + 1];
    return sh ? ((hi << sh) | (uint64_t)(words[i + 2] >> (32 - sh))) : hi;
}

// =============================================================================
// Prefix decode (d.c:514-551) with a 12-bit first-level table and a sorted
// fallback for longer codes.  A Shannon-Fano code (any prefix-free code) sorted
// by its left-aligned bits has the property that the code matching a window is
// the LARGEST left-aligned code <= window.
// =============================================================================
constexpr int DEC_LUT_BITS = 12;
constexpr int DEC_LUT_SIZE = 1 << DEC_LUT_BITS;
constexpr int DEC_LUT2_BITS = 13;       // window of the count-only table: one byte per entry, so 8 KiB as well
constexpr int DEC_LUT2_SIZE = 1 << DEC_LUT2_BITS;
constexpr uint32_t DEC_LONG = 0xFF;   // lut entry low byte: 1..12 length, 0 invalid, 0xFF -> fallback

struct DecTables {
    const uint16_t *lut;     // [4096]  (sym << 8) | len
    const uint64_t *sleft;   // [256] left-aligned codes, ascending (absent = ~0)
    const uint8_t *slen;     // [256]
    const uint8_t *ssym;     // [256]
};

// returns code length (>= 1) and *sym, or 0 if the window matches no code
SHAFA_HD uint32_t dec_symbol(const DecTables &t, uint64_t window, uint32_t *sym)
{
    const uint32_t e = t.lut[window >> (64 - DEC_LUT_BITS)];
    const uint32_t l = e & 0xFFu;
    if (l - 1u < (uint32_t)DEC_LUT_BITS) {
        *sym = e >> 8;
        return l;
    }
    if (l != DEC_LONG) return 0;
    int lo = 0, hi = 256;  // count of entries <= window
    while (lo < hi) {
        const int m = (lo + hi) >> 1;
        if (t.sleft[m] <= window)
            lo = m + 1;
        else
            hi = m;
    }
    if (lo == 0) return 0;
    const int r = lo - 1;
    const uint32_t len = t.slen[r];
    if (len == 0) return 0;
    if ((window >> (64 - len)) != (t.sleft[r] >> (64 - len))) return 0;
    *sym = t.ssym[r];
    return len;
}

constexpr int DEC_SUB_BITS = 128;                      // sub-sequence (one thread)
constexpr int DEC_THREADS = 256;                       // sub-sequences per sequence (one CTA)
constexpr int DEC_SEQ_BYTES = DEC_SUB_BITS * DEC_THREADS / 8;  // 4096
constexpr int DEC_SEQ_WORDS = DEC_SEQ_BYTES / 4;       // 1024

struct DecSub {
    uint32_t end;    // bit position (relative to the sub-sequence) where decoding stopped; may pass DEC_SUB_BITS
    uint32_t count;  // symbols that START inside [start, stop)
    uint32_t bad;    // a window matched no code
};

// ---- sequence-level decoding -------------------------------------------------------------------
// Bit positions are relative to the sub-sequence.  A window is only ever read at a position < 128 and
// the longest code has 64 bits, so a sub-sequence needs the six big-endian words w[0..5] that start at
// its first bit.  They are kept as a ROW of five overlapping pairs, row[k] = (w[k] << 32) | w[k+1]:
// the 32 bits starting at any bit q are one 64-bit load and a funnel shift, with no refill branch to
// diverge a warp and no address arithmetic beyond row + (q >> 5).  Rows are 40 bytes apart, so the 32
// lanes of a warp reading their own rows do not collide on shared-memory banks more than chance.
constexpr int DEC_ROW = 5;                               // pairs per row
constexpr int DEC_ROW_WORDS = DEC_ROW + 1;               // words per row
constexpr int DEC_PAIR_SLOTS = DEC_ROW * DEC_THREADS;    // 1280 slots of 8 bytes per sequence

SHAFA_HD void dec_row_from_words(const uint32_t *w, uint64_t *row)
{
    for (int k = 0; k < DEC_ROW; k++) row[k] = ((uint64_t)w[k] << 32) | (uint64_t)w[k + 1];
}

// V is a view of the tables (DecView below, or the kernels' shared-memory view with the same accessors);
// `row` is the slot of the sub-sequence's first pair (sub-sequence index * DEC_ROW)
template <typename V>
SHAFA_HD uint32_t dec_window32(const V &t, uint32_t row, uint32_t q)
{
    const uint64_t pr = t.pair(row + (q >> 5));
    return (uint32_t)((pr << (q & 31u)) >> 32);
}
template <typename V>
SHAFA_HD uint64_t dec_window64(const V &t, uint32_t row, uint32_t q)
{
    const uint32_t i = q >> 5, sh = q & 31u;
    const uint64_t hi = t.pair(row + i);
    return sh ? ((hi << sh) | ((t.pair(row + i + 1) & 0xffffffffull) >> (32 - sh))) : hi;  // low word of pair i+1 is w[i+2]
}

// Tables of one block as seen by a decoding thread (shared memory on the GPU).
//   lut1[4096]  (sym << 8) | len, len 1..12; 0 = no code; 0xFF = longer than 12 bits
//   lut2[8192]  (nsym << 4) | bits, one byte: how many whole codes fit in the 13-bit window and their total length; 0 = use lut1
//   lut3[4096]  (bits << 24) | (nsym << 16) | (sym1 << 8) | sym0: the first one or two whole codes of the window; 0 = use lut1
struct DecView {
    const uint64_t *pairs;
    const uint16_t *lut1;
    const uint8_t *lut2;
    const uint32_t *lut3;
    const uint8_t *lut16;
    const uint64_t *sleft;
    const uint8_t *slen;
    const uint8_t *ssym;
    SHAFA_HD uint64_t pair(uint32_t slot) const { return pairs[slot]; }
    SHAFA_HD uint32_t l1(uint32_t i) const { return lut1[i]; }
    SHAFA_HD uint32_t l1len(uint32_t i) const { return lut1[i] & 0xFFu; }
    SHAFA_HD uint32_t l16(uint32_t i) const { return lut16[i]; }
    SHAFA_HD uint32_t l2(uint32_t i) const { return lut2[i]; }
    SHAFA_HD uint32_t l3(uint32_t i) const { return lut3[i]; }
    SHAFA_HD bool has_l3() const { return lut3 != nullptr; }
};

// one symbol at bit q: returns its length (>= 1), *sym; *bad is set if the bits are no code
template <typename V>
SHAFA_HD uint32_t dec_one(const V &t, uint32_t row, uint32_t q, uint32_t *sym, uint32_t *bad)
{
    const uint32_t e = t.l1(dec_window32(t, row, q) >> (32 - DEC_LUT_BITS));
    uint32_t l = e & 0xFFu;
    *sym = e >> 8;
    if (l - 1u >= (uint32_t)DEC_LUT_BITS) {  // long code or no code: exact 64-bit window + sorted search
        DecTables dt;
        dt.lut = t.lut1;
        dt.sleft = t.sleft;
        dt.slen = t.slen;
        dt.ssym = t.ssym;
        l = dec_symbol(dt, dec_window64(t, row, q), sym);
        if (l == 0) {
            *bad = 1;
            l = 1;
        }
    }
    return l;
}

// the same from lut3 alone (first code of the entry) with the sorted search for long codes: no lut1 needed
template <typename V>
SHAFA_HD uint32_t dec_one3(const V &t, uint32_t row, uint32_t q, uint32_t *sym, uint32_t *bad)
{
    const uint32_t e3 = t.l3(dec_window32(t, row, q) >> (32 - DEC_LUT_BITS));
    if (e3) {
        *sym = e3 & 0xFFu;
        return (e3 >> 20) & 0xFu;
    }
    // long code or no code: exact 64-bit window + sorted search (largest left-aligned code <= window)
    const uint64_t window = dec_window64(t, row, q);
    int lo = 0, hi = 256;
    while (lo < hi) {
        const int m = (lo + hi) >> 1;
        if (t.sleft[m] <= window)
            lo = m + 1;
        else
            hi = m;
    }
    if (lo > 0) {
        const int r = lo - 1;
        const uint32_t len = t.slen[r];
        if (len != 0 && (window >> (64 - len)) == (t.sleft[r] >> (64 - len))) {
            *sym = t.ssym[r];
            return len;
        }
    }
    *bad = 1;
    *sym = 0;
    return 1;
}

// Count the symbols that START in [start, stop) of the sub-sequence whose row is `row` (stop <= 128)
// and find where decoding leaves the range.  While a whole 12-bit window lies before `stop`, all codes
// inside it are taken at once (lut2).
template <typename V>
SHAFA_HD DecSub dec_count_subsequence(const V &t, uint32_t row, uint32_t start, uint32_t stop)
{
    DecSub r;
    r.count = 0;
    r.bad = 0;
    int32_t q = (int32_t)start;
    const int32_t stop13 = (int32_t)stop - DEC_LUT2_BITS;
    for (;;) {
        while (q <= stop13) {
            const uint32_t e2 = t.l2(dec_window32(t, row, (uint32_t)q) >> (32 - DEC_LUT2_BITS));
            if (e2 == 0) break;
            q += (int32_t)(e2 & 0xFu);
            r.count += e2 >> 4;
        }
        if (q >= (int32_t)stop) break;
        uint32_t sym;  // a code longer than the window, or the last few bits of the range: one code
        q += (int32_t)dec_one(t, row, (uint32_t)q, &sym, &r.bad);
        r.count++;
    }
    r.end = (uint32_t)q;
    return r;
}

// Decode the symbols that start in [start, stop).  While a whole 12-bit window lies before `stop`, the
// first two codes inside it are taken at once (lut3).  The sink keeps the running symbol index itself:
// em.one(sym) appends one symbol, em.two(e3) the one or two symbols of a lut3 entry (e3 & 0xFF, then
// (e3 >> 8) & 0xFF if e3 & 0x20000; byte 2 of e3 is how many), em.count() is how many were appended.
template <typename V, typename Sink>
SHAFA_HD DecSub dec_emit_to(const V &t, uint32_t row, uint32_t start, uint32_t stop, Sink &em)
{
    DecSub r;
    r.bad = 0;
    int32_t q = (int32_t)start;
    const int32_t stop12 = t.has_l3() ? (int32_t)stop - DEC_LUT_BITS : -1;
    for (;;) {
        while (q <= stop12) {
            const uint32_t e3 = t.l3(dec_window32(t, row, (uint32_t)q) >> (32 - DEC_LUT_BITS));
            if (e3 == 0) break;
            em.two(e3);
            q += (int32_t)(e3 >> 24);
        }
        if (q >= (int32_t)stop) break;
        uint32_t sym;
        q += (int32_t)(t.has_l3() ? dec_one3(t, row, (uint32_t)q, &sym, &r.bad) : dec_one(t, row, (uint32_t)q, &sym, &r.bad));
        em.one(sym);
    }
    r.end = (uint32_t)q;
    r.count = em.count();
    return r;
}

// the same with a plain callable emit(i, sym)
template <typename F>
struct DecFnSink {
    F f;
    uint32_t n;
    SHAFA_HD void one(uint32_t sym) { f(n++, sym); }
    SHAFA_HD void two(uint32_t e3)
    {
        f(n++, e3 & 0xFFu);
        if (e3 & 0x20000u) f(n++, (e3 >> 8) & 0xFFu);
    }
    SHAFA_HD uint32_t count() const { return n; }
};
template <typename V, typename Emit>
SHAFA_HD DecSub dec_emit_subsequence(const V &t, uint32_t row, uint32_t start, uint32_t stop, Emit emit)
{
    DecFnSink<Emit> sink{emit, 0u};
    return dec_emit_to(t, row, start, stop, sink);
}

struct DecNoEmit {
    SHAFA_HD void operator()(uint32_t, uint32_t) const {}
};

// ---- the same walks with the sub-sequence's six big-endian words IN REGISTERS ---------------------------
// The word index is a compile-time loop counter, so a window is one funnel shift of two registers: no
// shared-memory load (and no address arithmetic) per lookup for the payload, only for the table.  The
// kernels use these for a thread's OWN sub-sequence; the row versions above remain for the rare rounds in
// which a thread re-decodes somebody else's.
template <typename V>
SHAFA_HD uint32_t dec_long_words(const V &t, const uint32_t (&w)[DEC_ROW_WORDS], int i, uint32_t q, uint32_t *sym, uint32_t *bad)
{
    const uint32_t sh = q & 31u;
    const uint64_t hi64 = ((uint64_t)w[i] << 32) | (uint64_t)w[i + 1];
    const uint64_t window = sh ? ((hi64 << sh) | ((uint64_t)w[i + 2] >> (32u - sh))) : hi64;
    int lo = 0, hi = 256;  // count of entries <= window
    while (lo < hi) {
        const int m = (lo + hi) >> 1;
        if (t.sleft[m] <= window)
            lo = m + 1;
        else
            hi = m;
    }
    if (lo > 0) {
        const int r = lo - 1;
        const uint32_t len = t.slen[r];
        if (len != 0 && (window >> (64 - len)) == (t.sleft[r] >> (64 - len))) {
            *sym = t.ssym[r];
            return len;
        }
    }
    *bad = 1;
    *sym = 0;
    return 1;
}

// dec_count_subsequence on registers (needs lut1 and lut2)
template <typename V>
SHAFA_HD DecSub dec_count_words(const V &t, const uint32_t (&w)[DEC_ROW_WORDS], uint32_t start, uint32_t stop)
{
    DecSub r;
    r.count = 0;
    r.bad = 0;
    uint32_t q = start;
    const int32_t stop13 = (int32_t)stop - DEC_LUT2_BITS;
#pragma unroll
    for (int i = 0; i < DEC_SUB_BITS / 32; i++) {
        const uint32_t lim = stop < 32u * (uint32_t)(i + 1) ? stop : 32u * (uint32_t)(i + 1);
        while (q < lim) {
            const uint32_t win = funnel_l(w[i + 1], w[i], q);
            if ((int32_t)q <= stop13) {
                const uint32_t e2 = t.l2(win >> (32 - DEC_LUT2_BITS));
                if (e2) {  // all whole codes of the 13-bit window at once
                    q += e2 & 0xFu;
                    r.count += e2 >> 4;
                    continue;
                }
            }
            const uint32_t e = t.l1(win >> (32 - DEC_LUT_BITS));
            uint32_t l = e & 0xFFu;
            if (l - 1u >= (uint32_t)DEC_LUT_BITS) {
                uint32_t sym;
                l = dec_long_words(t, w, i, q, &sym, &r.bad);
            }
            q += l;
            r.count++;
        }
    }
    r.end = q;
    return r;
}

// dec_emit_to on registers (needs lut3 only: its entries carry the first code's length)
template <typename V, typename Sink>
SHAFA_HD DecSub dec_emit_words(const V &t, const uint32_t (&w)[DEC_ROW_WORDS], uint32_t start, uint32_t stop, Sink &em)
{
    DecSub r;
    r.bad = 0;
    uint32_t q = start;
    const int32_t stop12 = (int32_t)stop - DEC_LUT_BITS;
#pragma unroll
    for (int i = 0; i < DEC_SUB_BITS / 32; i++) {
        const uint32_t lim = stop < 32u * (uint32_t)(i + 1) ? stop : 32u * (uint32_t)(i + 1);
        while (q < lim) {
            const uint32_t e3 = t.l3(funnel_l(w[i + 1], w[i], q) >> (32 - DEC_LUT_BITS));
            if (e3) {
                if ((int32_t)q <= stop12) {  // both codes of the entry start before `stop`
                    em.two(e3);
                    q += e3 >> 24;
                } else {
                    em.one(e3 & 0xFFu);
                    q += (e3 >> 20) & 0xFu;
                }
            } else {
                uint32_t sym;
                q += dec_long_words(t, w, i, q, &sym, &r.bad);
                em.one(sym);
            }
        }
    }
    r.end = q;
    r.count = em.count();
    return r;
}

// lut2 entry for the 13-bit window value e, derived from lut1 (codes of up to 12 bits): how many whole codes
// the window holds and their total length
SHAFA_HD uint8_t dec_lut2_entry(const uint16_t *lut1, uint32_t e)
{
    uint32_t pos = 0, n = 0;
    while (pos < (uint32_t)DEC_LUT2_BITS) {
        // the next 12 bits of the window, zero padded
        const uint32_t l = lut1[((e << pos) >> (DEC_LUT2_BITS - DEC_LUT_BITS)) & (DEC_LUT_SIZE - 1)] & 0xFFu;
        if (l - 1u >= (uint32_t)DEC_LUT_BITS || pos + l > (uint32_t)DEC_LUT2_BITS) break;  // must fit entirely
        pos += l;
        n++;
    }
    return (uint8_t)(n ? ((n << 4) | pos) : 0);  // n, pos <= 13
}

// ---- codes of nearly one length ("lean" blocks: complete code, no code longer than the lut1 window, 2 * shortest >
// longest) ----------------------------------------------------------------------------------------------------------
// Such codes never re-synchronise, so k_dec_sync_lean walks every sub-sequence from every start that can occur.  The
// walks are bound by shared-memory table gathers (random 12-bit indices: 2.8 bank-conflict wavefronts per warp load,
// LSU pipe 95 % busy on C4), so the remedy is fewer loads: lut16 maps a 16-bit window to ALL whole codes inside it
// (two 8-bit codes on uniform bytes), one byte per entry: (codes << 5) | bits, at most seven codes, bits <= 16.
constexpr int DEC_LUT16_BITS = 16;
constexpr int DEC_LUT16_SIZE = 1 << DEC_LUT16_BITS;
SHAFA_HD uint8_t dec_lut16_entry(const uint16_t *lut1, uint32_t v)
{
    uint32_t pos = 0, n = 0;
    while (n < 7u) {
        const uint32_t l = lut1[(((v << pos) & 0xFFFFu) >> (DEC_LUT16_BITS - DEC_LUT_BITS))] & 0xFFu;  // the 12 bits from pos, zero padded
        if (l - 1u >= (uint32_t)DEC_LUT_BITS || pos + l > (uint32_t)DEC_LUT16_BITS) break;              // must fit entirely
        pos += l;
        n++;
    }
    return n ? (uint8_t)((n << 5) | pos) : (uint8_t)((1u << 5) | 1u);  // (n = 0 cannot happen for a lean block; one code of one bit still advances)
}

// (count, end) of the chains of codes from the NS consecutive starts start0, start0 + 1, ... while they are inside
// [.., stop): several codes per lookup while a whole 16-bit window lies before `stop`, one code per lookup (length = low
// byte of lut1) for the last few bits.  The NS chains are independent, so their table loads are in flight together:
// the kernel runs two CTAs of eight warps per SM (the table is 64 KiB) and would otherwise wait for one load at a time.
template <int NS, typename V>
SHAFA_HD void dec_lean_walk(const V &t, const uint32_t (&w)[DEC_ROW_WORDS], uint32_t start0, uint32_t stop, uint32_t (&count)[NS], uint32_t (&end)[NS])
{
    uint32_t q[NS], cnt[NS];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < NS; k++) {
        q[k] = start0 + (uint32_t)k;
        cnt[k] = 0;
    }
    const uint32_t multi_end = stop >= (uint32_t)DEC_LUT16_BITS ? stop - (uint32_t)DEC_LUT16_BITS + 1u : 0u;  // q < multi_end: q + 16 <= stop
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < DEC_SUB_BITS / 32; i++) {  // the word of q is a compile-time index: a window is one funnel shift of two registers
        const uint32_t lim = stop < 32u * (uint32_t)(i + 1) ? stop : 32u * (uint32_t)(i + 1);
        const uint32_t lim16 = multi_end < lim ? multi_end : lim;
        // (the exit tests come BEFORE the loads: a trailing iteration that only finds every chain done would cost as many
        // table loads as a useful one -- measured: 2.2 x the instructions and 3 x the shared-memory wavefronts)
        auto any_below = [&](uint32_t bound) {
            bool any = false;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int k = 0; k < NS; k++) any |= q[k] < bound;
            return any;
        };
        while (any_below(lim16)) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int k = 0; k < NS; k++) {  // branch-free per chain: a chain that has left the word just does not move
                const bool a = q[k] < lim16;
                const uint32_t e = t.l16(funnel_l(w[i + 1], w[i], q[k]) >> (32 - DEC_LUT16_BITS));
                q[k] += a ? (e & 31u) : 0u;
                cnt[k] += a ? (e >> 5) : 0u;
            }
        }
        while (any_below(lim)) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int k = 0; k < NS; k++) {
                const bool a = q[k] < lim;
                const uint32_t l = t.l1len(funnel_l(w[i + 1], w[i], q[k]) >> (32 - DEC_LUT_BITS));
                q[k] += a ? l : 0u;
                cnt[k] += a ? 1u : 0u;
            }
        }
    }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < NS; k++) {
        count[k] = cnt[k];
        end[k] = q[k];
    }
}

// lut3 entry for the 12-bit window value e: its first one or two whole codes.
//   [7:0] sym0  [15:8] sym1  [17:16] how many (1|2)  [23:20] bits of the first code  [31:24] bits of the codes taken
// The first code's own length makes the table usable for single-code steps too, so the write kernel does not need
// lut1 in shared memory (8 KiB: one more resident CTA per SM).
SHAFA_HD uint32_t dec_lut3_entry(const uint16_t *lut1, uint32_t e)
{
    const uint32_t e0 = lut1[e];
    const uint32_t l0 = e0 & 0xFFu;
    if (l0 - 1u >= (uint32_t)DEC_LUT_BITS) return 0;
    const uint32_t e1 = lut1[(e << l0) & (DEC_LUT_SIZE - 1)];
    const uint32_t l1 = e1 & 0xFFu;
    if (l1 - 1u < (uint32_t)DEC_LUT_BITS && l0 + l1 <= (uint32_t)DEC_LUT_BITS)
        return ((l0 + l1) << 24) | (l0 << 20) | (2u << 16) | ((e1 >> 8) << 8) | (e0 >> 8);
    return (l0 << 24) | (l0 << 20) | (1u << 16) | (e0 >> 8);
}


// =============================================================================
// Fused single-pass decoder (sf_decode_fused.cu): decode ONCE, speculatively, and keep the symbols.
//
// A thread decodes its 128-bit sub-sequence from bit 0 before it knows where the first code really
// starts, and keeps what it decoded as a list of RECORDS (one per table lookup: one or two symbols and
// the bit position they start at).  When the true start s arrives (the predecessor's end) the thread
// walks single codes from s only until it lands on a position where a record starts: from there on
// the speculative path IS the true path (decoding from a given bit is deterministic), so the final
// symbols are a short HEAD (the codes walked from s) followed by the tail of the record list.  A
// prefix code re-synchronises within a few codes, so every payload bit is decoded about 1.2 times
// instead of once per synchronisation round plus once more for the output.
//
//   pair-LUT entry (12-bit window -> its first one or two whole codes):
//     [7:0] sym0  [15:8] sym1  [22:16] bits of the codes taken  [27:24] bits of the first code  [29:28] how many (1|2)
//     0 = the first code is longer than 12 bits (or the window is no code): sorted search
//   record = entry with byte 2 replaced by the start position (< 128) inside the sub-sequence
// =============================================================================
constexpr int DECF_CAP = 24;    // records per sub-sequence (more than ~48 symbols in 128 bits overflow into the direct path)
constexpr int DECF_HEAD = 16;   // head symbols kept per sub-sequence before the list is rebuilt from the new start

SHAFA_HD uint32_t decf_total(uint32_t e) { return (e >> 16) & 0x7Fu; }
SHAFA_HD uint32_t decf_len0(uint32_t e) { return (e >> 24) & 0xFu; }
SHAFA_HD uint32_t decf_n(uint32_t e) { return e >> 28; }
SHAFA_HD uint32_t decf_pos(uint32_t rec) { return (rec >> 16) & 0xFFu; }

// pair-LUT entry for the 12-bit window value e, from lut1 ((sym << 8) | len, 0xFF = long, 0 = none)
SHAFA_HD uint32_t decf_entry(const uint16_t *lut1, uint32_t e)
{
    const uint32_t e0 = lut1[e];
    const uint32_t l0 = e0 & 0xFFu;
    if (l0 - 1u >= (uint32_t)DEC_LUT_BITS) return 0;
    const uint32_t e1 = lut1[(e << l0) & (DEC_LUT_SIZE - 1)];
    const uint32_t l1 = e1 & 0xFFu;
    if (l1 - 1u < (uint32_t)DEC_LUT_BITS && l0 + l1 <= (uint32_t)DEC_LUT_BITS)
        return (e0 >> 8) | ((e1 >> 8) << 8) | ((l0 + l1) << 16) | (l0 << 24) | (2u << 28);
    return (e0 >> 8) | (l0 << 16) | (l0 << 24) | (1u << 28);
}

// code at the 64-bit window by sorted search (codes longer than the LUT window): length, or 0 if no code
template <typename V>
SHAFA_HD uint32_t decf_long(const V &t, uint64_t window, uint32_t *sym)
{
    int lo = 0, hi = 256;  // count of entries <= window
    while (lo < hi) {
        const int m = (lo + hi) >> 1;
        if (t.sleft[m] <= window)
            lo = m + 1;
        else
            hi = m;
    }
    if (lo == 0) return 0;
    const int r = lo - 1;
    const uint32_t len = t.slen[r];
    if (len == 0 || (window >> (64 - len)) != (t.sleft[r] >> (64 - len))) return 0;
    *sym = t.ssym[r];
    return len;
}

// entry (1 or 2 symbols) for the code(s) at bit q; a long or invalid code comes back as a one-symbol entry
template <typename V>
SHAFA_HD uint32_t decf_lookup(const V &t, uint32_t row, uint32_t q, uint32_t *bad)
{
    uint32_t e = t.lp(dec_window32(t, row, q) >> (32 - DEC_LUT_BITS));
    if (e == 0) {
        uint32_t sym = 0;
        uint32_t l = decf_long(t, dec_window64(t, row, q), &sym);
        if (l == 0) {
            *bad = 1;
            l = 1;
        }
        e = sym | (l << 16) | (1u << 28);
    }
    return e;
}

struct DecfMain {      // the record list of a sub-sequence
    uint32_t m;        // records stored
    uint32_t k;        // symbols in them (all symbols of the walk if norec)
    uint32_t end;      // bit position (relative to the sub-sequence) where the walk left [.., stop)
    uint32_t norec;    // the list overflowed DECF_CAP: k and end are right, the records are not usable
    uint32_t bad;
};

// Walk the sub-sequence from `start`, two codes per lookup where a 12-bit window holds them, and store one
// record per lookup at slot + 4 j.  Symbols are those that START in [start, stop).
template <typename V>
SHAFA_HD DecfMain decf_pass(const V &t, uint32_t row, uint32_t slot, uint32_t start, uint32_t stop)
{
    DecfMain r;
    r.norec = 0;
    r.bad = 0;
    uint32_t q = start, k = 0, j = 0;
    while (q < stop && j < (uint32_t)DECF_CAP) {
        const uint32_t e = decf_lookup(t, row, q, &r.bad);
        t.st32(slot + 4u * j, (e & 0xFF00FFFFu) | (q << 16));
        j++;
        q += decf_total(e);
        k += decf_n(e);
    }
    if (q >= stop) {
        if (j) {  // the last lookup may have taken a second code that starts at or after `stop`: it belongs to the next range
            const uint32_t rec = t.ld32(slot + 4u * (j - 1));
            const uint32_t q0 = decf_pos(rec);
            if (decf_n(rec) == 2u && q0 + decf_len0(rec) >= stop) {
                t.st32(slot + 4u * (j - 1), (rec & 0x0FFFFFFFu) | (1u << 28));
                k--;
                q = q0 + decf_len0(rec);
            }
        }
    } else {  // out of record space (codes of 1-2 bits): count on, one code at a time
        r.norec = 1;
        while (q < stop) {
            const uint32_t e = decf_lookup(t, row, q, &r.bad);
            q += decf_n(e) == 2u ? decf_len0(e) : decf_total(e);
            k++;
        }
    }
    r.m = j;
    r.k = k;
    r.end = q;
    return r;
}

struct DecfSplice {
    uint32_t kind;   // 0 = met the record list at record jsp, 1 = left the range without meeting it, 2 = head full (rebuild the list)
    uint32_t hn;     // head symbols written
    uint32_t jsp;    // first record that belongs to the true path
    uint32_t kskip;  // symbols in the records before it
    uint32_t end;    // kind 1: where the walk left the range
    uint32_t bad;
};

// The true start `s` is known: walk single codes from it until a record of the list starts at the same bit.
template <typename V>
SHAFA_HD DecfSplice decf_splice(const V &t, uint32_t row, uint32_t slot, uint32_t head, const DecfMain &mn, uint32_t s, uint32_t stop)
{
    DecfSplice r;
    r.hn = r.jsp = r.kskip = r.end = r.bad = 0;
    const uint32_t m = mn.norec ? 0u : mn.m;
    uint32_t q = s, j = 0, kskip = 0, hn = 0;
    for (;;) {
        if (q >= stop) {
            r.kind = 1;
            r.end = q;
            break;
        }
        uint32_t p = 0xFFFFFFFFu;
        while (j < m) {
            const uint32_t rec = t.ld32(slot + 4u * j);
            p = decf_pos(rec);
            if (p >= q) break;
            kskip += decf_n(rec);
            j++;
        }
        if (j < m && p == q) {
            r.kind = 0;
            break;
        }
        if (hn == (uint32_t)DECF_HEAD) {
            r.kind = 2;
            break;
        }
        const uint32_t e = decf_lookup(t, row, q, &r.bad);
        t.st8(head + hn, e & 0xFFu);
        hn++;
        q += decf_n(e) == 2u ? decf_len0(e) : decf_total(e);
    }
    r.hn = hn;
    r.jsp = j;
    r.kskip = kskip;
    return r;
}

// What a sub-sequence finally holds: `hn` head symbols, then the symbols of records [jsp, m)
struct DecfFinal {
    uint32_t start, end, count, hn, jsp;
};

}  // namespace shafa
