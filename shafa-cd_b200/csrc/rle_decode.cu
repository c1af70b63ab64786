// rle_decode.cu -- expansion of Shafa's escape/run RLE (reference
// rle_block_decompressor, d.c:116-197) for all blocks in ONE pass, outputs
// compacted back to back (= the decoded file, d.c:218,613).  Algorithmic bytes:
// R read + N written.
//
// The byte-serial parser of the reference is a 3-state automaton (core.h).  A
// thread owns 64 contiguous RLE bytes; only escape bytes (0x00 met in state 0)
// change the state, so its entry->exit map and its emission masks come from
// walking the zero-byte bitmask, not the bytes.  Maps are composed by a CTA scan
// (composition = one lookup in a 4 KiB table) and chain #1 (decoupled look-back
// inside the block) carries the state across 16 KiB tiles -- tiles whose map is
// constant (almost all) publish their exit state without waiting.  A CTA scan of
// output sizes plus chain #2 (over all tiles of the file) places the tile in the
// output.  The tile's size is published right after the scan; its output is
// materialised in a 24 KiB staging buffer at tile-relative offsets, the prefix is
// resolved behind that work, and the copy-out realigns (copy_out_unaligned).  A
// tile that expands past the buffer (long runs) is placed first and expanded
// straight into global memory, each warp filling the runs of its own 32 threads
// with 16-byte stores.
#include "common.cuh"
#include "core.h"
#include "internal.h"

namespace shafa {

constexpr int RLD_THREADS = 256;
constexpr int RLD_WARPS = RLD_THREADS / 32;
constexpr int RLD_VECS = 4;
constexpr int RLD_PER_THREAD = 16 * RLD_VECS;             // 64 RLE bytes per thread
constexpr int RLD_TILE = RLD_THREADS * RLD_PER_THREAD;    // 16 KiB of RLE bytes per tile
constexpr int RLD_STAGE = 24576;                          // staging window (bytes of output)
constexpr uint64_t RLE_BLOCK_CAP = 67108864ull + 1024ull;  // d.c:135,163

struct RleDecShared {
    alignas(16) uint8_t stage[RLD_STAGE + 32];  // + the read-ahead of copy_out_unaligned
    uint8_t compose[64 * 64];  // compose[a*64+b] = first a, then b
    uint32_t scan[RLD_WARPS + 1];
    uint32_t wmap[RLD_WARPS];
    uint32_t wexcl[RLD_WARPS];
    uint64_t out_prefix;
    uint32_t tile_map;
    uint32_t entry;
    uint32_t ticket;
    uint32_t block;
};

__device__ __forceinline__ uint4 rld_ld16(const void *p)
{
    uint4 v;  // default L1 policy: the four vectors of a thread share sectors
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// One position of the uniform literal pass, as predicated instructions on a 32-bit shared address:
// store the byte if the position is a literal (bit `BIT` of l16) and advance by one; if it is a count byte
// (bit of r16) advance by the run's length, max(count, 1) (d.c:179-184: count 0 -> one copy).
template <uint32_t BIT>
__device__ __forceinline__ void rld_lit_step(uint32_t &o_s, uint32_t cur, uint32_t l16, uint32_t r16)
{
    asm volatile(
        "{\n\t.reg .pred pl, pr;\n\t.reg .u32 t;\n\t"
        "and.b32 t, %2, %4;\n\tsetp.ne.u32 pl, t, 0;\n\t"
        "and.b32 t, %3, %4;\n\tsetp.ne.u32 pr, t, 0;\n\t"
        "@pl st.shared.u8 [%0], %1;\n\t"
        "@pl add.u32 %0, %0, 1;\n\t"
        "max.u32 t, %1, 1;\n\t"
        "@pr add.u32 %0, %0, t;\n\t}"
        : "+r"(o_s)
        : "r"(cur), "r"(l16), "r"(r16), "n"(BIT)
        : "memory");
}
template <int J>
__device__ __forceinline__ void rld_lit_steps(uint32_t &o_s, const uint4 &x, uint32_t l16, uint32_t r16)
{
    if constexpr (J < 16) {
        const uint32_t w = (J < 4) ? x.x : (J < 8) ? x.y : (J < 12) ? x.z : x.w;
        rld_lit_step<(1u << J)>(o_s, __byte_perm(w, 0, 0x4440 | (J & 3)), l16, r16);
        rld_lit_steps<J + 1>(o_s, x, l16, r16);
    }
}

// `len` copies of `sym` at dst (shared memory): bytes up to a 4-byte boundary, words, tail bytes
__device__ __noinline__ void rld_fill(uint8_t *dst, uint32_t sym, uint32_t len)
{
    if (len < 8) {
        for (uint32_t k = 0; k < len; k++) dst[k] = (uint8_t)sym;
        return;
    }
    const uint32_t head = (4u - ((uint32_t)(uintptr_t)dst & 3u)) & 3u;
    for (uint32_t k = 0; k < head; k++) dst[k] = (uint8_t)sym;
    dst += head;
    len -= head;
    const uint32_t w = sym * 0x01010101u;
    uint32_t *d32 = reinterpret_cast<uint32_t *>(dst);
    const uint32_t nw = len >> 2;
    for (uint32_t k = 0; k < nw; k++) d32[k] = w;
    for (uint32_t k = nw * 4; k < len; k++) dst[k] = (uint8_t)sym;
}

// one output piece [o, o+len) of byte `sym` (tile-relative offsets), clipped to the window [wlo, whi)
template <bool CLIP>
__device__ __forceinline__ void rld_emit(uint8_t *stage_base, uint32_t wlo, uint32_t whi, uint32_t o, uint32_t len, uint32_t sym)
{
    if (!CLIP) {
        if (len == 1)
            stage_base[o] = (uint8_t)sym;
        else
            rld_fill(stage_base + o, sym, len);
    } else {
        const uint32_t lo = o > wlo ? o : wlo;
        const uint32_t hi = (o + len < whi) ? (o + len) : whi;
        if (lo < hi) rld_fill(stage_base + lo, sym, hi - lo);
    }
}

// `len` (<= 255) copies of `sym` at dst in global memory, by the 32 lanes of a warp: bytes up to the first
// 16-byte boundary, 16-byte stores, tail bytes
__device__ __forceinline__ void rld_fill_global(uint8_t *dst, uint32_t sym, uint32_t len, unsigned lane)
{
    const uint32_t mis = (16u - (uint32_t)((uintptr_t)dst & 15u)) & 15u;
    const uint32_t head = mis < len ? mis : len;
    if (lane < head) dst[lane] = (uint8_t)sym;
    const uint32_t body = (len - head) >> 4;  // <= 15 chunks
    const uint32_t w = sym * 0x01010101u;
    if (lane < body) st_stream16(dst + head + 16 * lane, make_uint4(w, w, w, w));
    const uint32_t done = head + 16 * body;
    if (lane < len - done) dst[done + lane] = (uint8_t)sym;
}

// A warp expands the records of its own 32 threads straight into global memory, one record at a time (tiles
// whose output does not fit the staging buffer).  `out_tile` is where the tile's output starts; lit / run are the
// thread's literal and count-byte positions relative to its first valid byte j0.  Not inlined: keeps the
// registers of this rare path out of the common one.
__device__ __noinline__ void rld_expand_direct(const uint8_t *__restrict__ in, uint8_t *__restrict__ out_tile, uint64_t pos, int j0,
                                               uint32_t obytes, uint32_t my_out, uint64_t lit, uint64_t run, uint32_t prevb)
{
    const unsigned lane = lane_id();
    for (int src = 0; src < 32; src++) {
        if (__shfl_sync(FULL, obytes, src) == 0) continue;  // uniform
        const uint64_t lt = __shfl_sync(FULL, lit, src), rn = __shfl_sync(FULL, run, src);
        const uint32_t prev_src = __shfl_sync(FULL, prevb, src);
        const uint8_t *psrc = in + (pos - ((uint64_t)lane - (uint64_t)src) * RLD_PER_THREAD) + __shfl_sync(FULL, j0, src);
        uint8_t *o = out_tile + __shfl_sync(FULL, my_out, src);
        uint64_t eb = lt | rn;
        while (eb) {  // the same trip for every lane: all operands are the source lane's
            const int q = ctz64(eb);
            if ((rn >> q) & 1ull) {
                const uint32_t cnt = psrc[q];  // every lane reads the same byte (broadcast)
                const uint32_t len = cnt ? cnt : 1u;
                const uint32_t sym = q ? (uint32_t)psrc[q - 1] : prev_src;
                rld_fill_global(o, sym, len, lane);
                o += len;
                eb &= eb - 1;
            } else {
                // literals at consecutive positions are consecutive in the output: copy the segment
                const uint64_t rest = lt >> q;
                const uint32_t seg = (~rest) ? (uint32_t)ctz64(~rest) : 64u;
                for (uint32_t k = lane; k < seg; k += 32) o[k] = psrc[q + k];
                o += seg;
                eb &= (seg >= 64u) ? 0ull : ~(((1ull << seg) - 1ull) << q);
            }
        }
    }
}

#ifndef SHAFA_RLD_MIN_CTAS
#define SHAFA_RLD_MIN_CTAS 6
#endif
__global__ void __launch_bounds__(RLD_THREADS, SHAFA_RLD_MIN_CTAS)
k_rle_decode(const uint8_t *__restrict__ in, const uint64_t *__restrict__ off, const uint64_t *__restrict__ tile0,
             uint32_t nb, uint8_t *__restrict__ out, uint64_t out_capacity, uint64_t *__restrict__ blk_end,
             uint64_t *__restrict__ desc_state_chain, uint64_t *__restrict__ desc_out, uint32_t *__restrict__ counter,
             uint32_t *__restrict__ err)
{
    __shared__ RleDecShared sh;
    const unsigned tid = threadIdx.x, lane = lane_id(), warp = warp_id();
    const uint64_t total_tiles = tile0[nb];
    for (int i = tid; i < 64 * 64; i += RLD_THREADS) sh.compose[i] = (uint8_t)fsm_compose((uint32_t)i >> 6, (uint32_t)i & 63u);
    const uint32_t compose_s = smem_opaque(sh.compose), stage_s = smem_opaque(sh.stage);

    // A ticket is taken only when the CTA is ready to work on it (a tile reserved ahead of time stalls every later
    // tile of the look-back chains: measured).  Warp 0 finds the ticket's block by probing the 32 blocks after its
    // previous one at once instead of a dependent binary search.
    uint32_t last_block = 0;
    for (;;) {
        __syncthreads();
        if (warp == 0) {
            uint32_t tk = 0;
            if (lane == 0) tk = atomicAdd(counter, 1u);
            tk = __shfl_sync(FULL, tk, 0);
            if (tk < total_tiles) last_block = find_block_forward(tile0, nb, tk, last_block);
            if (lane == 0) {
                sh.ticket = tk;
                sh.block = last_block;
            }
        }
        __syncthreads();
        const uint64_t t = sh.ticket;
        if (t >= total_tiles) break;
        const uint32_t b = sh.block;
        const uint64_t lo = off[b], hi = off[b + 1];
        const uint64_t first_tile = tile0[b];
        const bool is_first = (t == first_tile);
        const bool is_last = (t + 1 == tile0[b + 1]);
        const uint64_t base = (lo / RLD_TILE + (t - first_tile)) * RLD_TILE;
        const uint64_t pos = base + (uint64_t)tid * RLD_PER_THREAD;

        // ---- load; valid range [j0, j1) of this thread's 64 bytes ----
        int j0 = 0, j1 = 0;
        if (pos + RLD_PER_THREAD > lo && pos < hi) {
            j0 = (lo > pos) ? (int)(lo - pos) : 0;
            j1 = (hi < pos + RLD_PER_THREAD) ? (int)(hi - pos) : RLD_PER_THREAD;
        }
        const int n = j1 - j0;
        const bool full = (n == RLD_PER_THREAD);
        // The bytes themselves are not kept: the write pass below reads them again (L1 hits), which
        // keeps the kernel at 64 registers without spilling.
        uint64_t Z = 0;  // bit q = (valid byte q == 0), relative to j0
        uint32_t last_byte = 0;
#pragma unroll
        for (int c = 0; c < RLD_VECS; c++) {
            uint4 x = make_uint4(0, 0, 0, 0);
            if (16 * c < j1 && 16 * c + 16 > j0) x = rld_ld16(in + pos + 16 * c);
            Z |= (uint64_t)zero_mask16(x) << (16 * c);
            if (c == RLD_VECS - 1) last_byte = x.w >> 24;
        }
        Z = (j0 < 64) ? (Z >> j0) : 0ull;
        Z &= (n >= 64) ? ~0ull : ((1ull << n) - 1ull);

        // ---- state maps: thread -> warp -> CTA (composition through the table) ----
        uint32_t incl = fsm_map64(Z, n);
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(FULL, incl, d);
            if (lane >= (unsigned)d) incl = lds_u8(compose_s + o * 64 + incl);
        }
        if (lane == 31) sh.wmap[warp] = incl;
        uint32_t excl_map = __shfl_up_sync(FULL, incl, 1);
        if (lane == 0) excl_map = FSM_IDENTITY;
        // byte just before this thread's first valid byte (needed when a count byte is my first byte)
        uint32_t prevb = __shfl_up_sync(FULL, last_byte, 1);
        if ((lane == 0 || !full) && n > 0 && pos + j0 > lo) prevb = in[pos + j0 - 1];
        __syncthreads();

        // ---- warp 0: scan of the warp maps, then chain #1 (automaton state at the tile start) ----
        if (warp == 0) {
            uint32_t wm = (lane < RLD_WARPS) ? sh.wmap[lane] : FSM_IDENTITY;
            uint32_t wi = wm;
#pragma unroll
            for (int d = 1; d < RLD_WARPS; d <<= 1) {
                const uint32_t o = __shfl_up_sync(FULL, wi, d);
                if (lane >= (unsigned)d) wi = lds_u8(compose_s + o * 64 + wi);
            }
            uint32_t we = __shfl_up_sync(FULL, wi, 1);
            if (lane == 0) we = FSM_IDENTITY;
            if (lane < RLD_WARPS) sh.wexcl[lane] = we;
            const uint32_t tile_map = __shfl_sync(FULL, wi, RLD_WARPS - 1);
            uint32_t entry = 0;
            if (is_first) {
                if (lane == 0) desc_store(desc_state_chain + t, desc_pack(DESC_PREFIX, fsm_apply(tile_map, 0)));
            } else {
                const bool cst = fsm_is_const(tile_map);
                if (lane == 0)
                    desc_store(desc_state_chain + t,
                               cst ? desc_pack(DESC_PREFIX, fsm_apply(tile_map, 0)) : desc_pack(DESC_AGG, tile_map));
                // walk back: pending = composition of the maps between the found prefix and this tile
                uint32_t pending = FSM_IDENTITY;
                int64_t p = (int64_t)t - 1;
                for (;;) {
                    const int64_t idx = p - (int64_t)lane;
                    uint64_t d = desc_pack(DESC_PREFIX, 0);  // before the block: state 0
                    if (idx >= (int64_t)first_tile) d = desc_load(desc_state_chain + idx);
                    while (__any_sync(FULL, desc_state(d) == DESC_INVALID)) {
                        if (idx >= (int64_t)first_tile && desc_state(d) == DESC_INVALID) d = desc_load(desc_state_chain + idx);
                    }
                    const unsigned pm = __ballot_sync(FULL, desc_state(d) == DESC_PREFIX);
                    const int upto = pm ? (__ffs(pm) - 1) : 32;  // lanes 0..upto-1 hold maps, lane upto the state
                    uint32_t win = FSM_IDENTITY;
                    for (int i = upto - 1; i >= 0; i--) {  // oldest first
                        const uint32_t m = (uint32_t)desc_value(__shfl_sync(FULL, d, i));
                        win = fsm_compose(win, m);
                    }
                    pending = fsm_compose(win, pending);
                    if (pm) {
                        const uint32_t st = (uint32_t)desc_value(__shfl_sync(FULL, d, upto));
                        entry = fsm_apply(pending, st);
                        break;
                    }
                    p -= 32;
                }
                if (!cst && lane == 0)
                    desc_store(desc_state_chain + t, desc_pack(DESC_PREFIX, fsm_apply(tile_map, entry)));
            }
            if (lane == 0) {
                sh.entry = entry;
                sh.tile_map = tile_map;
            }
        }
        __syncthreads();
        const uint32_t tile_entry = sh.entry;
        const uint32_t my_entry = fsm_apply(excl_map, fsm_apply(sh.wexcl[warp], tile_entry));

        // ---- what do my bytes emit: literal positions, count-byte positions, output size ----
        const RleDecMasks m = rle_dec_masks64(fsm_walk64(Z, n, my_entry).esc, n, my_entry);
        uint32_t obytes = (uint32_t)__popcll(m.lit);
        {
            uint64_t rb = m.run;
            while (rb) {  // sparse: one iteration per triple
                const int q = ctz64(rb);
                rb &= rb - 1;
                const uint32_t cnt = in[pos + j0 + q];  // L1 hit: the line was just loaded
                obytes += cnt ? cnt : 1u;               // d.c:179-184: count 0 -> one copy
            }
        }
        uint32_t tile_out;
        const uint32_t my_out = block_excl_scan<RLD_THREADS>(obytes, sh.scan, &tile_out);
        if (tid == 0) lookback_publish(desc_out, (int64_t)t, 0, tile_out);  // the size now, the place after the staging

        const bool single = tile_out <= (uint32_t)RLD_STAGE;
        if (!single) {
            // ---- the tile expands past the staging buffer (long runs): place it first, then every warp expands
            // the records of its own 32 threads straight into global memory, one record at a time ----
            if (warp == 0) {
                const uint64_t pfx = lookback_resolve(desc_out, (int64_t)t, 0, tile_out);
                if (lane == 0) sh.out_prefix = pfx;
            }
            __syncthreads();
            const uint64_t big0 = sh.out_prefix;
            if (is_last && tid == 0) {
                blk_end[b] = big0 + tile_out;
                if (fsm_apply(sh.tile_map, tile_entry) != 0) atomicOr(err, (uint32_t)ERR_RLE_TRUNCATED);  // block ends inside a triple
            }
            if (big0 + tile_out > out_capacity) {
                if (tid == 0) atomicOr(err, (uint32_t)ERR_CAPACITY);
                continue;
            }
            rld_expand_direct(in, out + big0, pos, j0, obytes, my_out, m.lit, m.run, prevb);
            continue;
        }

        // ---- materialise the tile's output at tile-relative offsets in the staging buffer ----
        {
            const uint32_t w0 = 0, w1 = tile_out;  // the whole tile is one window of the stage
            // my output occupies tile-space [a, a+obytes)
            const uint32_t a = my_out;
            if (obytes) {
                uint8_t *sb = sh.stage;  // sb[x] is tile-space position x
                uint32_t o = a;
                if (full) {
                    uint32_t o_s = stage_s + a;
                    // literals: one uniform, fully predicated pass -- the offset of position q is
                    // (literals before q) + (bytes of the runs before q), kept as a running sum
#pragma unroll 1
                    for (int c = 0; c < RLD_VECS; c++) {
                        const uint4 x = rld_ld16(in + pos + 16 * c);
                        const uint32_t l16 = (uint32_t)(m.lit >> (16 * c)) & 0xffffu, r16 = (uint32_t)(m.run >> (16 * c)) & 0xffffu;
                        if (__all_sync(__activemask(), r16 == 0 && l16 == 0xffffu)) {  // whole warp: sixteen literals each
#pragma unroll
                            for (int j = 0; j < 16; j++)
                                asm volatile("st.shared.u8 [%0], %1;" ::"r"(o_s + j), "r"(byte_of(x, j)) : "memory");
                            o_s += 16;
                        } else {
                            rld_lit_steps<0>(o_s, x, l16, r16);
                        }
                    }
                } else {
                    // partial threads (block edges) and multi-window tiles: generic walk over the emitting positions
                    uint64_t eb = m.lit | m.run;
#pragma unroll 1
                    while (eb) {
                        const int q = ctz64(eb);
                        eb &= eb - 1;
                        const uint32_t cur = in[pos + j0 + q];
                        if ((m.run >> q) & 1ull) {
                            const uint32_t sym = (q > 0) ? (uint32_t)in[pos + j0 + q - 1] : prevb;
                            const uint32_t len = cur ? cur : 1u;
                            rld_emit<true>(sb, w0, w1, o, len, sym);
                            o += len;
                        } else {
                            rld_emit<true>(sb, w0, w1, o, 1u, cur);
                            o += 1;
                        }
                    }
                }
            }
            {
                // runs.  Short ones (<= 16 bytes: the common case in text) are written by their own lane, all
                // lanes in parallel; long ones are filled by the whole warp, one run at a time.
                uint64_t long_runs = 0;
                if (full) {
                    uint64_t rb = m.run;
                    uint32_t cum = 0;
#pragma unroll 1
                    while (rb) {
                        const int q = ctz64(rb);
                        rb &= rb - 1;
                        const uint32_t cnt = in[pos + q];  // L1 hit: the line was just loaded
                        const uint32_t len = cnt ? cnt : 1u;
                        if (len <= 16) {
                            const uint32_t sym = q ? (uint32_t)in[pos + q - 1] : prevb;
                            const uint32_t d = stage_s + my_out + (uint32_t)__popcll(m.lit & ((1ull << q) - 1ull)) + cum;
                            for (uint32_t k = 0; k < len; k++) sts_u8(d + k, sym);
                        } else {
                            long_runs |= 1ull << q;
                        }
                        cum += len;
                    }
                }
                unsigned has = __ballot_sync(FULL, long_runs != 0);
                while (has) {
                    const int src = __ffs(has) - 1;
                    has &= has - 1;
                    uint64_t rb = __shfl_sync(FULL, m.run, src);  // that lane's count-byte positions
                    const uint64_t lr = __shfl_sync(FULL, long_runs, src);
                    const uint64_t lt = __shfl_sync(FULL, m.lit, src);
                    const uint32_t a_src = __shfl_sync(FULL, my_out, src);
                    const uint32_t prev_src = __shfl_sync(FULL, prevb, src);
                    const uint64_t pos_src = pos - ((uint64_t)lane - (uint64_t)src) * RLD_PER_THREAD;
                    uint32_t cum = 0;  // bytes of the runs before the current one
                    while (rb) {
                        const int q = ctz64(rb);
                        rb &= rb - 1;
                        const uint32_t cnt = in[pos_src + q];  // every lane reads the same byte (broadcast)
                        const uint32_t len = cnt ? cnt : 1u;
                        if ((lr >> q) & 1ull) {
                            const uint32_t sym = q ? (uint32_t)in[pos_src + q - 1] : prev_src;
                            const uint32_t o = a_src + (uint32_t)__popcll(lt & ((1ull << q) - 1ull)) + cum;
                            for (uint32_t k = lane; k < len; k += 32) sts_u8(stage_s + o + k, sym);
                        }
                        cum += len;
                    }
                }
            }
            // the tile's place in the output, looked up behind the staging work
            if (warp == 0) {
                const uint64_t pfx = lookback_resolve(desc_out, (int64_t)t, 0, tile_out);
                if (lane == 0) sh.out_prefix = pfx;
            }
            __syncthreads();
            const uint64_t out0 = sh.out_prefix;
            if (is_last && tid == 0) {
                blk_end[b] = out0 + tile_out;
                if (fsm_apply(sh.tile_map, tile_entry) != 0) atomicOr(err, (uint32_t)ERR_RLE_TRUNCATED);  // block ends inside a triple
            }
            if (out0 + tile_out > out_capacity) {
                if (tid == 0) atomicOr(err, (uint32_t)ERR_CAPACITY);
                continue;
            }
            if (tile_out) copy_out_unaligned<RLD_THREADS>(out + out0, sh.stage, tile_out);
        }
    }
}

int launch_rle_decode(shafa_cuda_ctx *ctx, const uint8_t *d_rle, const uint64_t *h_off, uint64_t nb, uint8_t *d_out,
                      uint64_t out_capacity, uint64_t *h_final_sizes, uint32_t *d_err)
{
    if (nb == 0) return SHAFA_SUCCESS;
    if (((uintptr_t)d_rle & 15) || ((uintptr_t)d_out & 15)) return set_error(ctx, SHAFA_FILE_STREAM_FAILED, "rle_decode: buffers must be 16-byte aligned");
    int &occ = ctx->occ[OCC_RLE_DEC];
    if (!occ) {
        SHAFA_CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_rle_decode, RLD_THREADS, 0));
        if (occ < 1) occ = 1;
    }
    uint64_t *hp = (uint64_t *)pinned_get(ctx, (nb + 1) * 2 * sizeof(uint64_t));
    uint64_t *dp = (uint64_t *)scratch_get(ctx, SCR_PLAN, (nb + 1) * 2 * sizeof(uint64_t));
    if (!hp || !dp) return SHAFA_LACK_OF_MEMORY;
    uint64_t tiles = 0;
    for (uint64_t b = 0; b < nb; b++) {
        hp[b] = h_off[b];
        hp[nb + 1 + b] = tiles;
        const uint64_t lo = h_off[b], hi = h_off[b + 1];
        if (hi < lo) return set_error(ctx, SHAFA_FILE_STREAM_FAILED, "rle_decode: offsets not monotone");
        if (hi > lo) tiles += (hi - 1) / RLD_TILE - lo / RLD_TILE + 1;
    }
    hp[nb] = h_off[nb];
    hp[2 * nb + 1] = tiles;
    if (tiles > 0xfffffff0ull) return set_error(ctx, SHAFA_LACK_OF_MEMORY, "rle_decode: too many tiles");
    SHAFA_CUDA_TRY(ctx, cudaMemcpyAsync(dp, hp, (nb + 1) * 2 * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    uint64_t *desc = (uint64_t *)scratch_get(ctx, SCR_DESC, 2 * (tiles + 1) * sizeof(uint64_t));
    uint64_t *blk_end = (uint64_t *)scratch_get(ctx, SCR_SIZES, nb * sizeof(uint64_t));
    uint32_t *counter = misc_ptr(ctx, MISC_COUNTER);
    if (!desc || !blk_end || !counter) return SHAFA_LACK_OF_MEMORY;
    SHAFA_CUDA_TRY(ctx, cudaMemsetAsync(desc, 0, 2 * (tiles + 1) * sizeof(uint64_t), ctx->stream));
    SHAFA_CUDA_TRY(ctx, cudaMemsetAsync(counter, 0, 4, ctx->stream));
    // empty blocks never run a tile: their ends are filled in on the host side below
    SHAFA_CUDA_TRY(ctx, cudaMemsetAsync(blk_end, 0xff, nb * sizeof(uint64_t), ctx->stream));
    if (tiles) {
        uint64_t grid = (uint64_t)ctx->sm_count * occ;
        if (grid > tiles) grid = tiles;
        SHAFA_LAUNCH(ctx, "rle_decode", k_rle_decode, dim3((unsigned)grid), dim3(RLD_THREADS), 0, d_rle, dp, dp + nb + 1,
                     (uint32_t)nb, d_out, out_capacity, blk_end, desc, desc + tiles + 1, counter, d_err);
        SHAFA_CHECK_LAUNCH(ctx);
    }
    uint64_t *h_end = (uint64_t *)pinned_get(ctx, nb * sizeof(uint64_t));
    if (!h_end) return SHAFA_LACK_OF_MEMORY;
    SHAFA_CUDA_TRY(ctx, cudaMemcpyAsync(h_end, blk_end, nb * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    SHAFA_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    uint64_t prev = 0;
    int rc = SHAFA_SUCCESS;
    for (uint64_t b = 0; b < nb; b++) {
        const uint64_t e = (h_end[b] == ~0ull) ? prev : h_end[b];
        h_final_sizes[b] = e - prev;
        if (h_final_sizes[b] > RLE_BLOCK_CAP) rc = SHAFA_FILE_UNRECOGNIZABLE;
        prev = e;
    }
    if (rc) return set_error(ctx, rc, "rle_decode: a block expands past 64 MiB + 1 KiB (d.c:165-168)");
    return SHAFA_SUCCESS;
}

}  // namespace shafa
