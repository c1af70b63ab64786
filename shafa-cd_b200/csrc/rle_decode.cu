// rle_decode.cu -- expansion of Shafa's escape/run RLE (reference
// rle_block_decompressor, d.c:116-197) for all blocks in ONE pass, outputs
// compacted back to back (= the decoded file, d.c:218,613).  Algorithmic bytes:
// R read + N written.
//
// The byte-serial parser of the reference becomes a 3-state automaton
// (core.h): each thread summarises its 16 bytes as an entry->exit state map,
// maps are composed by a CTA scan (composition = one lookup in a 4 KiB table),
// and chain #1 (decoupled look-back inside the block) carries the state across
// tiles -- tiles whose map is constant (almost all) publish their exit state
// without waiting.  With the entry state known every thread knows its records
// {length, byte}; a CTA scan plus chain #2 (over all tiles of the file) places
// the tile in the output.  Two ways to materialise a tile:
//   * small expansion (<= 16 KiB, the literal-heavy case): every thread writes
//     its own bytes into a staging buffer that is co-aligned with the
//     destination, then the CTA copies it out with 16-byte stores;
//   * large expansion (long runs): the records go to a list in shared memory and
//     the CTA expands it cooperatively, one aligned 16-byte chunk per thread per
//     step (binary search for the chunk's first record), so long runs become
//     full-width stores and no thread ever owns a 255-byte memset.
#include "common.cuh"
#include "core.h"
#include "internal.h"

namespace shafa {

constexpr int RLD_THREADS = 256;
constexpr int RLD_WARPS = RLD_THREADS / 32;
constexpr int RLD_TILE = RLD_THREADS * RLE_ITEMS;          // 4096 RLE bytes
constexpr int RLD_STAGE = 16384;                           // staged expansion limit (bytes)
constexpr uint64_t RLE_BLOCK_CAP = 67108864ull + 1024ull;  // d.c:135,163

struct RleDecShared {
    union {
        struct {
            uint32_t rec_start[RLD_TILE + 1];
            uint8_t rec_sym[RLD_TILE];
        } list;
        alignas(16) uint8_t stage[RLD_STAGE + 32];
    } u;
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

__global__ void __launch_bounds__(RLD_THREADS)
k_rle_decode(const uint8_t *__restrict__ in, const uint64_t *__restrict__ off, const uint64_t *__restrict__ tile0,
             uint32_t nb, uint8_t *__restrict__ out, uint64_t out_capacity, uint64_t *__restrict__ blk_end,
             uint64_t *__restrict__ desc_state_chain, uint64_t *__restrict__ desc_out, uint32_t *__restrict__ counter,
             uint32_t *__restrict__ err)
{
    __shared__ RleDecShared sh;
    const unsigned tid = threadIdx.x, lane = lane_id(), warp = warp_id();
    const uint64_t total_tiles = tile0[nb];
    for (int i = tid; i < 64 * 64; i += RLD_THREADS) sh.compose[i] = (uint8_t)fsm_compose((uint32_t)i >> 6, (uint32_t)i & 63u);

    for (;;) {
        __syncthreads();
        if (tid == 0) {
            const uint32_t tk = atomicAdd(counter, 1u);
            sh.ticket = tk;
            if (tk < total_tiles) sh.block = find_block(tile0, nb, tk);
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
        const uint64_t pos = base + (uint64_t)tid * 16;

        // ---- load, valid range [j0, j1) of this thread ----
        uint4 v = make_uint4(0, 0, 0, 0);
        int j0 = 0, j1 = 0;
        if (pos + 16 > lo && pos < hi) {
            v = ld_stream16(in + pos);
            j0 = (lo > pos) ? (int)(lo - pos) : 0;
            j1 = (hi < pos + 16) ? (int)(hi - pos) : 16;
        }
        const int n = j1 - j0;
        const uint32_t zmask = (zero_mask16(v) >> j0) & ((n >= 16) ? 0xffffu : ((1u << n) - 1u));

        // ---- state maps: thread -> warp -> CTA (composition through the table) ----
        uint32_t incl = fsm_thread_map_fast(zmask, n);
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(FULL, incl, d);
            if (lane >= (unsigned)d) incl = sh.compose[o * 64 + incl];
        }
        if (lane == 31) sh.wmap[warp] = incl;
        uint32_t excl_map = __shfl_up_sync(FULL, incl, 1);
        if (lane == 0) excl_map = FSM_IDENTITY;
        __syncthreads();

        // ---- warp 0: scan of the warp maps, then chain #1 (automaton state at the tile start) ----
        if (warp == 0) {
            uint32_t wm = (lane < RLD_WARPS) ? sh.wmap[lane] : FSM_IDENTITY;
            uint32_t wi = wm;
#pragma unroll
            for (int d = 1; d < RLD_WARPS; d <<= 1) {
                const uint32_t o = __shfl_up_sync(FULL, wi, d);
                if (lane >= (unsigned)d) wi = sh.compose[o * 64 + wi];
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

        // ---- records of this thread: count + output bytes ----
        const bool all_literal = (n == 16) && (zmask == 0) && (my_entry == 0);
        uint32_t prevb = 0;
        if (n > 0 && my_entry == 2 && pos + j0 > lo) prevb = in[pos + j0 - 1];
        auto bytef = [&](int q) -> uint32_t { return byte_of(v, j0 + q); };
        uint32_t recs = 16, obytes = 16;
        if (!all_literal) {
            recs = 0;
            obytes = 0;
            rle_parse_thread(my_entry, n, prevb, bytef, [&](uint32_t len, uint32_t) {
                recs++;
                obytes += len;
            });
        }
        uint32_t packed_total;
        const uint32_t packed_excl = block_excl_scan<RLD_THREADS>(obytes | (recs << 19), sh.scan, &packed_total);
        const uint32_t tile_out = packed_total & ((1u << 19) - 1u);
        const uint32_t tile_recs = packed_total >> 19;
        if (warp == 0) {
            const uint64_t pfx = lookback_sum(desc_out, (int64_t)t, 0, tile_out);
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
        if (tile_out == 0) continue;
        const uint32_t my_out = packed_excl & ((1u << 19) - 1u);

        if (tile_out <= (uint32_t)RLD_STAGE) {
            // ---- staged expansion: every thread writes its own bytes ----
            const uint32_t shift = (uint32_t)(out0 & 15u);
            uint8_t *dst = sh.u.stage + shift + my_out;
            if (all_literal) {
#pragma unroll
                for (int j = 0; j < 16; j++) dst[j] = (uint8_t)byte_of(v, j);
            } else {
                rle_parse_thread(my_entry, n, prevb, bytef, [&](uint32_t len, uint32_t sym) {
                    for (uint32_t k = 0; k < len; k++) dst[k] = (uint8_t)sym;
                    dst += len;
                });
            }
            __syncthreads();
            uint8_t *g = out + (out0 - shift);  // 16-byte aligned
            const uint32_t end = shift + tile_out;
            for (uint32_t c = tid * 16; c < end; c += RLD_THREADS * 16) {
                if (c >= shift && c + 16 <= end) {
                    st_stream16(g + c, *reinterpret_cast<const uint4 *>(sh.u.stage + c));
                } else {
                    for (uint32_t k = (c > shift ? c : shift); k < c + 16 && k < end; k++) g[k] = sh.u.stage[k];
                }
            }
        } else {
            // ---- large expansion: record list + cooperative fill ----
            {
                uint32_t o = my_out;
                uint32_t r = packed_excl >> 19;
                if (all_literal) {
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        sh.u.list.rec_start[r + j] = o + j;
                        sh.u.list.rec_sym[r + j] = (uint8_t)byte_of(v, j);
                    }
                } else {
                    rle_parse_thread(my_entry, n, prevb, bytef, [&](uint32_t len, uint32_t sym) {
                        sh.u.list.rec_start[r] = o;
                        sh.u.list.rec_sym[r] = (uint8_t)sym;
                        r++;
                        o += len;
                    });
                }
                if (tid == RLD_THREADS - 1) sh.u.list.rec_start[tile_recs] = tile_out;
            }
            __syncthreads();
            const uint64_t a0 = out0 & ~15ull;
            const uint64_t a1 = out0 + tile_out;
            for (uint64_t g = a0 + (uint64_t)tid * 16; g < a1; g += (uint64_t)RLD_THREADS * 16) {
                const int64_t r0 = (int64_t)g - (int64_t)out0;  // tile-relative offset of the chunk (may be < 0)
                const uint32_t first = r0 < 0 ? 0u : (uint32_t)r0;
                uint32_t lo_i = 0, hi_i = tile_recs;  // largest ri with rec_start[ri] <= first
                while (hi_i - lo_i > 1) {
                    const uint32_t mid = (lo_i + hi_i) >> 1;
                    if (sh.u.list.rec_start[mid] <= first)
                        lo_i = mid;
                    else
                        hi_i = mid;
                }
                uint32_t ri = lo_i;
                const bool full = (r0 >= 0) && (r0 + 16 <= (int64_t)tile_out);
                if (full && sh.u.list.rec_start[ri + 1] >= first + 16) {
                    const uint32_t w = (uint32_t)sh.u.list.rec_sym[ri] * 0x01010101u;
                    st_stream16(out + g, make_uint4(w, w, w, w));
                } else {
                    uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
                    for (int k = 0; k < 16; k++) {
                        const int64_t r = r0 + k;
                        if (r >= 0 && r < (int64_t)tile_out) {
                            while (sh.u.list.rec_start[ri + 1] <= (uint32_t)r) ri++;
                            w[k >> 2] |= (uint32_t)sh.u.list.rec_sym[ri] << (8 * (k & 3));
                        }
                    }
                    if (full) {
                        st_stream16(out + g, make_uint4(w[0], w[1], w[2], w[3]));
                    } else {
#pragma unroll
                        for (int k = 0; k < 16; k++) {
                            const int64_t r = r0 + k;
                            if (r >= 0 && r < (int64_t)tile_out) out[g + k] = (uint8_t)(w[k >> 2] >> (8 * (k & 3)));
                        }
                    }
                }
            }
        }
    }
}

int launch_rle_decode(shafa_cuda_ctx *ctx, const uint8_t *d_rle, const uint64_t *h_off, uint64_t nb, uint8_t *d_out,
                      uint64_t out_capacity, uint64_t *h_final_sizes, uint32_t *d_err)
{
    if (nb == 0) return SHAFA_SUCCESS;
    if (((uintptr_t)d_rle & 15) || ((uintptr_t)d_out & 15)) return set_error(ctx, SHAFA_FILE_STREAM_FAILED, "rle_decode: buffers must be 16-byte aligned");
    static int occ = 0;
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
