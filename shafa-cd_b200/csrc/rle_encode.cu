// rle_encode.cu -- Shafa's escape/run RLE (reference block_compression, f.c:29-55)
// for all blocks of a file in ONE pass, with the histogram of the emitted bytes
// fused (f.c:310) and the blocks' outputs compacted back to back (= the .rle
// file, f.c:307).  Algorithmic bytes: N read + R written.
//
// Tile = 16 KiB of one block (tiles never straddle a block), 256 threads x 64
// contiguous bytes held in registers (four 16-byte vectors).  Per tile:
//   1. loads; "continuation" bits E (byte equals its predecessor) and zero bits Z as
//      64-bit masks per thread (exact byte tests + dp4a gathers, common.cuh); the
//      neighbours' edge bits arrive by warp shuffles, the warp edges read one extra
//      word.  ~E is the run-head bitmask, kept in shared memory together with a
//      288-position look-ahead halo (positions at/after the block end count as heads).
//   2. chain #1 (decoupled look-back inside the block): length of the run that
//      is still open at the tile start (needed for k mod 255).
//   3. what every position emits: the bit-parallel rule shafa::rle_fast_masks64
//      (core.h); inside a run longer than 255 bytes the first position of every
//      further 255-byte chunk is turned into a head first (rle_chunk_heads).
//   4. CTA scan of emitted bytes; the tile's size is published at once (chain #2,
//      look-back over ALL tiles of the file).  The chain's prefix frontier moves
//      one look-back window per L2 round trip, which is why tiles are 16 KiB and
//      the window 128 descriptors.
//   5. bytes are staged in shared memory at tile-relative offsets (uniform
//      predicated literal pass + sparse triple loop); chain #2 is resolved behind
//      that work and the copy-out realigns (copy_out_unaligned).  The same staging
//      buffer feeds the fused 256-bin histogram (8 striped sub-histograms to
//      spread equal-byte hits).
#include "common.cuh"
#include "core.h"
#include "internal.h"

namespace shafa {

constexpr int RLE_WARPS = RLE_THREADS / 32;
constexpr int RLE_PER_THREAD = RLE_ITEMS * RLE_VECS;     // 64
constexpr int RLE_STAGE = 2 * RLE_TILE + 48;             // worst case 2n+3 (f.c:244) + the read-ahead of copy_out_unaligned
constexpr int RLE_HIST_COPIES = 8;
constexpr int RLE_SPARSE_MAX = 3;  // literals per 16-byte vector up to which every lane of a warp takes the sparse literal loop

struct RleEncShared {
    alignas(16) uint8_t stage[RLE_STAGE];
    uint32_t heads[RLE_HEAD_WORDS + 2];
    uint32_t hist[256 * RLE_HIST_COPIES];
    uint32_t scan[RLE_WARPS + 1];
    int32_t lasthead[RLE_WARPS];
    uint64_t out_prefix;
    uint64_t run_in;
    uint32_t ticket;
};

__device__ __forceinline__ uint4 rle_ld16(const void *p)
{
    uint4 v;  // default L1 policy: the four vectors of a thread share sectors
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// One position of the uniform literal pass, as predicated instructions on a 32-bit shared address: store the
// byte and advance by one if the position emits a literal (bit `BIT` of l16); a triple only reserves 3 bytes.
template <uint32_t BIT>
__device__ __forceinline__ void rle_lit_step(uint32_t &o_s, uint32_t cur, uint32_t l16, uint32_t t16)
{
    asm volatile(
        "{\n\t.reg .pred pl, pt;\n\t.reg .u32 t;\n\t"
        "and.b32 t, %2, %4;\n\tsetp.ne.u32 pl, t, 0;\n\t"
        "and.b32 t, %3, %4;\n\tsetp.ne.u32 pt, t, 0;\n\t"
        "@pl st.shared.u8 [%0], %1;\n\t"
        "@pl add.u32 %0, %0, 1;\n\t"
        "@pt add.u32 %0, %0, 3;\n\t}"
        : "+r"(o_s)
        : "r"(cur), "r"(l16), "r"(t16), "n"(BIT)
        : "memory");
}
template <int J>
__device__ __forceinline__ void rle_lit_steps(uint32_t &o_s, const uint4 &x, uint32_t l16, uint32_t t16)
{
    if constexpr (J < 16) {
        const uint32_t w = (J < 4) ? x.x : (J < 8) ? x.y : (J < 12) ? x.z : x.w;
        rle_lit_step<(1u << J)>(o_s, __byte_perm(w, 0, 0x4440 | (J & 3)), l16, t16);
        rle_lit_steps<J + 1>(o_s, x, l16, t16);
    }
}

template <bool WRITE, bool HIST>
__global__ void __launch_bounds__(RLE_THREADS, 5)
k_rle_encode(const uint8_t *__restrict__ in, uint64_t n, uint64_t block_size, uint32_t tiles_per_block,
             uint64_t total_tiles, uint8_t *__restrict__ out, uint64_t out_capacity,
             unsigned long long *__restrict__ hist, uint64_t *__restrict__ blk_end, uint64_t *__restrict__ desc_run,
             uint64_t *__restrict__ desc_out, uint32_t *__restrict__ counter, uint32_t *__restrict__ err)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RleEncShared &sh = *reinterpret_cast<RleEncShared *>(smem_raw);
    const unsigned tid = threadIdx.x, lane = lane_id(), warp = warp_id();

    // A ticket is taken only when the CTA is ready to work on it.  (Taking it one tile ahead, to hide the atomic's
    // round trip, was measured: 1.23 -> 1.77 ms on C2.  A reserved tile that nobody works on yet makes every later
    // tile of both look-back chains wait for it.)
    // (Even taking it a microsecond early, just before the previous tile's copy-out, cost 7 % here.)
    for (;;) {
        __syncthreads();
        if (tid == 0) sh.ticket = atomicAdd(counter, 1u);
        if (HIST) {
            for (int i = tid; i < 256 * RLE_HIST_COPIES; i += RLE_THREADS) sh.hist[i] = 0;
        }
        __syncthreads();
        const uint64_t t = sh.ticket;
        if (t >= total_tiles) break;
        const uint64_t b = t / tiles_per_block;
        const uint32_t ti = (uint32_t)(t % tiles_per_block);
        const uint64_t blk_lo = b * block_size;
        const uint64_t blk_hi = (blk_lo + block_size < n) ? blk_lo + block_size : n;
        const uint64_t tile_lo = blk_lo + (uint64_t)ti * RLE_TILE;
        const uint32_t tile_len = (uint32_t)((blk_hi - tile_lo < RLE_TILE) ? (blk_hi - tile_lo) : RLE_TILE);
        const bool last_tile = (tile_lo + tile_len == blk_hi);
        const uint64_t first_tile_of_block = b * tiles_per_block;
        const uint32_t valid_len = (uint32_t)((blk_hi - tile_lo < (uint64_t)(RLE_TILE + RLE_HALO)) ? (blk_hi - tile_lo)
                                                                                                  : (RLE_TILE + RLE_HALO));
        // ---- 1. load 64 bytes; continuation bits of my positions and of my neighbours' edges ----
        const uint32_t p0 = tid * RLE_PER_THREAD;
        const uint64_t pos = tile_lo + p0;
        const int nvalid = (p0 >= tile_len) ? 0 : (int)((tile_len - p0 < (uint32_t)RLE_PER_THREAD) ? (tile_len - p0) : RLE_PER_THREAD);
        const uint64_t valid64 = (nvalid >= 64) ? ~0ull : ((1ull << nvalid) - 1ull);
        uint4 v[RLE_VECS];
#pragma unroll
        for (int c = 0; c < RLE_VECS; c++) {
            v[c] = make_uint4(0, 0, 0, 0);
            if ((int)(16 * c) < nvalid) v[c] = rle_ld16(in + pos + 16 * c);
        }
        uint32_t prevb = __shfl_up_sync(FULL, v[RLE_VECS - 1].w >> 24, 1);
        uint32_t e_before = 0;  // bit 0 = E[p0-2], bit 1 = E[p0-1]
        if (lane == 0 && pos > blk_lo && nvalid) {
            const uint32_t w = *reinterpret_cast<const uint32_t *>(in + pos - 4);  // bytes pos-4 .. pos-1, inside the block
            prevb = w >> 24;
            const uint32_t b1 = (w >> 8) & 0xffu, b2 = (w >> 16) & 0xffu;
            e_before = ((b2 == b1) ? 1u : 0u) | ((prevb == b2) ? 2u : 0u);
        }
        uint64_t E = 0, Z = 0;
#pragma unroll
        for (int c = 0; c < RLE_VECS; c++) {
            const uint32_t pb = (c == 0) ? prevb : (v[c - 1].w >> 24);
            E |= (uint64_t)eq_prev_mask16(v[c], pb) << (16 * c);
            Z |= (uint64_t)zero_mask16(v[c]) << (16 * c);
        }
        E &= valid64;                   // positions at/after the block end: E = 0
        if (pos == blk_lo) E &= ~1ull;  // the block start is a head
        {
            const uint32_t up = __shfl_up_sync(FULL, (uint32_t)(E >> 62), 1);
            if (lane != 0) e_before = up;
        }
        uint32_t e_after = __shfl_down_sync(FULL, (uint32_t)E & 7u, 1);  // bit i = E[p0+64+i]
        if (lane == 31) {
            e_after = 0;
            if (nvalid == RLE_PER_THREAD && pos + RLE_PER_THREAD < blk_hi) {
                const uint32_t w = *reinterpret_cast<const uint32_t *>(in + pos + RLE_PER_THREAD);  // next 4 bytes
                const uint32_t c0 = w & 0xffu, c1 = (w >> 8) & 0xffu, c2 = (w >> 16) & 0xffu;
                const uint64_t left = blk_hi - (pos + RLE_PER_THREAD);  // valid bytes from there on
                e_after = (c0 == (v[RLE_VECS - 1].w >> 24)) ? 1u : 0u;
                if (left > 1 && c1 == c0) e_after |= 2u;
                if (left > 2 && c2 == c1) e_after |= 4u;
            }
        }
        const uint64_t my_heads = ~E;  // includes every position >= valid
        sh.heads[2 * tid] = (uint32_t)my_heads;
        sh.heads[2 * tid + 1] = (uint32_t)(my_heads >> 32);
        if (tid < RLE_HALO / 16) {  // look-ahead halo heads, straight from global memory
            const uint32_t q0 = RLE_TILE + tid * 16;
            uint32_t hm = 0xffffu;
            if (q0 < valid_len) {
                const uint4 hv = rle_ld16(in + tile_lo + q0);
                const uint32_t pb = in[tile_lo + q0 - 1];
                hm = ~eq_prev_mask16(hv, pb) & 0xffffu;
                if (q0 + 16 > valid_len) hm |= (0xffffu << (valid_len - q0)) & 0xffffu;
            }
            reinterpret_cast<uint16_t *>(sh.heads)[RLE_TILE / 16 + tid] = (uint16_t)hm;
        }
        // ---- last head strictly before each thread's first position ----
        int32_t lh = my_heads ? (int32_t)(p0 + 63 - __clzll((long long)my_heads)) : -1;  // highest head inside my 64
        {
            int32_t incl = lh;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int32_t o = __shfl_up_sync(FULL, incl, d);
                if (lane >= (unsigned)d && o > incl) incl = o;
            }
            if (lane == 31) sh.lasthead[warp] = incl;
            int32_t prev = __shfl_up_sync(FULL, incl, 1);
            if (lane == 0) prev = -1;
            lh = prev;
        }
        __syncthreads();  // heads[] and lasthead[] complete
        {
            int32_t before = -1;
            for (unsigned w = 0; w < warp; w++) {
                const int32_t o = sh.lasthead[w];
                if (o > before) before = o;
            }
            if (before > lh) lh = before;
        }

        // ---- 2. chain #1: equal-byte run still open at the tile start ----
        if (warp == 0) {
            int32_t lastv = -1;  // last head among the tile's valid positions
            for (int w = (int)((tile_len + 31) / 32) - 1 - (int)lane; w >= 0; w -= 32) {
                uint32_t word = sh.heads[w];
                const uint32_t lim = tile_len - (uint32_t)w * 32;
                if (lim < 32) word &= (1u << lim) - 1u;
                if (word) {
                    lastv = w * 32 + 31 - clz32(word);
                    break;
                }
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                const int32_t o = __shfl_xor_sync(FULL, lastv, d);
                if (o > lastv) lastv = o;
            }
            const bool head0 = (sh.heads[0] & 1u) != 0;
            uint64_t run_in = 0;
            if (lastv >= 0) {
                // the run open at the tile end starts inside this tile: publish it right away
                if (lane == 0) desc_store(desc_run + t, desc_pack(DESC_PREFIX, (uint64_t)(tile_len - (uint32_t)lastv)));
                if (!head0) run_in = lookback_read(desc_run, (int64_t)t, (int64_t)first_tile_of_block);
            } else {
                // no head in the whole tile: the run passes through; aggregate = tile_len
                run_in = lookback_sum(desc_run, (int64_t)t, (int64_t)first_tile_of_block, tile_len);
            }
            if (lane == 0) sh.run_in = run_in;
        }
        __syncthreads();
        const uint64_t run_in = sh.run_in;

        // ---- 3. what does each position emit ----
        uint64_t lit = 0, tri = 0;
        if (nvalid) {
            const uint64_t k0 = (lh >= 0) ? (uint64_t)(p0 - (uint32_t)lh) : (uint64_t)p0 + run_in;  // k at p0 if p0 is no head
            // deep inside a long run the first position of every further 255-byte chunk acts as a head
            uint64_t Ev = E;
            uint32_t bv = e_before, av = e_after;
            if ((E & 1ull) && k0 + (RLE_PER_THREAD + 2) >= 255) rle_chunk_heads(k0, &Ev, &bv, &av);
            const RleMasks64 m = rle_fast_masks64(Ev, bv, av, Z & valid64, valid64);
            lit = m.lit;
            tri = m.tri;
        }
        const uint32_t my_bytes = (uint32_t)__popcll(lit) + 3u * (uint32_t)__popcll(tri);

        // ---- 4. offsets: publish the tile's size at once, resolve its place only after the staging ----
        uint32_t tile_bytes;
        const uint32_t excl = block_excl_scan<RLE_THREADS>(my_bytes, sh.scan, &tile_bytes);
        if (tid == 0) lookback_publish(desc_out, (int64_t)t, 0, tile_bytes);
        if (!WRITE && !HIST) {
            if (warp == 0 && last_tile) {
                const uint64_t pfx = lookback_resolve(desc_out, (int64_t)t, 0, tile_bytes);
                if (lane == 0) blk_end[b] = pfx + tile_bytes;
            } else if (warp == 0) {
                lookback_resolve(desc_out, (int64_t)t, 0, tile_bytes);
            }
            continue;
        }

        // ---- 5. stage at tile-relative offsets (needs no prefix: the look-back latency hides behind it) ----
        {
            // literals: one uniform, fully predicated pass with a running offset; a triple just reserves 3 bytes
            uint8_t *const dst0 = sh.stage + excl;
            uint32_t o_s = smem_opaque(sh.stage) + excl;
#pragma unroll
            for (int c = 0; c < RLE_VECS; c++) {
                const uint32_t l16 = (uint32_t)(lit >> (16 * c)) & 0xffffu, t16 = (uint32_t)(tri >> (16 * c)) & 0xffffu;
                if (__all_sync(__activemask(), t16 == 0 && l16 == 0xffffu)) {  // whole warp: sixteen literals each
#pragma unroll
                    for (int j = 0; j < 16; j++) sts_u8(o_s + j, byte_of(v[c], j));
                    o_s += 16;
                } else if (__any_sync(__activemask(), __popc(l16) > RLE_SPARSE_MAX)) {
                    rle_lit_steps<0>(o_s, v[c], l16, t16);  // dense: sixteen predicated steps, whatever the masks are
                } else {
                    // sparse (run-heavy data: a vector emits a triple or two and hardly any literal): one iteration per
                    // literal instead of sixteen predicated steps per vector -- nothing at all for a vector without literals
                    uint32_t lb = l16;
                    while (lb) {
                        const uint32_t j = (uint32_t)ctz32(lb);
                        lb &= lb - 1u;
                        const uint32_t below = (1u << j) - 1u;
                        sts_u8(o_s + (uint32_t)__popc(l16 & below) + 3u * (uint32_t)__popc(t16 & below), in[pos + 16 * c + j]);  // L1 hit
                    }
                    o_s += (uint32_t)__popc(l16) + 3u * (uint32_t)__popc(t16);
                }
            }
            // triples: sparse loop over their positions (the count needs a look-ahead in the head bitmask)
            uint64_t tb = tri;
#pragma unroll 1
            while (tb) {
                const int q = ctz64(tb);
                tb &= tb - 1;
                const uint64_t below = (1ull << q) - 1ull;
                uint8_t *d = dst0 + (uint32_t)__popcll(lit & below) + 3u * (uint32_t)__popcll(tri & below);
                d[0] = 0;
                d[1] = in[pos + q];  // L1 hit: the line was just loaded
                d[2] = (uint8_t)rle_run_remaining(sh.heads, p0 + (uint32_t)q);
            }
        }
        if (warp == 0) {
            const uint64_t pfx = lookback_resolve(desc_out, (int64_t)t, 0, tile_bytes);
            if (lane == 0) sh.out_prefix = pfx;
        }
        __syncthreads();
        const uint64_t out0 = sh.out_prefix;
        if (last_tile && tid == 0) blk_end[b] = out0 + tile_bytes;

        // ---- 6. copy out + fused histogram ----
        if (WRITE) {
            if (out0 + tile_bytes > out_capacity) {
                if (tid == 0) atomicOr(err, (uint32_t)ERR_CAPACITY);
            } else {
                copy_out_unaligned<RLE_THREADS>(out + out0, sh.stage, tile_bytes);
            }
        }
        if (HIST) {
            // bin address = this lane's sub-histogram + byte * 32: byte extract, multiply-add, reduction
            const uint32_t mine_s = smem_opaque(sh.hist) + (lane & (RLE_HIST_COPIES - 1)) * 4u;
            auto bump = [&](uint32_t byte) {
                asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(mine_s + byte * (RLE_HIST_COPIES * 4u)), "r"(1u) : "memory");
            };
            const uint32_t *stage32 = reinterpret_cast<const uint32_t *>(sh.stage);
            const uint32_t whole = tile_bytes >> 2;
            for (uint32_t w = tid; w < whole; w += RLE_THREADS) {
                const uint32_t word = stage32[w];
                bump(__byte_perm(word, 0, 0x4440));
                bump(__byte_perm(word, 0, 0x4441));
                bump(__byte_perm(word, 0, 0x4442));
                bump(__byte_perm(word, 0, 0x4443));
            }
            if (tid < (tile_bytes & 3u)) bump(sh.stage[whole * 4 + tid]);
            __syncthreads();
            uint32_t sum = 0;
#pragma unroll
            for (int c = 0; c < RLE_HIST_COPIES; c++) sum += sh.hist[tid * RLE_HIST_COPIES + c];
            if (sum) atomicAdd(hist + b * 256 + tid, (unsigned long long)sum);
        }
    }
}

int launch_rle_encode(shafa_cuda_ctx *ctx, const uint8_t *d_in, uint64_t n, uint64_t block_size, uint8_t *d_rle,
                      uint64_t rle_capacity, uint64_t *h_rle_sizes, uint64_t *d_hist, uint32_t *d_err)
{
    if (n == 0) return SHAFA_SUCCESS;
    if (block_size == 0 || (block_size & 15)) return set_error(ctx, SHAFA_FILE_STREAM_FAILED, "rle_encode: block_size must be a positive multiple of 16");
    if (((uintptr_t)d_in & 15) || ((uintptr_t)d_rle & 15)) return set_error(ctx, SHAFA_FILE_STREAM_FAILED, "rle_encode: buffers must be 16-byte aligned");
    const uint64_t nb = (n + block_size - 1) / block_size;
    const uint64_t tpb = (block_size + RLE_TILE - 1) / RLE_TILE;
    const uint64_t last_len = n - (nb - 1) * block_size;
    const uint64_t total_tiles = (nb - 1) * tpb + (last_len + RLE_TILE - 1) / RLE_TILE;
    if (tpb > 0xffffffffull || total_tiles > 0xfffffff0ull) return set_error(ctx, SHAFA_LACK_OF_MEMORY, "rle_encode: too many tiles");

    int *occ = ctx->occ + OCC_RLE_ENC;
    const size_t smem = sizeof(RleEncShared);
    const int variant = (d_rle ? 2 : 0) | (d_hist ? 1 : 0);
    void (*kern)(const uint8_t *, uint64_t, uint64_t, uint32_t, uint64_t, uint8_t *, uint64_t, unsigned long long *,
                 uint64_t *, uint64_t *, uint64_t *, uint32_t *, uint32_t *) =
        variant == 3 ? k_rle_encode<true, true>
        : variant == 2 ? k_rle_encode<true, false>
        : variant == 1 ? k_rle_encode<false, true>
                       : k_rle_encode<false, false>;
    if (!occ[variant]) {
        SHAFA_CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SHAFA_CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[variant], kern, RLE_THREADS, smem));
        if (occ[variant] < 1) occ[variant] = 1;
    }
    uint64_t *desc = (uint64_t *)scratch_get(ctx, SCR_DESC, 2 * (total_tiles + 1) * sizeof(uint64_t));
    uint64_t *blk_end = (uint64_t *)scratch_get(ctx, SCR_SIZES, nb * sizeof(uint64_t));
    uint32_t *counter = misc_ptr(ctx, MISC_COUNTER);
    if (!desc || !blk_end || !counter) return SHAFA_LACK_OF_MEMORY;
    SHAFA_CUDA_TRY(ctx, cudaMemsetAsync(desc, 0, 2 * (total_tiles + 1) * sizeof(uint64_t), ctx->stream));
    SHAFA_CUDA_TRY(ctx, cudaMemsetAsync(counter, 0, 4, ctx->stream));
    if (d_hist) SHAFA_CUDA_TRY(ctx, cudaMemsetAsync(d_hist, 0, nb * 256 * sizeof(uint64_t), ctx->stream));
    uint64_t grid = (uint64_t)ctx->sm_count * occ[variant];
    if (grid > total_tiles) grid = total_tiles;
    SHAFA_LAUNCH(ctx, d_rle ? "rle_encode" : "rle_encode_sizes", kern, dim3((unsigned)grid), dim3(RLE_THREADS), smem,
                 d_in, n, block_size, (uint32_t)tpb, total_tiles, d_rle, rle_capacity, (unsigned long long *)d_hist,
                 blk_end, desc, desc + total_tiles + 1, counter, d_err);
    SHAFA_CHECK_LAUNCH(ctx);
    // sizes back to the host (this is the one synchronisation of the call)
    uint64_t *h_end = (uint64_t *)pinned_get(ctx, nb * sizeof(uint64_t));
    if (!h_end) return SHAFA_LACK_OF_MEMORY;
    SHAFA_CUDA_TRY(ctx, cudaMemcpyAsync(h_end, blk_end, nb * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    SHAFA_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    uint64_t prev = 0;
    for (uint64_t b = 0; b < nb; b++) {
        h_rle_sizes[b] = h_end[b] - prev;
        prev = h_end[b];
    }
    return SHAFA_SUCCESS;
}

}  // namespace shafa
