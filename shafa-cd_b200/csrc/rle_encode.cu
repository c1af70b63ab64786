// rle_encode.cu -- Shafa's escape/run RLE (reference block_compression, f.c:29-55)
// for all blocks of a file in ONE pass, with the histogram of the emitted bytes
// fused (f.c:310) and the blocks' outputs compacted back to back (= the .rle
// file, f.c:307).  Algorithmic bytes: N read + R written.
//
// Tile = 4096 bytes of one block (tiles never straddle a block), 256 threads x
// 16 bytes.  Per tile:
//   1. 16-byte loads; run-head bitmask of the tile plus a 288-position look-ahead
//      halo (positions at/after the block end count as heads);
//   2. chain #1 (decoupled look-back inside the block): length of the equal-byte
//      run that is still open at the tile start, needed for k mod 255;
//   3. shafa::rle_emit_thread decides what each position emits (core.h);
//   4. CTA scan of emitted bytes, chain #2 (look-back over ALL tiles of the
//      file) gives the tile's offset in the compacted output;
//   5. bytes are staged in shared memory co-aligned with the destination and
//      copied out with 16-byte stores; the same staging buffer feeds the fused
//      256-bin histogram.
#include "common.cuh"
#include "core.h"
#include "internal.h"

namespace shafa {

constexpr int RLE_THREADS = RLE_TILE / RLE_ITEMS;  // 256
constexpr int RLE_STAGE = 2 * RLE_TILE + 3 + 16 + 13;  // worst case 2n+3 (f.c:244) + alignment slack, multiple of 16

struct RleEncShared {
    alignas(16) uint8_t bytes[16 + RLE_TILE + RLE_HALO + 16];  // [16 back halo][tile][look-ahead]
    alignas(16) uint8_t stage[RLE_STAGE];
    uint32_t heads[RLE_HEAD_WORDS + 2];
    uint32_t hist[RLE_THREADS / 32][256];
    uint32_t scan[RLE_THREADS / 32 + 1];
    int32_t lasthead[RLE_THREADS / 32];
    uint64_t out_prefix;
    uint64_t run_in;
    uint32_t ticket;
};

template <bool WRITE, bool HIST>
__global__ void __launch_bounds__(RLE_THREADS)
k_rle_encode(const uint8_t *__restrict__ in, uint64_t n, uint64_t block_size, uint32_t tiles_per_block,
             uint64_t total_tiles, uint8_t *__restrict__ out, uint64_t out_capacity,
             unsigned long long *__restrict__ hist, uint64_t *__restrict__ blk_end, uint64_t *__restrict__ desc_run,
             uint64_t *__restrict__ desc_out, uint32_t *__restrict__ counter, uint32_t *__restrict__ err)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RleEncShared &sh = *reinterpret_cast<RleEncShared *>(smem_raw);
    const unsigned tid = threadIdx.x, lane = lane_id(), warp = warp_id();
    uint8_t *tileb = sh.bytes + 16;  // tileb[p] = byte at tile position p; tileb[-1] = byte before the tile

    for (;;) {
        __syncthreads();
        if (tid == 0) sh.ticket = atomicAdd(counter, 1u);
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

        // ---- 1. load tile + halo (16-byte vectors; the block start is 16-byte aligned) ----
        uint4 v = make_uint4(0, 0, 0, 0);
        {
            const uint64_t pos = tile_lo + (uint64_t)tid * 16;
            if (pos < blk_hi) v = ld_stream16(in + pos);
            *reinterpret_cast<uint4 *>(tileb + tid * 16) = v;
            if (tid < RLE_HALO / 16) {  // look-ahead halo
                uint4 h = make_uint4(0, 0, 0, 0);
                const uint64_t hp = tile_lo + RLE_TILE + (uint64_t)tid * 16;
                if (hp < blk_hi) h = ld_stream16(in + hp);
                *reinterpret_cast<uint4 *>(tileb + RLE_TILE + tid * 16) = h;
            }
            if (tid == RLE_THREADS - 1) tileb[-1] = (ti != 0) ? in[tile_lo - 1] : 0;
        }
        __syncthreads();
        // ---- heads: bit p set iff position p starts a run (or lies at/after the block end) ----
        const uint32_t valid_len = (uint32_t)((blk_hi - tile_lo < (uint64_t)(RLE_TILE + RLE_HALO)) ? (blk_hi - tile_lo)
                                                                                                  : (RLE_TILE + RLE_HALO));
        uint32_t my_heads;
        {
            const uint32_t p0h = tid * 16;
            uint32_t m = ~eq_prev_mask16(v, tileb[(int)p0h - 1]) & 0xffffu;
            if (p0h + 16 > valid_len) m |= (p0h >= valid_len) ? 0xffffu : (0xffffu << (valid_len - p0h)) & 0xffffu;
            if (tid == 0 && ti == 0) m |= 1u;
            my_heads = m;
            reinterpret_cast<uint16_t *>(sh.heads)[tid] = (uint16_t)m;
            if (tid < RLE_HALO / 16) {  // halo chunks, from shared memory
                const uint32_t q0 = RLE_TILE + tid * 16;
                const uint4 hv = *reinterpret_cast<const uint4 *>(tileb + q0);
                uint32_t hm = ~eq_prev_mask16(hv, tileb[q0 - 1]) & 0xffffu;
                if (q0 + 16 > valid_len) hm |= (q0 >= valid_len) ? 0xffffu : (0xffffu << (valid_len - q0)) & 0xffffu;
                reinterpret_cast<uint16_t *>(sh.heads)[RLE_TILE / 16 + tid] = (uint16_t)hm;
            }
        }
        // ---- last head at or before each thread's first position ----
        int32_t lh = my_heads ? (int32_t)(tid * 16 + 31 - clz32(my_heads)) : -1;  // highest head inside my 16
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
            lh = prev;  // exclusive within the warp
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
            // run length that is open at the END of this tile, if it can be known locally
            // last head among valid positions:
            int32_t lastv = -1;
            for (int w = (int)((tile_len + 31) / 32) - 1 - (int)lane; w >= 0; w -= 32) {
                uint32_t word = sh.heads[w];
                const uint32_t lim = tile_len - (uint32_t)w * 32;  // valid bits in this word
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
                // the open run at the tile end starts inside this tile: publish it right away
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
        const uint32_t p0 = tid * 16;
        const int nvalid = (p0 >= tile_len) ? 0 : (int)((tile_len - p0 < 16) ? (tile_len - p0) : 16);
        uint32_t c0 = 0;
        if (nvalid) {
            const uint64_t k = (lh >= 0) ? (uint64_t)(p0 - (uint32_t)lh) : (uint64_t)p0 + run_in;
            c0 = (uint32_t)(k % 255);
        }
        auto symf = [&](int j) -> uint32_t { return byte_of(v, j); };
        const RleEmit em = rle_emit_thread(sh.heads, p0, nvalid, c0, symf);

        // ---- 4. offsets ----
        uint32_t tile_bytes;
        const uint32_t excl = block_excl_scan<RLE_THREADS>(em.bytes, sh.scan, &tile_bytes);
        if (warp == 0) {
            const uint64_t pfx = lookback_sum(desc_out, (int64_t)t, 0, tile_bytes);
            if (lane == 0) sh.out_prefix = pfx;
        }
        if (HIST) {
            for (int i = tid; i < (RLE_THREADS / 32) * 256; i += RLE_THREADS) (&sh.hist[0][0])[i] = 0;
        }
        __syncthreads();
        const uint64_t out0 = sh.out_prefix;
        if (last_tile && tid == 0) blk_end[b] = out0 + tile_bytes;

        // ---- 5. stage + copy out + fused histogram ----
        const uint32_t shift = (uint32_t)(out0 & 15u);  // co-align staging with the destination
        rle_write_thread(em, sh.stage + shift + excl, symf);
        __syncthreads();
        if (WRITE) {
            if (out0 + tile_bytes > out_capacity) {
                if (tid == 0) atomicOr(err, 32u);
            } else {
                uint8_t *dst = out + (out0 - shift);  // 16-byte aligned
                const uint32_t end = shift + tile_bytes;
                for (uint32_t c = tid * 16; c < end; c += RLE_THREADS * 16) {
                    if (c >= shift && c + 16 <= end) {
                        st_stream16(dst + c, *reinterpret_cast<const uint4 *>(sh.stage + c));
                    } else {
                        for (uint32_t k = (c > shift ? c : shift); k < c + 16 && k < end; k++) dst[k] = sh.stage[k];
                    }
                }
            }
        }
        if (HIST) {
            uint32_t *mine = sh.hist[warp];
            for (uint32_t k = shift + tid; k < shift + tile_bytes; k += RLE_THREADS) atomicAdd(&mine[sh.stage[k]], 1u);
            __syncthreads();
            uint32_t sum = 0;
#pragma unroll
            for (int w = 0; w < RLE_THREADS / 32; w++) sum += sh.hist[w][tid];
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

    static int occ[4] = {0, 0, 0, 0};
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
