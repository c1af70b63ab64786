// hist.cu -- per-block 256-bin byte histograms (reference make_freq, f.c:63-79).
//
// HBM-bound in principle: every input byte is read once (algorithmic bytes = N), the
// output is 2 KiB per block.  Tiles are 16 KiB windows aligned in the stream's address
// space, intersected with the block's byte range, so all loads are aligned 16-byte
// vectors even for ragged blocks (the .rle stream).
//
// Persistent CTAs, each owning a contiguous chunk of tiles (no chain: counts commute).
// Counts live in shared memory as 32 copies per bin, bins[sym * 32 + lane]: lane l of
// every warp only touches bank l, so a warp's 32 increments never collide on a bank
// or on an address however skewed the data is; the eight warps meet through
// shared-memory reductions.  A CTA adds its copies up and issues one 64-bit global
// atomic per non-empty bin only when its chunk crosses into another block (and at
// the end), instead of once per tile.
#include "common.cuh"
#include "internal.h"

namespace shafa {

constexpr int HIST_THREADS = 256;
constexpr int HIST_VECS = 4;                               // 16-byte vectors per thread
constexpr int HIST_TILE = HIST_THREADS * 16 * HIST_VECS;   // 16384 bytes
constexpr int HIST_COPIES = 32;
constexpr uint32_t HIST_FLUSH_TILES = 1u << 17;            // 2 GiB: a 32-bit copy cannot overflow before this

__device__ __forceinline__ void hist_inc(uint32_t saddr)
{
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(saddr), "r"(1u) : "memory");
}

// add the CTA's copies of every bin into block b's histogram and clear them (all threads)
__device__ __forceinline__ void hist_flush(uint32_t *bins, unsigned long long *hist, uint32_t b)
{
    __syncthreads();
    const unsigned tid = threadIdx.x;
    uint32_t sum = 0;
#pragma unroll
    for (int k = 0; k < HIST_COPIES; k++) {
        const uint32_t idx = tid * HIST_COPIES + ((tid + k) & (HIST_COPIES - 1));  // rotated: one bank per lane
        sum += bins[idx];
        bins[idx] = 0;
    }
    if (sum) atomicAdd(hist + (uint64_t)b * 256 + tid, (unsigned long long)sum);
    __syncthreads();
}

__global__ void __launch_bounds__(HIST_THREADS, 4)
k_histogram(const uint8_t *__restrict__ in, const uint64_t *__restrict__ off, const uint64_t *__restrict__ tile0,
            uint32_t nb, uint64_t total_tiles, unsigned long long *__restrict__ hist)
{
    __shared__ uint32_t bins[256 * HIST_COPIES];
    const unsigned tid = threadIdx.x;
    for (int i = tid; i < 256 * HIST_COPIES; i += HIST_THREADS) bins[i] = 0;
    const uint64_t per = (total_tiles + gridDim.x - 1) / gridDim.x;
    const uint64_t t_first = (uint64_t)blockIdx.x * per;
    const uint64_t t_last = (t_first + per < total_tiles) ? (t_first + per) : total_tiles;
    if (t_first >= t_last) return;
    __syncthreads();
    uint32_t lane_s = smem_u32(bins) + lane_id() * 4u;  // this lane's column
    asm volatile("" : "+r"(lane_s));

    uint32_t b = find_block(tile0, nb, t_first);
    uint64_t b_first = tile0[b], b_next = tile0[b + 1];
    uint64_t lo = off[b], hi = off[b + 1];
    uint32_t since_flush = 0;
    for (uint64_t t = t_first; t < t_last; t++) {
        if (t >= b_next || since_flush >= HIST_FLUSH_TILES) {
            hist_flush(bins, hist, b);
            since_flush = 0;
            while (t >= b_next) {  // skips empty blocks
                b++;
                b_first = b_next;
                b_next = tile0[b + 1];
            }
            lo = off[b];
            hi = off[b + 1];
        }
        since_flush++;
        const uint64_t base = (lo / HIST_TILE + (t - b_first)) * HIST_TILE;
        uint4 v[HIST_VECS];
        uint64_t pos[HIST_VECS];
#pragma unroll
        for (int c = 0; c < HIST_VECS; c++) {
            pos[c] = base + (uint64_t)(c * HIST_THREADS + tid) * 16;
            v[c] = make_uint4(0, 0, 0, 0);
            if (pos[c] + 16 > lo && pos[c] < hi) v[c] = ld_stream16(in + pos[c]);
        }
#pragma unroll
        for (int c = 0; c < HIST_VECS; c++) {
            if (pos[c] >= lo && pos[c] + 16 <= hi) {
#pragma unroll
                for (int j = 0; j < 16; j++) hist_inc(lane_s + byte_of(v[c], j) * (HIST_COPIES * 4));
            } else if (pos[c] + 16 > lo && pos[c] < hi) {
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    const uint64_t p = pos[c] + j;
                    if (p >= lo && p < hi) hist_inc(lane_s + byte_of(v[c], j) * (HIST_COPIES * 4));
                }
            }
        }
    }
    hist_flush(bins, hist, b);
}

int launch_histogram(shafa_cuda_ctx *ctx, const uint8_t *d_in, const uint64_t *h_off, uint64_t nb, uint64_t *d_hist)
{
    if (nb == 0) return SHAFA_SUCCESS;
    if (((uintptr_t)d_in & 15) != 0) return set_error(ctx, SHAFA_FILE_STREAM_FAILED, "histogram: d_in not 16-byte aligned");
    // plan: [off (nb+1)] [tile0 (nb+1)]
    uint64_t *hp = (uint64_t *)pinned_get(ctx, (nb + 1) * 2 * sizeof(uint64_t));
    uint64_t *dp = (uint64_t *)scratch_get(ctx, SCR_PLAN, (nb + 1) * 2 * sizeof(uint64_t));
    if (!hp || !dp) return SHAFA_LACK_OF_MEMORY;
    uint64_t tiles = 0;
    for (uint64_t b = 0; b < nb; b++) {
        hp[b] = h_off[b];
        hp[nb + 1 + b] = tiles;
        const uint64_t lo = h_off[b], hi = h_off[b + 1];
        if (hi < lo) return set_error(ctx, SHAFA_FILE_STREAM_FAILED, "histogram: offsets not monotone");
        if (hi > lo) tiles += (hi - 1) / HIST_TILE - lo / HIST_TILE + 1;
    }
    hp[nb] = h_off[nb];
    hp[2 * nb + 1] = tiles;
    SHAFA_CUDA_TRY(ctx, cudaMemcpyAsync(dp, hp, (nb + 1) * 2 * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    SHAFA_CUDA_TRY(ctx, cudaMemsetAsync(d_hist, 0, nb * 256 * sizeof(uint64_t), ctx->stream));
    if (tiles == 0) return SHAFA_SUCCESS;
    if (tiles > 0x7fffffffull) return set_error(ctx, SHAFA_LACK_OF_MEMORY, "histogram: too many tiles");
    int &occ = ctx->occ[OCC_HIST];
    if (!occ) {
        SHAFA_CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_histogram, HIST_THREADS, 0));
        if (occ < 1) occ = 1;
    }
    uint64_t grid = (uint64_t)ctx->sm_count * occ;
    if (grid > tiles) grid = tiles;
    SHAFA_LAUNCH(ctx, "histogram", k_histogram, dim3((unsigned)grid), dim3(HIST_THREADS), 0, d_in, dp, dp + nb + 1,
                 (uint32_t)nb, tiles, (unsigned long long *)d_hist);
    SHAFA_CHECK_LAUNCH(ctx);
    return SHAFA_SUCCESS;
}

}  // namespace shafa
