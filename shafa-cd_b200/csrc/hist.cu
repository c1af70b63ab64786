// hist.cu -- per-block 256-bin byte histograms (reference make_freq, f.c:63-79).
//
// HBM-bound: every input byte is read once (algorithmic bytes = N), the output
// is 2 KiB per block.  Tiles are 16 KiB windows aligned in the stream's address
// space, intersected with the block's byte range, so all loads are aligned
// 16-byte vectors even for ragged blocks (the .rle stream).  Counts are
// privatised per warp in shared memory and flushed with one 64-bit atomic per
// non-empty bin per tile.
#include "common.cuh"
#include "internal.h"

namespace shafa {

constexpr int HIST_THREADS = 256;
constexpr int HIST_VECS = 4;                               // 16-byte vectors per thread
constexpr int HIST_TILE = HIST_THREADS * 16 * HIST_VECS;   // 16384 bytes
constexpr int HIST_WARPS = HIST_THREADS / 32;

__global__ void __launch_bounds__(HIST_THREADS)
k_histogram(const uint8_t *__restrict__ in, const uint64_t *__restrict__ off, const uint64_t *__restrict__ tile0,
            uint32_t nb, unsigned long long *__restrict__ hist)
{
    __shared__ uint32_t sh[HIST_WARPS][256];
    __shared__ uint32_t s_block;
    const unsigned tid = threadIdx.x;
    for (int i = tid; i < HIST_WARPS * 256; i += HIST_THREADS) (&sh[0][0])[i] = 0;
    const uint64_t t = blockIdx.x;
    if (tid == 0) s_block = find_block(tile0, nb, t);
    __syncthreads();
    const uint32_t b = s_block;
    const uint64_t lo = off[b], hi = off[b + 1];
    const uint64_t base = (lo / HIST_TILE + (t - tile0[b])) * HIST_TILE;
    uint32_t *mine = sh[warp_id()];

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
            for (int j = 0; j < 16; j++) atomicAdd(&mine[byte_of(v[c], j)], 1u);
        } else if (pos[c] + 16 > lo && pos[c] < hi) {
#pragma unroll
            for (int j = 0; j < 16; j++) {
                const uint64_t p = pos[c] + j;
                if (p >= lo && p < hi) atomicAdd(&mine[byte_of(v[c], j)], 1u);
            }
        }
    }
    __syncthreads();
    uint32_t sum = 0;
#pragma unroll
    for (int w = 0; w < HIST_WARPS; w++) sum += sh[w][tid];
    if (sum) atomicAdd(hist + (uint64_t)b * 256 + tid, (unsigned long long)sum);
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
    SHAFA_LAUNCH(ctx, "histogram", k_histogram, dim3((unsigned)tiles), dim3(HIST_THREADS), 0, d_in, dp, dp + nb + 1,
                 (uint32_t)nb, (unsigned long long *)d_hist);
    SHAFA_CHECK_LAUNCH(ctx);
    return SHAFA_SUCCESS;
}

}  // namespace shafa
