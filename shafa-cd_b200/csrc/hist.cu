// hist.cu -- per-block 256-bin byte histograms (reference make_freq, f.c:63-79).
//
// HBM-bound in principle: every input byte is read once (algorithmic bytes = N), the
// output is 2 KiB per block.  Tiles are 16 KiB windows aligned in the stream's address
// space, intersected with the block's byte range, so all loads are aligned 16-byte
// vectors even for ragged blocks (the .rle stream).
//
// Persistent CTAs, each owning a contiguous chunk of tiles (no chain: counts commute).  The
// input is staged through shared memory by bulk async copies (cp.async.bulk, "TMA" 1-D) on a
// ring of three 16 KiB buffers with one mbarrier each: the copies of the next tiles are in
// flight while a tile is counted, and no load instruction or register is spent on the way in.
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

// ---- bulk async copy (TMA 1-D) + mbarrier ----------------------------------------------------
__device__ __forceinline__ void hist_mbar_init(uint64_t *bar)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void hist_bulk_load(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar)
{
    const uint32_t b = smem_u32(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(b)
                 : "memory");
}
__device__ __forceinline__ void hist_mbar_wait(uint64_t *bar, uint32_t parity)
{
    const uint32_t b = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(b), "r"(parity)
                     : "memory");
    } while (!done);
}

constexpr int HIST_STAGES = 3;  // 16 KiB tiles in flight per CTA (bulk copies land while earlier tiles are counted)

struct HistShared {
    alignas(16) uint8_t buf[HIST_STAGES][HIST_TILE];
    uint32_t bins[256 * HIST_COPIES];
    alignas(8) uint64_t full[HIST_STAGES];
};

// where tile t of the plan lies: incremental cursor over the blocks (a CTA walks its chunk in order)
struct HistCursor {
    uint32_t b;
    uint64_t b_first, b_next, lo, hi;
    __device__ __forceinline__ void init(const uint64_t *off, const uint64_t *tile0, uint32_t nb, uint64_t t)
    {
        b = find_block(tile0, nb, t);
        b_first = tile0[b];
        b_next = tile0[b + 1];
        lo = off[b];
        hi = off[b + 1];
    }
    // returns true if the block changed
    __device__ __forceinline__ bool advance_to(const uint64_t *off, const uint64_t *tile0, uint64_t t)
    {
        if (t < b_next) return false;
        while (t >= b_next) {  // skips empty blocks
            b++;
            b_first = b_next;
            b_next = tile0[b + 1];
        }
        lo = off[b];
        hi = off[b + 1];
        return true;
    }
    __device__ __forceinline__ uint64_t base(uint64_t t) const { return (lo / HIST_TILE + (t - b_first)) * HIST_TILE; }
};

// Persistent CTAs over contiguous chunks of tiles.  Thread 0 keeps HIST_STAGES bulk async copies (cp.async.bulk,
// 16 KiB each, completion on an mbarrier) ahead of the counting, so the input never passes through registers on its
// way in and the HBM pipe stays full while the shared-memory reductions run.
__global__ void __launch_bounds__(HIST_THREADS, 2)
k_histogram(const uint8_t *__restrict__ in, const uint64_t *__restrict__ off, const uint64_t *__restrict__ tile0,
            uint32_t nb, uint64_t total_tiles, unsigned long long *__restrict__ hist)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    HistShared &sh = *reinterpret_cast<HistShared *>(smem_raw);
    uint32_t *bins = sh.bins;
    const unsigned tid = threadIdx.x;
    for (int i = tid; i < 256 * HIST_COPIES; i += HIST_THREADS) bins[i] = 0;
    const uint64_t per = (total_tiles + gridDim.x - 1) / gridDim.x;
    const uint64_t t_first = (uint64_t)blockIdx.x * per;
    const uint64_t t_last = (t_first + per < total_tiles) ? (t_first + per) : total_tiles;
    if (t_first >= t_last) return;
    const uint64_t readable = (off[nb] + 15) & ~15ull;  // the stream may be read up to its next 16-byte boundary
    HistCursor prod;                                    // thread 0: the tile whose copy is issued next
    uint64_t t_issue = t_first;
    auto issue = [&](int slot) {                        // thread 0 only
        prod.advance_to(off, tile0, t_issue);
        const uint64_t base = prod.base(t_issue);
        const uint64_t left = readable - base;
        hist_bulk_load(sh.buf[slot], in + base, left < (uint64_t)HIST_TILE ? (uint32_t)left : (uint32_t)HIST_TILE, &sh.full[slot]);
        t_issue++;
    };
    if (tid == 0) {
        for (int s = 0; s < HIST_STAGES; s++) hist_mbar_init(&sh.full[s]);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        prod.init(off, tile0, nb, t_first);
        for (int s = 0; s < HIST_STAGES && t_issue < t_last; s++) issue(s);
    }
    __syncthreads();
    uint32_t lane_s = smem_u32(bins) + lane_id() * 4u;  // this lane's column
    asm volatile("" : "+r"(lane_s));

    HistCursor cur;
    cur.init(off, tile0, nb, t_first);
    uint32_t since_flush = 0, slot = 0, parity = 0;
    for (uint64_t t = t_first; t < t_last; t++) {
        if (t >= cur.b_next || since_flush >= HIST_FLUSH_TILES) {
            hist_flush(bins, hist, cur.b);
            since_flush = 0;
            cur.advance_to(off, tile0, t);
        }
        since_flush++;
        const uint64_t base = cur.base(t), lo = cur.lo, hi = cur.hi;
        hist_mbar_wait(&sh.full[slot], parity);
        const uint4 *tile = reinterpret_cast<const uint4 *>(sh.buf[slot]);
#pragma unroll
        for (int c = 0; c < HIST_VECS; c++) {
            const uint32_t k = c * HIST_THREADS + tid;  // consecutive lanes, consecutive 16-byte vectors: no bank conflicts
            const uint64_t pos = base + (uint64_t)k * 16;
            if (pos >= lo && pos + 16 <= hi) {
                const uint4 v = tile[k];
#pragma unroll
                for (int j = 0; j < 16; j++) hist_inc(lane_s + byte_of(v, j) * (HIST_COPIES * 4));
            } else if (pos + 16 > lo && pos < hi) {  // block edges
                const uint4 v = tile[k];
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    const uint64_t q = pos + j;
                    if (q >= lo && q < hi) hist_inc(lane_s + byte_of(v, j) * (HIST_COPIES * 4));
                }
            }
        }
        __syncthreads();  // everybody has read the slot: it can be refilled
        if (tid == 0 && t_issue < t_last) issue((int)slot);
        if (++slot == HIST_STAGES) {
            slot = 0;
            parity ^= 1u;
        }
    }
    hist_flush(bins, hist, cur.b);
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
    const size_t smem = sizeof(HistShared);
    if (!occ) {
        SHAFA_CUDA_TRY(ctx, cudaFuncSetAttribute(k_histogram, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SHAFA_CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_histogram, HIST_THREADS, smem));
        if (occ < 1) occ = 1;
    }
    uint64_t grid = (uint64_t)ctx->sm_count * occ;
    if (grid > tiles) grid = tiles;
    SHAFA_LAUNCH(ctx, "histogram", k_histogram, dim3((unsigned)grid), dim3(HIST_THREADS), smem, d_in, dp, dp + nb + 1,
                 (uint32_t)nb, tiles, (unsigned long long *)d_hist);
    SHAFA_CHECK_LAUNCH(ctx);
    return SHAFA_SUCCESS;
}

}  // namespace shafa
