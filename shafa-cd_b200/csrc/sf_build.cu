// sf_build.cu -- Shannon-Fano code construction, one CTA per block
// (reference insert_sort / not_null / best_Division / sf_codes, t.c:74-210).
//
// 256 threads.  (1) stable descending rank of every symbol by counting
// (ties -> lower symbol first, t.c:87); (2) prefix sums over the rank order;
// (3) the reference's recursion is replaced by level-synchronous splitting:
// every rank keeps its current interval [lo,hi], all ranks of an interval
// compute the same split with shafa::sf_split (binary search on the prefix sums)
// and append '0' (left) or '1' (right), until every interval is a single rank.
// Codes never leave the device: module C reads d_len/d_code directly.
#include "common.cuh"
#include "core.h"
#include "internal.h"

namespace shafa {

__global__ void __launch_bounds__(256)
k_sf_build(const uint64_t *__restrict__ hist, uint8_t *__restrict__ len_out, uint64_t *__restrict__ code_out,
           uint32_t *__restrict__ err)
{
    __shared__ uint32_t cnt[256];      // by symbol
    __shared__ uint32_t sorted[256];   // by rank
    __shared__ uint32_t symr[256];     // rank -> symbol
    __shared__ uint32_t P[257];        // P[r+1] = sum of ranks 0..r
    __shared__ uint32_t scan[9];
    __shared__ unsigned long long s_total;
    const unsigned s = threadIdx.x;
    const uint64_t b = blockIdx.x;
    const uint64_t h = hist[b * 256 + s];
    if (s == 0) s_total = 0;
    __syncthreads();
    {
        const uint64_t ws = warp_sum64(h);
        if (lane_id() == 0) atomicAdd(&s_total, (unsigned long long)ws);
    }
    const int bad_count = __syncthreads_or(h >= (1ull << 31));
    // t.c:133-141 does this arithmetic in int: totals >= 2^31 are undefined there
    if (bad_count || s_total >= (1ull << 31)) {
        if (s == 0) atomicOr(err, 1u);
        len_out[b * 256 + s] = 0;
        code_out[b * 256 + s] = 0;
        return;
    }
    const uint32_t my = (uint32_t)h;
    cnt[s] = my;
    __syncthreads();
    uint32_t rank = 0;
#pragma unroll 8
    for (int t = 0; t < 256; t++) {
        const uint32_t ft = cnt[t];
        rank += (ft > my) || (ft == my && (unsigned)t < s);
    }
    sorted[rank] = my;
    symr[rank] = s;
    __syncthreads();
    // from here on thread r stands for rank r
    const int r = (int)s;
    const uint32_t v = sorted[r];
    uint32_t total;
    const uint32_t excl = block_excl_scan<256>(v, scan, &total);
    P[r + 1] = excl + v;
    if (r == 0) P[0] = 0;
    const int nn = __syncthreads_count(v != 0) - 1;  // t.c:202-210 last non-zero rank
    const bool active = r <= nn;
    int lo = 0, hi = nn;
    uint64_t code = 0;
    uint32_t len = 0;
    for (;;) {
        const bool split = active && lo != hi;  // t.c:185
        if (!__syncthreads_or(split)) break;
        if (split) {
            const int d = sf_split(P, lo, hi);
            if (r <= d) {
                hi = d;
                code <<= 1;  // t.c:189
            } else {
                lo = d + 1;
                code = (code << 1) | 1ull;  // t.c:190
            }
            len++;
        }
    }
    if (len > 64) {
        atomicOr(err, 2u);
        len = 0;
    }
    const uint32_t sym = symr[r];
    len_out[b * 256 + sym] = active ? (uint8_t)len : 0;
    code_out[b * 256 + sym] = active ? code : 0ull;
}

int launch_sf_build(shafa_cuda_ctx *ctx, const uint64_t *d_hist, uint64_t nb, uint8_t *d_len, uint64_t *d_code,
                    uint32_t *d_err)
{
    if (nb == 0) return SHAFA_SUCCESS;
    SHAFA_LAUNCH(ctx, "sf_build", k_sf_build, dim3((unsigned)nb), dim3(256), 0, d_hist, d_len, d_code, d_err);
    SHAFA_CHECK_LAUNCH(ctx);
    return SHAFA_SUCCESS;
}

}  // namespace shafa
