// common.cuh -- device-side helpers shared by all kernels of libshafa_cuda.
// sm_100a only.  Nothing here is GEMM-shaped: these are HBM-bound byte/bit
// kernels, so the helpers are about coalesced 16-byte access, warp scans and the
// decoupled look-back chain (Merrill & Garland) used by every single-pass kernel.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace shafa {

constexpr unsigned FULL = 0xffffffffu;

// 32-bit shared-window address of a shared-memory object (operand of explicit ld/st/red.shared)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// Explicit shared-memory accesses through such addresses.  The compiler otherwise tends to rebuild the CTA's
// shared window base (S2R SR_CgaCtaId, MOV, LEA, IADD) in front of dynamically indexed or predicated accesses.
__device__ __forceinline__ uint32_t smem_opaque(const void *p)
{
    uint32_t a = smem_u32(p);
    asm volatile("" : "+r"(a));  // keeps the optimiser from rematerialising the address at every use
    return a;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t saddr)
{
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}
__device__ __forceinline__ void sts_u8(uint32_t saddr, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(saddr), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ unsigned warp_id() { return threadIdx.x >> 5; }

// ---- streaming 16-byte access -------------------------------------------------
__device__ __forceinline__ uint4 ld_stream16(const void *p)
{
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream16(void *p, uint4 v)
{
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
                 "r"(v.z), "r"(v.w)
                 : "memory");
}
__device__ __forceinline__ void st_stream4(void *p, uint32_t v)
{
    asm volatile("st.global.L1::no_allocate.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// byte j (0..15) of a uint4 in memory order
__device__ __forceinline__ uint32_t byte_of(const uint4 &v, int j)
{
    const uint32_t w = (j < 4) ? v.x : (j < 8) ? v.y : (j < 12) ? v.z : v.w;
    // one PRMT; as (w >> 8k) & 0xff the byte costs an instruction more wherever it is scaled into an address
    return __byte_perm(w, 0, 0x4440 | (j & 3));
}

// ---- look-back descriptors ------------------------------------------------------
// One 64-bit word per tile: [63:62] state, [61:0] value, so state and value are
// published by a single store and no fence is needed.
constexpr uint64_t DESC_INVALID = 0, DESC_AGG = 1, DESC_PREFIX = 2;
constexpr uint64_t DESC_VALUE_MASK = (1ull << 62) - 1;

__device__ __forceinline__ uint64_t desc_pack(uint64_t state, uint64_t value) { return (state << 62) | value; }
__device__ __forceinline__ uint64_t desc_state(uint64_t d) { return d >> 62; }
__device__ __forceinline__ uint64_t desc_value(uint64_t d) { return d & DESC_VALUE_MASK; }

__device__ __forceinline__ void desc_store(uint64_t *p, uint64_t v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t desc_load(const uint64_t *p)
{
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void st_relaxed_u32(uint32_t *p, uint32_t v)
{
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// ---- warp primitives ------------------------------------------------------------
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(FULL, v, d);
        if (lane_id() >= (unsigned)d) v += t;
    }
    return v;
}
__device__ __forceinline__ uint64_t warp_sum64(uint64_t v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
    return v;
}

// Exclusive scan of one uint32 per thread over a CTA of NT threads (NT multiple
// of 32, <= 1024).  `smem` needs NT/32 + 1 words.  Returns the exclusive prefix;
// *total receives the CTA total.  Contains two __syncthreads().
template <int NT>
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *smem, uint32_t *total)
{
    constexpr int NW = NT / 32;
    const uint32_t incl = warp_incl_scan(v);
    if (lane_id() == 31) smem[warp_id()] = incl;
    __syncthreads();
    if (warp_id() == 0) {
        uint32_t w = (lane_id() < NW) ? smem[lane_id()] : 0u;
        uint32_t wi = warp_incl_scan(w);
        if (lane_id() < NW) smem[lane_id()] = wi - w;
        if (lane_id() == NW - 1) smem[NW] = wi;
    }
    __syncthreads();
    *total = smem[NW];
    return smem[warp_id()] + incl - v;
}

// Decoupled look-back (Merrill & Garland) over a chain of tiles [first, tile).
// Called by all 32 lanes of ONE warp.  Walks the predecessors of `tile` back to
// the nearest one that has published an inclusive prefix and returns the sum of
// everything in between (= the exclusive prefix of `tile`).  Tiles before
// `first` count as an inclusive prefix of 0.
//
// The prefix frontier of a chain advances by at most one look-back window per L2
// round trip, so the window is LB_WINDOW = 128 descriptors (4 per lane), not 32:
// with 4 KiB tiles a 32-wide window capped a file-wide chain at ~250 GB/s.
#ifndef SHAFA_LB_SPINS
#define SHAFA_LB_SPINS 2  // failed probes before a look-back starts to sleep between probes
#endif
#ifndef SHAFA_LB_PER_LANE
#define SHAFA_LB_PER_LANE 4
#endif
constexpr int LB_PER_LANE = SHAFA_LB_PER_LANE;
constexpr int LB_WINDOW = 32 * LB_PER_LANE;

__device__ __forceinline__ uint64_t lookback_read(const uint64_t *desc, int64_t tile, int64_t first)
{
    const unsigned lane = lane_id();
    uint64_t excl = 0;
    int64_t pos = tile - 1;
    for (;;) {
        // lane l looks at pos - l - 32*j: j = 0 is the nearest group of 32 predecessors
        uint64_t d[LB_PER_LANE];
        bool pending;
        unsigned spins = 0;
        do {
            pending = false;
#pragma unroll
            for (int j = 0; j < LB_PER_LANE; j++) {
                const int64_t idx = pos - (int64_t)lane - 32 * j;
                d[j] = (idx >= first) ? desc_load(desc + idx) : desc_pack(DESC_PREFIX, 0);
            }
            // only descriptors nearer than the first published prefix have to be valid
#pragma unroll
            for (int j = 0; j < LB_PER_LANE; j++) {
                const unsigned inv = __ballot_sync(FULL, desc_state(d[j]) == DESC_INVALID);
                const unsigned pfx = __ballot_sync(FULL, desc_state(d[j]) == DESC_PREFIX);
                if (pfx) {
                    const unsigned nearer = (1u << (__ffs(pfx) - 1)) - 1u;  // lanes before the first prefix
                    if (inv & nearer) pending = true;
                    break;
                }
                if (inv) {
                    pending = true;
                    break;
                }
            }
            if (pending && ++spins > SHAFA_LB_SPINS) __nanosleep(40);
        } while (pending);
        bool done = false;
#pragma unroll
        for (int j = 0; j < LB_PER_LANE; j++) {
            if (!done) {
                const unsigned pm = __ballot_sync(FULL, desc_state(d[j]) == DESC_PREFIX);
                const unsigned upto = pm ? (unsigned)(__ffs(pm) - 1) : 31u;  // lanes 0..upto contribute
                excl += warp_sum64(lane <= upto ? desc_value(d[j]) : 0ull);
                if (pm) done = true;
            }
        }
        if (done) break;
        pos -= LB_WINDOW;
    }
    return excl;
}

// Publish this tile's aggregate, look back, publish the inclusive prefix; returns
// the exclusive prefix.  Values must stay below 2^62.
__device__ __forceinline__ uint64_t lookback_sum(uint64_t *desc, int64_t tile, int64_t first, uint64_t aggregate)
{
    const unsigned lane = lane_id();
    if (tile == first) {
        if (lane == 0) desc_store(desc + tile, desc_pack(DESC_PREFIX, aggregate));
        return 0;
    }
    if (lane == 0) desc_store(desc + tile, desc_pack(DESC_AGG, aggregate));
    const uint64_t excl = lookback_read(desc, tile, first);
    if (lane == 0) desc_store(desc + tile, desc_pack(DESC_PREFIX, excl + aggregate));
    return excl;
}

// The same in two steps, so that work which does not need the prefix can run in between: publish the
// aggregate (one thread) as soon as it is known, resolve (one warp) when the prefix is needed.
__device__ __forceinline__ void lookback_publish(uint64_t *desc, int64_t tile, int64_t first, uint64_t aggregate)
{
    desc_store(desc + tile, desc_pack(tile == first ? DESC_PREFIX : DESC_AGG, aggregate));
}
__device__ __forceinline__ uint64_t lookback_resolve(uint64_t *desc, int64_t tile, int64_t first, uint64_t aggregate)
{
    if (tile == first) return 0;
    const uint64_t excl = lookback_read(desc, tile, first);
    if (lane_id() == 0) desc_store(desc + tile, desc_pack(DESC_PREFIX, excl + aggregate));
    return excl;
}

// Copy n bytes of shared memory (src 16-byte aligned, readable up to src + n + 32) to global memory at
// dst (ANY alignment) with 16-byte stores, by all NT threads of the CTA.  The tile is staged at
// tile-relative offsets -- before its place in the output is known -- so the two are not co-aligned:
// each 16-byte destination chunk is cut out of two aligned 16-byte shared loads with funnel shifts.
// The misalignment is the same for every chunk, so the word selection is a uniform branch.
template <int NT>
__device__ __forceinline__ void copy_out_unaligned(uint8_t *dst, const uint8_t *src, uint32_t n)
{
    const uint32_t tid = threadIdx.x;
    const uint32_t r = (16u - (uint32_t)((uintptr_t)dst & 15u)) & 15u;  // bytes before dst's first 16-byte boundary
    const uint32_t head = r < n ? r : n;
    if (tid < head) dst[tid] = src[tid];
    const uint32_t body = (n - head) >> 4;  // whole 16-byte chunks; chunk k = src bytes [r + 16k, r + 16k + 16)
    uint8_t *const d16 = dst + head;
    if (r == 0) {
        for (uint32_t k = tid; k < body; k += NT) st_stream16(d16 + 16 * k, *reinterpret_cast<const uint4 *>(src + 16 * k));
    } else {
        const uint32_t wr = r >> 2, sh8 = (r & 3u) * 8u;
        for (uint32_t k = tid; k < body; k += NT) {
            const uint4 a = *reinterpret_cast<const uint4 *>(src + 16 * k);
            const uint4 b = *reinterpret_cast<const uint4 *>(src + 16 * k + 16);
            uint32_t w0, w1, w2, w3, w4;  // the five words that hold the chunk
            if (wr == 0)
                w0 = a.x, w1 = a.y, w2 = a.z, w3 = a.w, w4 = b.x;
            else if (wr == 1)
                w0 = a.y, w1 = a.z, w2 = a.w, w3 = b.x, w4 = b.y;
            else if (wr == 2)
                w0 = a.z, w1 = a.w, w2 = b.x, w3 = b.y, w4 = b.z;
            else
                w0 = a.w, w1 = b.x, w2 = b.y, w3 = b.z, w4 = b.w;
            uint4 o;
            o.x = __funnelshift_r(w0, w1, sh8);
            o.y = __funnelshift_r(w1, w2, sh8);
            o.z = __funnelshift_r(w2, w3, sh8);
            o.w = __funnelshift_r(w3, w4, sh8);
            st_stream16(d16 + 16 * k, o);
        }
    }
    const uint32_t done = head + 16 * body;
    if (tid < n - done) dst[done + tid] = src[done + tid];
}

// 0x80 in every byte of the result whose byte of w is zero (exact, no carries between bytes)
__device__ __forceinline__ uint32_t zero_flags(uint32_t w) { return ~(((w & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | w | 0x7F7F7F7Fu); }

// bit j (j = 0..15) of the result = (byte j of the 16 bytes x,y,z,w is zero).  The flag bytes (0x80 or 0) are
// gathered with byte dot products against the bit weights 1,2,4,8 / 16,32,64,128 -- 128 * (the 8-bit mask) per
// pair of words -- so a vector costs four dp4a, one multiply-add and one shift on top of the flag tests.
__device__ __forceinline__ uint32_t zero_mask16_words(uint32_t x, uint32_t y, uint32_t z, uint32_t w)
{
    const uint32_t lo = __dp4a(zero_flags(y), 0x80402010u, __dp4a(zero_flags(x), 0x08040201u, 0u));
    const uint32_t hi = __dp4a(zero_flags(w), 0x80402010u, __dp4a(zero_flags(z), 0x08040201u, 0u));
    return (hi * 256u + lo) >> 7;
}

// bit j of the result = (byte j of v equals the byte before it); `prev` is the
// byte preceding byte 0
__device__ __forceinline__ uint32_t eq_prev_mask16(const uint4 &v, uint32_t prev)
{
    const uint32_t sx = (v.x << 8) | (prev & 0xffu);
    const uint32_t sy = __funnelshift_l(v.x, v.y, 8);
    const uint32_t sz = __funnelshift_l(v.y, v.z, 8);
    const uint32_t sw = __funnelshift_l(v.z, v.w, 8);
    return zero_mask16_words(v.x ^ sx, v.y ^ sy, v.z ^ sz, v.w ^ sw);
}

// bit j = (byte j of v == 0)
__device__ __forceinline__ uint32_t zero_mask16(const uint4 &v) { return zero_mask16_words(v.x, v.y, v.z, v.w); }

// Largest b with tile0[b] <= t  (tile0 has nb+1 monotone entries, tile0[nb] = total)
__device__ __forceinline__ uint32_t find_block(const uint64_t *__restrict__ tile0, uint32_t nb, uint64_t t)
{
    uint32_t lo = 0, hi = nb;  // invariant: tile0[lo] <= t < tile0[hi]
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(tile0 + mid) <= t)
            lo = mid;
        else
            hi = mid;
    }
    return lo;
}

// The same for a warp that walks tickets in increasing order and remembers the block of its last ticket: the 32
// lanes probe the next 32 blocks at once (one load latency instead of a dependent binary search); only a jump
// past them falls back to the search.  All lanes return the block.
__device__ __forceinline__ uint32_t find_block_forward(const uint64_t *__restrict__ tile0, uint32_t nb, uint64_t t, uint32_t from)
{
    const uint32_t cand = from + lane_id();
    const bool ok = cand < nb && __ldg(tile0 + cand) <= t;
    const unsigned m = __ballot_sync(FULL, ok);  // lane 0 is set: tile0[from] <= t by construction
    if (m == FULL && from + 32u < nb) {          // maybe further ahead than this window
        uint32_t b = 0;
        if (lane_id() == 0) b = from + 31u + find_block(tile0 + from + 31u, nb - (from + 31u), t);
        return __shfl_sync(FULL, b, 0);
    }
    return from + (31u - (uint32_t)__clz((int)m));
}

}  // namespace shafa
