// sf_encode.cu -- Shannon-Fano bit packing (reference compress_to_buffer +
// binary_coding, c.c:52-83,91-237): per block, the codes of its symbols
// concatenated MSB-first and zero padded to a byte.
//
// Single pass over the symbol stream (algorithmic bytes S + Z):
//   * exact payload sizes come from histogram x code length (k_sf_block_bits),
//     so every block gets a 16-byte aligned payload slot before encoding starts;
//   * tiles are 16 KiB aligned windows of the stream intersected with the
//     block (aligned 16-byte loads even for the ragged .rle stream), 64 symbols
//     per thread;
//   * per tile: conflict-free code LUT lookups (one private bank per lane), a
//     length pass and a CTA scan of bit lengths; the tile's bit count is published
//     at once.  Codes are packed at tile-relative bit positions through a 64-bit
//     register accumulator (whole words as plain shared stores, only a thread's
//     last partial word is OR-ed in); the decoupled look-back along the block's
//     tiles is resolved behind the packing, and the copy-out shifts by the
//     prefix's low five bits, byte-swaps and writes 16-byte groups.  A 32-bit
//     word that straddles two tiles is written by the tile holding its last bit,
//     which rebuilds the preceding bits from the <= 32 preceding symbols, so no
//     output word is ever touched twice and no memset/atomics on global memory
//     are needed.
//   * blocks whose longest code exceeds 26 bits (never seen on real data; the
//     format allows it) take k_sf_encode_long, a simple one-CTA-per-block kernel.
#include "common.cuh"
#include "core.h"
#include "internal.h"

namespace shafa {

constexpr int ENC_COPIES = 32;  // one LUT copy per lane -> no bank conflicts

// ---- exact sizes -------------------------------------------------------------------
// one warp per block: bits = sum hist*len; flags: bit0 = has a code > ENC_SHORT_MAX,
// bit1 = a symbol occurs but has no code
__global__ void __launch_bounds__(256)
k_sf_block_bits(const uint64_t *__restrict__ hist, const uint8_t *__restrict__ lens, uint32_t nb,
                uint64_t *__restrict__ bits_out, uint32_t *__restrict__ flags_out)
{
    const uint32_t b = blockIdx.x * 8 + warp_id();
    if (b >= nb) return;
    uint64_t bits = 0;
    uint32_t fl = 0, present = 0, uncoded = 0, coded = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int s = k * 32 + lane_id();
        const uint64_t h = hist[(uint64_t)b * 256 + s];
        const uint32_t l = lens[(uint64_t)b * 256 + s];
        bits += h * l;
        if (l > (uint32_t)ENC_SHORT_MAX) fl |= 1u;
        present += h != 0;
        uncoded += (h != 0 && l == 0);
        coded += l != 0;
    }
    bits = warp_sum64(bits);
    present = __reduce_add_sync(FULL, present);
    uncoded = __reduce_add_sync(FULL, uncoded);
    coded = __reduce_add_sync(FULL, coded);
    // A block with ONE distinct symbol gets the empty code (t.c:185: start == end appends no bit) and the
    // reference packer then writes "@0@" with no payload (c.c:65-80): bits = 0 here, not an error.  Any other
    // symbol that occurs without a code is a malformed .cod.
    if (uncoded && !(present == 1 && coded == 0)) fl |= 2u;
    fl = __reduce_or_sync(FULL, fl);
    if (lane_id() == 0) {
        bits_out[b] = bits;
        flags_out[b] = fl;
    }
}

// single CTA: z[b] = ceil(bits/8), payload_off[b] = exclusive scan of roundup16(z)
// summary[0] = total payload bytes, summary[1] = OR of flags
__global__ void __launch_bounds__(1024)
k_sf_offsets(const uint64_t *__restrict__ bits, const uint32_t *__restrict__ flags, uint32_t nb,
             uint64_t *__restrict__ z_out, uint64_t *__restrict__ off_out, uint64_t *__restrict__ summary)
{
    __shared__ uint64_t warp_tot[32];
    __shared__ uint64_t s_carry;
    __shared__ uint32_t s_flags;
    if (threadIdx.x == 0) {
        s_carry = 0;
        s_flags = 0;
    }
    __syncthreads();
    uint32_t fl = 0;
    for (uint32_t base = 0; base < nb; base += 1024) {
        const uint32_t b = base + threadIdx.x;
        uint64_t z = 0, a = 0;
        if (b < nb) {
            z = (bits[b] + 7) >> 3;
            a = (z + 15) & ~15ull;
            fl |= flags[b];
        }
        // inclusive warp scan (64-bit)
        uint64_t incl = a;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint64_t t = __shfl_up_sync(FULL, incl, d);
            if (lane_id() >= (unsigned)d) incl += t;
        }
        if (lane_id() == 31) warp_tot[warp_id()] = incl;
        __syncthreads();
        uint64_t wbase = 0;
        for (unsigned w = 0; w < warp_id(); w++) wbase += warp_tot[w];
        const uint64_t carry = s_carry;
        if (b < nb) {
            z_out[b] = z;
            off_out[b] = carry + wbase + incl - a;
        }
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + wbase + incl;
        __syncthreads();
    }
    fl = __reduce_or_sync(FULL, fl);
    if (lane_id() == 0 && fl) atomicOr(&s_flags, fl);
    __syncthreads();
    if (threadIdx.x == 0) {
        summary[0] = s_carry;
        summary[1] = s_flags;
    }
}

// ---- fast path ------------------------------------------------------------------------
struct EncShared {
    uint32_t lut[256 * ENC_COPIES];       // lut[sym*32 + lane]
    uint32_t stage[ENC_STAGE_WORDS + 8];  // big-endian words of (a window of) this tile's bit range; its first 256 words double as
                                          // the scratch of the LUT build at a block change (barrier-separated from the tile's use)
    uint32_t scan[ENC_THREADS / 32 + 1];
    uint64_t prefix;
    uint32_t ticket;
    uint32_t block;
};

// ld.global.nc with the default L1 policy: a thread reads 64 contiguous bytes as four
// 16-byte vectors, so neighbouring vectors share sectors and should hit L1
__device__ __forceinline__ uint4 ld_nc16(const void *p)
{
    uint4 v;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// Shared memory through explicit 32-bit shared addresses: one ADD per access instead of the
// generic->shared window arithmetic the compiler emits for dynamic shared memory, and the
// conditional OR as ONE predicated instruction instead of a branch around an atomic.
__device__ __forceinline__ uint32_t lds32(uint32_t saddr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}
// LUT entry of byte j (0..15, compile-time) of a 16-byte vector in this lane's column: byte extract (PRMT),
// multiply-add, load -- the generic (w >> 8k) & 0xff form costs one instruction more per lookup
__device__ __forceinline__ uint32_t enc_lookup(uint32_t lut_lane, const uint4 &x, int j)
{
    const uint32_t w = (j < 4) ? x.x : (j < 8) ? x.y : (j < 12) ? x.z : x.w;
    return lds32(lut_lane + __byte_perm(w, 0, 0x4440 | (j & 3)) * (ENC_COPIES * 4));
}
__device__ __forceinline__ void red_or_if(uint32_t saddr, uint32_t value, bool cond)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p red.shared.or.b32 [%0], %1;\n\t}" ::"r"(saddr), "r"(value),
                 "r"((uint32_t)cond)
                 : "memory");
}

// Place the thread's codes into the (zeroed) staging window, branch-free: a code of l bits at bit
// offset o lands in word o>>5 (and spills into the next word when it straddles); both pieces are
// OR-ed in with shared-memory reductions, so no data-dependent "a word is full" branch diverges the warp.
// orel = bit offset of the thread's first bit relative to the window start (negative if before it).
__device__ __forceinline__ void st_shared_if(uint32_t saddr, uint32_t value, uint32_t cond)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.shared.b32 [%0], %1;\n\t}" ::"r"(saddr), "r"(value), "r"(cond) : "memory");
}

// Common case: all 64 symbols valid and every bit of the thread inside the staging window.  The codes
// are shifted into a 64-bit accumulator and leave it one whole 32-bit word at a time, as plain stores:
// a word the thread completes holds no bits of later threads, and the bits of EARLIER threads in its
// first word (zeros here) are OR-ed in afterwards -- every thread's last, partial word is returned and
// merged with a reduction only after a barrier (see the kernel).
// PAIR: no code of the block is longer than 16 bits, so two codes are joined before they enter the
// accumulator (31 pending bits + 32 <= 63) and the flush test runs once per two symbols.
template <bool PAIR>
__device__ __forceinline__ void enc_pack_fast(uint32_t lut_lane, uint32_t stage_s, uint4 (&v)[ENC_VECS], uint32_t orel,
                                              uint32_t *tail_addr, uint32_t *tail_word)
{
    uint32_t nbits = orel & 31u;                  // bits of the open word that belong to earlier threads
    uint32_t waddr = stage_s + (orel >> 5) * 4u;  // the open word
    uint64_t acc = 0;
#pragma unroll
    for (int c = 0; c < ENC_VECS; c++) {
        // opaque: the table addresses of later vectors must not be precomputed (they would be spilled)
        asm volatile("" : "+r"(v[c].x), "+r"(v[c].y), "+r"(v[c].z), "+r"(v[c].w));
#pragma unroll
        for (int j = 0; j < 16; j += (PAIR ? 2 : 1)) {
            const uint32_t e0 = enc_lookup(lut_lane, v[c], j);
            uint32_t code = enc_code(e0), l = enc_len(e0);
            if (PAIR) {
                const uint32_t e1 = enc_lookup(lut_lane, v[c], j + 1);
                const uint32_t l1 = enc_len(e1);
                code = (code << l1) | enc_code(e1);
                l += l1;
            }
            acc = (acc << l) | code;
            nbits += l;
            const uint32_t full = nbits >> 5;  // 0 or 1
            st_shared_if(waddr, (uint32_t)(acc >> ((nbits - 32u) & 31u)), full);
            waddr += full * 4u;
            nbits &= 31u;
        }
    }
    *tail_addr = waddr;
    *tail_word = nbits ? (uint32_t)(acc << (32u - nbits)) : 0u;
}

template <bool ALL_VALID>
__device__ __forceinline__ void enc_pack_thread(uint32_t lut_lane, uint32_t stage_s, const uint4 (&v)[ENC_VECS], uint64_t vmask,
                                                int32_t orel, uint32_t nw)
{
    int32_t o = orel;
#pragma unroll
    for (int c = 0; c < ENC_VECS; c++) {
#pragma unroll
        for (int j = 0; j < 16; j++) {
            uint32_t ee = lds32(lut_lane + byte_of(v[c], j) * (ENC_COPIES * 4));
            if (!ALL_VALID) ee = ((vmask >> (c * 16 + j)) & 1ull) ? ee : 0u;
            const uint32_t l = enc_len(ee);
            const uint32_t q = (uint32_t)o & 31u;
            const uint32_t w = (uint32_t)(o >> 5);                             // arithmetic shift: before the window -> huge
            const uint64_t v64 = (uint64_t)enc_code(ee) << ((64u - q - l) & 63u);  // q + l <= 57; l = 0 carries no bits
            const uint32_t hi = (uint32_t)(v64 >> 32), lo = (uint32_t)v64;
            red_or_if(stage_s + w * 4u, hi, hi != 0 && w < nw);
            red_or_if(stage_s + w * 4u + 4u, lo, lo != 0 && w + 1u < nw);
            o += (int32_t)l;
        }
    }
}

#ifndef SHAFA_ENC_MIN_CTAS
#define SHAFA_ENC_MIN_CTAS 4
#endif
__global__ void __launch_bounds__(ENC_THREADS, SHAFA_ENC_MIN_CTAS)
k_sf_encode(const uint8_t *__restrict__ in, const uint64_t *__restrict__ off, const uint64_t *__restrict__ tile0,
            uint32_t nb, const uint8_t *__restrict__ lens, const uint64_t *__restrict__ codes,
            const uint64_t *__restrict__ pay_off, uint8_t *__restrict__ payload, uint64_t *__restrict__ desc,
            uint32_t *__restrict__ counter, uint32_t *__restrict__ err)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    EncShared &sh = *reinterpret_cast<EncShared *>(smem_raw);
    const unsigned tid = threadIdx.x, lane = lane_id(), warp = warp_id();
    const uint64_t total_tiles = tile0[nb];
    uint32_t cur_block = 0xffffffffu;
    bool cur_long = false, cur_pair = false;

    // A ticket is taken only when the CTA is ready to work on it (a tile reserved ahead of time stalls every later
    // tile of the look-back chains: measured).  Warp 0 finds the ticket's block by probing the 32 blocks after its
    // previous one at once instead of a dependent binary search.
    uint32_t last_block = 0;
    for (;;) {
        __syncthreads();  // everyone is done with the previous tile's shared state
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
        if (b != cur_block) {
            cur_block = b;
            const uint32_t l = lens[(uint64_t)b * 256 + tid];
            const uint64_t c = codes[(uint64_t)b * 256 + tid];
            cur_long = __syncthreads_or(l > (uint32_t)ENC_SHORT_MAX) != 0;
            cur_pair = __syncthreads_or(l > 16u) == 0;
            if (!cur_long) {
                sh.stage[tid] = enc_entry(c, l);
                __syncthreads();
#pragma unroll 4
                for (int i = 0; i < ENC_COPIES; i++) {
                    const int idx = i * ENC_THREADS + tid;  // consecutive lanes -> consecutive banks
                    sh.lut[idx] = sh.stage[idx / ENC_COPIES];
                }
                __syncthreads();
            }
        }
        if (cur_long) continue;  // k_sf_encode_long owns this block

        const uint64_t lo = off[b], hi = off[b + 1];
        const uint64_t first_tile = tile0[b];
        const bool is_first = (t == first_tile);
        const bool is_last = (t + 1 == tile0[b + 1]);
        const uint64_t base = (lo / ENC_TILE + (t - first_tile)) * ENC_TILE;
        const uint64_t pos = base + (uint64_t)tid * ENC_ITEMS;
        const bool all_valid = (pos >= lo) && (pos + ENC_ITEMS <= hi);

        uint4 v[ENC_VECS];
        uint64_t vmask = 0;  // bit k: symbol k of this thread lies inside the block
#pragma unroll
        for (int c = 0; c < ENC_VECS; c++) {
            const uint64_t pc = pos + 16 * c;
            v[c] = make_uint4(0, 0, 0, 0);
            if (pc + 16 > lo && pc < hi) v[c] = ld_nc16(in + pc);
        }
        const uint32_t lut_lane = smem_u32(sh.lut) + lane * 4u;  // this lane's private LUT column
        uint32_t bits = 0;
        if (all_valid) {
            vmask = ~0ull;
#pragma unroll
            for (int c = 0; c < ENC_VECS; c++)
#pragma unroll
                for (int j = 0; j < 16; j++) bits += enc_len(enc_lookup(lut_lane, v[c], j));
        } else {
#pragma unroll
            for (int c = 0; c < ENC_VECS; c++)
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    const uint64_t q = pos + 16 * c + j;
                    if (q >= lo && q < hi) {
                        vmask |= 1ull << (c * 16 + j);
                        bits += enc_len(lds32(lut_lane + byte_of(v[c], j) * (ENC_COPIES * 4)));
                    }
                }
        }
        // keep the compiler from carrying 64 precomputed table addresses from this pass to the packing pass
        // (they do not fit in registers and were spilled to local memory)
#pragma unroll
        for (int c = 0; c < ENC_VECS; c++) asm volatile("" : "+r"(v[c].x), "+r"(v[c].y), "+r"(v[c].z), "+r"(v[c].w));
        uint32_t tile_bits;
        const uint32_t excl = block_excl_scan<ENC_THREADS>(bits, sh.scan, &tile_bits);
        // The tile's size is published at once; its place in the block's bit stream (P) is resolved behind
        // the packing, which works in tile-relative bit positions.  stage[1 + j] is word j of the tile's own
        // bit string; stage[0] holds the 32 bits just before it, so that output word k of the tile -- bits
        // [32k - r, 32k - r + 32) of the tile with r = P mod 32 -- is one funnel shift of stage[k], stage[k+1].
        if (tid == 0) lookback_publish(desc, (int64_t)t, (int64_t)first_tile, tile_bits);
        const uint32_t Tw = (tile_bits + 31u) >> 5;                           // words of the tile's own bit string
        const uint32_t wf = excl >> 5, wl = (excl + bits + 31u) >> 5;         // my words: [wf, wl)
        uint32_t *out32 = reinterpret_cast<uint32_t *>(payload + pay_off[b]);
        const uint32_t stage_s = smem_u32(sh.stage) + 4u;                     // shared address of stage[1]
        uint64_t P = 0, W0 = 0;
        uint32_t r = 0, Kown = 0;

        // the bit string is materialised window by window (one window unless codes are very long here)
        for (uint32_t jw = 0; jw < Tw || jw == 0; jw += ENC_STAGE_WORDS) {
            const uint32_t nw = (Tw - jw < (uint32_t)ENC_STAGE_WORDS) ? (Tw - jw) : (uint32_t)ENC_STAGE_WORDS;
            if (jw) {  // the word just before this window
                if (tid == 0) sh.stage[0] = sh.stage[ENC_STAGE_WORDS];
                __syncthreads();
            }
            for (uint32_t k = tid; k <= nw; k += ENC_THREADS) sh.stage[1 + k] = 0;  // one more: read as the word after the end
            __syncthreads();
            const bool mine = bits && wl > jw && wf < jw + nw;
            const bool fast = mine && all_valid && wf >= jw && wl <= jw + nw;
            const int32_t orel = (int32_t)excl - (int32_t)(jw << 5);  // a tile holds < 2^19 bits
            uint32_t tail_addr = 0, tail_word = 0;
            if (fast) {
                if (cur_pair)
                    enc_pack_fast<true>(lut_lane, stage_s, v, (uint32_t)orel, &tail_addr, &tail_word);
                else
                    enc_pack_fast<false>(lut_lane, stage_s, v, (uint32_t)orel, &tail_addr, &tail_word);
            }
            __syncthreads();  // plain stores of whole words first, then every OR into a shared word
            if (fast) {
                red_or_if(tail_addr, tail_word, tail_word != 0);
            } else if (mine) {
                if (all_valid)
                    enc_pack_thread<true>(lut_lane, stage_s, v, vmask, orel, nw);
                else
                    enc_pack_thread<false>(lut_lane, stage_s, v, vmask, orel, nw);
            }
            if (jw == 0) {
                if (warp == 1) {
                    // stage[0]: the last 32 bits of the block's stream before this tile = the codes of the
                    // (up to) 32 symbols just before it, right-aligned (what lies further back is shifted out)
                    uint32_t ee = 0;
                    if (base >= lo + 1 + lane) ee = sh.lut[(uint32_t)in[base - 1 - lane] * ENC_COPIES + lane];
                    const uint32_t l = enc_len(ee);
                    const uint32_t s_prev = warp_incl_scan(l) - l;  // bits of nearer symbols
                    uint32_t w = (l && s_prev < 32u) ? (enc_code(ee) << s_prev) : 0u;
                    w = __reduce_or_sync(FULL, w);
                    if (lane == 0) sh.stage[0] = w;
                }
                if (warp == 0) {
                    const uint64_t pfx = lookback_resolve(desc, (int64_t)t, (int64_t)first_tile, tile_bits);
                    if (lane == 0) sh.prefix = pfx;
                }
            }
            __syncthreads();
            if (jw == 0) {
                P = sh.prefix;  // bits of this block before this tile
                const uint64_t E = P + tile_bits;
                W0 = P >> 5;
                r = (uint32_t)(P & 31u);
                const uint64_t Wown = is_last ? ((E + 31) >> 5) : (E >> 5);  // words whose last bit lies in [P, E) (all, for the last tile)
                Kown = (uint32_t)(Wown - W0);
            }
            // output words k in [jw, kend): word k = stage[k - jw], stage[k - jw + 1] shifted by r
            const bool last_win = (jw + nw >= Tw);
            uint32_t kend = last_win ? Kown : (jw + nw);
            if (kend > jw + nw + 1) kend = jw + nw + 1;
            if (kend > Kown) kend = Kown;
            if (kend > jw) {
                const uint64_t gfirst = W0 + jw, glast = W0 + kend;  // global words of the block's payload
                for (uint64_t Q = (gfirst >> 2) + tid; Q * 4 < glast; Q += ENC_THREADS) {
                    const int64_t s0 = (int64_t)(Q * 4) - (int64_t)gfirst;  // stage index of the quad's first "previous" word
                    if (Q * 4 >= gfirst && Q * 4 + 4 <= glast) {
                        const uint32_t a0 = sh.stage[s0], a1 = sh.stage[s0 + 1], a2 = sh.stage[s0 + 2], a3 = sh.stage[s0 + 3],
                                       a4 = sh.stage[s0 + 4];
                        uint4 o;
                        o.x = bswap32(__funnelshift_r(a1, a0, r));
                        o.y = bswap32(__funnelshift_r(a2, a1, r));
                        o.z = bswap32(__funnelshift_r(a3, a2, r));
                        o.w = bswap32(__funnelshift_r(a4, a3, r));
                        st_stream16(out32 + Q * 4, o);
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            const uint64_t g = Q * 4 + i;
                            if (g >= gfirst && g < glast)
                                st_stream4(out32 + g, bswap32(__funnelshift_r(sh.stage[s0 + i + 1], sh.stage[s0 + i], r)));
                        }
                    }
                }
            }
            if (!last_win) __syncthreads();  // the next window re-zeroes the stage
        }
    }
    (void)err;  // a symbol without a code is caught by k_sf_block_bits (histogram x length) before this kernel runs
}

// ---- generic path: any code length up to 64 bits, one CTA per block ---------------------
__global__ void __launch_bounds__(256)
k_sf_encode_long(const uint8_t *__restrict__ in, const uint64_t *__restrict__ off, uint32_t nb,
                 const uint8_t *__restrict__ lens, const uint64_t *__restrict__ codes,
                 const uint64_t *__restrict__ pay_off, const uint64_t *__restrict__ z, uint8_t *__restrict__ payload,
                 uint32_t *__restrict__ err)
{
    __shared__ uint64_t s_code[256];
    __shared__ uint32_t s_len[256];
    __shared__ uint32_t scan[9];
    __shared__ unsigned long long s_carry;
    const unsigned tid = threadIdx.x;
    const uint32_t b = blockIdx.x;
    const uint32_t l = lens[(uint64_t)b * 256 + tid];
    s_len[tid] = l;
    s_code[tid] = codes[(uint64_t)b * 256 + tid];
    if (tid == 0) s_carry = 0;
    if (!__syncthreads_or(l > (uint32_t)ENC_SHORT_MAX)) return;  // fast path handled it
    const uint64_t lo = off[b], hi = off[b + 1];
    uint32_t *out32 = reinterpret_cast<uint32_t *>(payload + pay_off[b]);
    const uint64_t nwords = (z[b] + 3) >> 2;
    for (uint64_t w = tid; w < nwords; w += 256) out32[w] = 0;
    __syncthreads();
    uint32_t bad = 0;
    for (uint64_t base = lo; base < hi; base += 256 * 8) {
        const uint64_t p0 = base + (uint64_t)tid * 8;
        uint32_t bits = 0;
        uint8_t sy[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            sy[j] = (p0 + j < hi) ? in[p0 + j] : 0;
            if (p0 + j < hi) {
                bits += s_len[sy[j]];
                if (s_len[sy[j]] == 0) bad = 1;
            }
        }
        uint32_t total;
        const uint32_t excl = block_excl_scan<256>(bits, scan, &total);
        uint64_t q = s_carry + excl;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (p0 + j < hi) {
                uint32_t len = s_len[sy[j]];
                const uint64_t code = s_code[sy[j]];
                while (len) {
                    const uint32_t room = 32 - (uint32_t)(q & 31u);
                    const uint32_t take = len < room ? len : room;
                    const uint32_t piece = (uint32_t)((code >> (len - take)) & ((take == 32) ? 0xffffffffu : ((1u << take) - 1u)));
                    // big-endian word, stored byte-swapped
                    atomicOr(out32 + (q >> 5), bswap32(piece << (room - take)));
                    q += take;
                    len -= take;
                }
            }
        }
        __syncthreads();
        if (tid == 0) s_carry += total;
        __syncthreads();
    }
    if (bad) atomicOr(err, 4u);
}

int launch_sf_sizes(shafa_cuda_ctx *ctx, const uint64_t *d_hist, const uint8_t *d_len, uint64_t nb, uint64_t *d_bits,
                    uint32_t *d_flags, uint64_t *d_z, uint64_t *d_payload_off, uint64_t *d_summary)
{
    SHAFA_LAUNCH(ctx, "sf_block_bits", k_sf_block_bits, dim3((unsigned)((nb + 7) / 8)), dim3(256), 0, d_hist, d_len,
                 (uint32_t)nb, d_bits, d_flags);
    SHAFA_LAUNCH(ctx, "sf_offsets", k_sf_offsets, dim3(1), dim3(1024), 0, d_bits, d_flags, (uint32_t)nb, d_z,
                 d_payload_off, d_summary);
    SHAFA_CHECK_LAUNCH(ctx);
    return SHAFA_SUCCESS;
}

int launch_sf_encode(shafa_cuda_ctx *ctx, const uint8_t *d_sym, const uint64_t *h_off, uint64_t nb,
                     const uint8_t *d_len, const uint64_t *d_code, const uint64_t *d_payload_off, const uint64_t *d_z,
                     uint8_t *d_payload, bool any_long, uint32_t *d_err)
{
    int &max_ctas_per_sm = ctx->occ[OCC_SF_ENCODE];
    const size_t smem = sizeof(EncShared);
    if (!max_ctas_per_sm) {
        SHAFA_CUDA_TRY(ctx, cudaFuncSetAttribute(k_sf_encode, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SHAFA_CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&max_ctas_per_sm, k_sf_encode, ENC_THREADS, smem));
        if (max_ctas_per_sm < 1) max_ctas_per_sm = 1;
    }
    // plan: [off (nb+1)] [tile0 (nb+1)]
    uint64_t *hp = (uint64_t *)pinned_get(ctx, (nb + 1) * 2 * sizeof(uint64_t));
    uint64_t *dp = (uint64_t *)scratch_get(ctx, SCR_PLAN, (nb + 1) * 2 * sizeof(uint64_t));
    if (!hp || !dp) return SHAFA_LACK_OF_MEMORY;
    uint64_t tiles = 0;
    for (uint64_t b = 0; b < nb; b++) {
        hp[b] = h_off[b];
        hp[nb + 1 + b] = tiles;
        const uint64_t lo = h_off[b], hi = h_off[b + 1];
        if (hi < lo) return set_error(ctx, SHAFA_FILE_STREAM_FAILED, "sf_encode: offsets not monotone");
        if (hi > lo) tiles += (hi - 1) / ENC_TILE - lo / ENC_TILE + 1;
    }
    hp[nb] = h_off[nb];
    hp[2 * nb + 1] = tiles;
    if (tiles > 0xfffffff0ull) return set_error(ctx, SHAFA_LACK_OF_MEMORY, "sf_encode: too many tiles");
    SHAFA_CUDA_TRY(ctx, cudaMemcpyAsync(dp, hp, (nb + 1) * 2 * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    uint64_t *desc = (uint64_t *)scratch_get(ctx, SCR_DESC, (tiles + 1) * sizeof(uint64_t));
    uint32_t *counter = misc_ptr(ctx, MISC_COUNTER);
    if (!desc || !counter) return SHAFA_LACK_OF_MEMORY;
    SHAFA_CUDA_TRY(ctx, cudaMemsetAsync(desc, 0, (tiles + 1) * sizeof(uint64_t), ctx->stream));
    SHAFA_CUDA_TRY(ctx, cudaMemsetAsync(counter, 0, 4, ctx->stream));
    if (tiles) {
        uint64_t grid = (uint64_t)ctx->sm_count * max_ctas_per_sm;
        if (grid > tiles) grid = tiles;
        SHAFA_LAUNCH(ctx, "sf_encode", k_sf_encode, dim3((unsigned)grid), dim3(ENC_THREADS), smem, d_sym, dp, dp + nb + 1,
                     (uint32_t)nb, d_len, d_code, d_payload_off, d_payload, desc, counter, d_err);
        if (any_long)
            SHAFA_LAUNCH(ctx, "sf_encode_long", k_sf_encode_long, dim3((unsigned)nb), dim3(256), 0, d_sym, dp,
                         (uint32_t)nb, d_len, d_code, d_payload_off, d_z, d_payload, d_err);
    }
    SHAFA_CHECK_LAUNCH(ctx);
    return SHAFA_SUCCESS;
}

}  // namespace shafa
