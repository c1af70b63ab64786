// sf_wide.cu -- Shannon-Fano packing and decoding for code tables with codes LONGER THAN 64 BITS.
//
// Module T cannot produce them (block totals < 2^31 bound a Shannon-Fano code at 52 bits), but the
// formats allow them: the reference's packer keeps up to 32 bytes per code (c.c:18, 115-160) and its decoder builds a tree
// from the strings of the .cod file whatever their length (d.c:422-504).  A hand-made .cod can therefore carry codes of
// up to 255 bits, and `shafa -m c` / `shafa -m d` have to honour it.  These kernels are the plain formulation -- one
// CTA (packing) or one warp (decoding) per block, nothing of the tiled machinery of sf_encode.cu / sf_decode.cu -- and
// run only when a table holds such a code: correctness for every file the reference accepts, not a fast path.
//
// A code is four 64-bit words, word 0 the least significant; its length is 1..255.
#include "common.cuh"
#include "core.h"
#include "internal.h"

namespace shafa {

constexpr int WIDE_WORDS = 4;       // 64-bit words per code
constexpr int WIDE_MAX_BITS = 255;  // the length is a byte, like everywhere else

// bits [lo, lo + take) of the code, take <= 32
__device__ __forceinline__ uint32_t wide_bits(const uint64_t *c, uint32_t lo, uint32_t take)
{
    const uint32_t w = lo >> 6, sh = lo & 63u;
    uint64_t v = c[w] >> sh;
    if (sh + take > 64u && w + 1 < (uint32_t)WIDE_WORDS) v |= c[w + 1] << (64u - sh);
    return (uint32_t)v & (take == 32u ? 0xffffffffu : ((1u << take) - 1u));
}

// ---- packing (c.c:52-83): one CTA per block, bits OR-ed into the zeroed payload -------------------------------------
__global__ void __launch_bounds__(256)
k_sf_encode_wide(const uint8_t *__restrict__ in, const uint64_t *__restrict__ off, const uint8_t *__restrict__ lens,
                 const uint64_t *__restrict__ codes, const uint64_t *__restrict__ pay_off, const uint64_t *__restrict__ z,
                 uint8_t *__restrict__ payload, uint32_t *__restrict__ err)
{
    __shared__ uint64_t s_code[256 * WIDE_WORDS];
    __shared__ uint32_t s_len[256];
    __shared__ uint32_t scan[9];
    __shared__ unsigned long long s_carry;
    const unsigned tid = threadIdx.x;
    const uint32_t b = blockIdx.x;
    s_len[tid] = lens[(uint64_t)b * 256 + tid];
#pragma unroll
    for (int k = 0; k < WIDE_WORDS; k++) s_code[tid * WIDE_WORDS + k] = codes[((uint64_t)b * 256 + tid) * WIDE_WORDS + k];
    if (tid == 0) s_carry = 0;
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
#pragma unroll 1
        for (int j = 0; j < 8; j++) {
            if (p0 + j < hi) {
                uint32_t len = s_len[sy[j]];
                const uint64_t *code = s_code + (uint32_t)sy[j] * WIDE_WORDS;
                while (len) {  // most significant bits first, at most up to the next 32-bit boundary of the payload
                    const uint32_t room = 32u - (uint32_t)(q & 31u);
                    const uint32_t take = len < room ? len : room;
                    const uint32_t piece = wide_bits(code, len - take, take);
                    atomicOr(out32 + (q >> 5), bswap32(piece << (room - take)));  // big-endian word, stored byte-swapped
                    q += take;
                    len -= take;
                }
            }
        }
        __syncthreads();
        if (tid == 0) s_carry += total;
        __syncthreads();
    }
    if (bad) atomicOr(err, (uint32_t)ERR_NO_CODE);
}

// ---- decoding (d.c:514-551): one warp per block ------------------------------------------------------------------------
// A 12-bit table answers the codes of up to 12 bits; everything else is matched by the whole warp against the left-aligned
// codes (eight symbols per lane, eight 32-bit words per code).  The payload is read as big-endian 32-bit windows.
__device__ __forceinline__ uint32_t wide_window32(const uint8_t *__restrict__ p, uint64_t z, uint64_t bit)
{
    const uint64_t byte = bit >> 3;
    uint64_t v = 0;
#pragma unroll
    for (int k = 0; k < 5; k++) v = (v << 8) | (byte + k < z ? (uint64_t)p[byte + k] : 0ull);
    return (uint32_t)(v >> (8u - (uint32_t)(bit & 7u)));
}

__global__ void __launch_bounds__(32)
k_sf_decode_wide(const uint8_t *__restrict__ payload, const uint64_t *__restrict__ pay_off, const uint64_t *__restrict__ zs,
                 const uint64_t *__restrict__ ns, const uint64_t *__restrict__ out_off, const uint8_t *__restrict__ lens,
                 const uint64_t *__restrict__ codes, uint8_t *__restrict__ out, uint32_t *__restrict__ err)
{
    __shared__ uint32_t s_left[256 * 8];  // code s left-aligned in 256 bits, word 0 = the first 32 bits
    __shared__ uint8_t s_len[256];
    __shared__ uint16_t s_lut[DEC_LUT_SIZE];  // (sym << 8) | len for codes of up to 12 bits, 0 = look further
    const unsigned lane = threadIdx.x;
    const uint32_t b = blockIdx.x;
    const uint8_t *p = payload + pay_off[b];
    const uint64_t z = zs[b], n = ns[b], zbits = z * 8;
    uint8_t *dst = out + out_off[b];
    for (int e = lane; e < DEC_LUT_SIZE; e += 32) s_lut[e] = 0;
    __syncwarp();
    for (int k = 0; k < 8; k++) {
        const int s = k * 32 + (int)lane;
        const uint32_t len = lens[(uint64_t)b * 256 + s];
        s_len[s] = (uint8_t)len;
        const uint64_t *c = codes + ((uint64_t)b * 256 + s) * WIDE_WORDS;
        const uint64_t cw[WIDE_WORDS] = {c[0], c[1], c[2], c[3]};
        for (int w = 0; w < 8; w++) {  // bits [256 - 32 (w + 1), 256 - 32 w) of (code << (256 - len))
            uint32_t word = 0;
            const int lo = (int)len - 32 * (w + 1);  // position of the word's lowest bit inside the code
            if (len) {
                if (lo >= 0)
                    word = wide_bits(cw, (uint32_t)lo, 32u);
                else if (lo > -32)
                    word = wide_bits(cw, 0u, (uint32_t)(32 + lo)) << (uint32_t)(-lo);
            }
            s_left[s * 8 + w] = word;
        }
    }
    __syncwarp();
    for (int k = 0; k < 8; k++) {  // short codes into the table (a prefix-free table never writes an entry twice)
        const int s = k * 32 + (int)lane;
        const uint32_t len = s_len[s];
        if (len >= 1 && len <= (uint32_t)DEC_LUT_BITS) {
            const uint32_t first = s_left[s * 8] >> (32 - DEC_LUT_BITS);
            for (uint32_t e = 0; e < (1u << (DEC_LUT_BITS - len)); e++) s_lut[first + e] = (uint16_t)((s << 8) | len);
        }
    }
    __syncwarp();
    uint64_t q = 0;
    uint32_t mine = 0, fail = 0;
    for (uint64_t i = 0; i < n; i++) {
        uint32_t sym = 0, len = 0;
        const uint32_t w0 = wide_window32(p, z, q);
        const uint32_t e = s_lut[w0 >> (32 - DEC_LUT_BITS)];
        if (e) {
            sym = e >> 8;
            len = e & 0xFFu;
        } else {
            // the next 256 bits: lanes 0..7 fetch a word each, everybody gets all eight
            const uint32_t part = lane < 8 ? wide_window32(p, z, q + 32ull * lane) : 0u;
            uint32_t W[8];
#pragma unroll
            for (int w = 0; w < 8; w++) W[w] = __shfl_sync(FULL, part, w);
            uint32_t found = 0xffffffffu;
            for (int k = 0; k < 8; k++) {
                const int s = k * 32 + (int)lane;
                const uint32_t l = s_len[s];
                if (l <= (uint32_t)DEC_LUT_BITS) continue;  // absent, or short (the table would have answered)
                bool same = true;
#pragma unroll
                for (int w = 0; w < 8; w++) {
                    const int covered = (int)l - 32 * w;  // how many bits of word w belong to the code
                    if (covered <= 0) break;
                    const uint32_t mask = covered >= 32 ? 0xffffffffu : ~(0xffffffffu >> covered);
                    if ((W[w] ^ s_left[s * 8 + w]) & mask) same = false;
                }
                if (same) found = (uint32_t)s;
            }
            const unsigned who = __ballot_sync(FULL, found != 0xffffffffu);
            if (who) {
                sym = __shfl_sync(FULL, found, __ffs((int)who) - 1);
                len = s_len[sym];
            }
        }
        if (len == 0 || q + len > zbits) {  // no code matches, or the payload ends inside the code (d.c:533-536 dereferences NULL there)
            fail = len == 0 ? (uint32_t)ERR_BAD_WINDOW : (uint32_t)ERR_PAYLOAD_SHORT;
            break;
        }
        q += len;
        if ((i & 31u) == lane) mine = sym;
        if ((i & 31u) == 31u) dst[i - 31u + lane] = (uint8_t)mine;  // 32 symbols per coalesced store
    }
    if (!fail && (n & 31u) && lane < (n & 31u)) dst[(n & ~31ull) + lane] = (uint8_t)mine;
    if (fail && lane == 0) atomicOr(err, fail);
}

// plan arrays on the device: enc [off (nb+1)]; dec [pay_off nb][z nb][n nb][out_off nb]
int launch_sf_encode_wide(shafa_cuda_ctx *ctx, const uint8_t *d_sym, const uint64_t *h_off, uint64_t nb, const uint8_t *d_len,
                          const uint64_t *d_code4, const uint64_t *d_payload_off, const uint64_t *d_z, uint8_t *d_payload,
                          uint32_t *d_err)
{
    uint64_t *hp = (uint64_t *)pinned_get(ctx, (nb + 1) * sizeof(uint64_t));
    uint64_t *dp = (uint64_t *)scratch_get(ctx, SCR_PLAN, (nb + 1) * sizeof(uint64_t));
    if (!hp || !dp) return SHAFA_LACK_OF_MEMORY;
    for (uint64_t b = 0; b <= nb; b++) {
        if (b && h_off[b] < h_off[b - 1]) return set_error(ctx, SHAFA_FILE_STREAM_FAILED, "sf_encode_wide: offsets not monotone");
        hp[b] = h_off[b];
    }
    SHAFA_CUDA_TRY(ctx, cudaMemcpyAsync(dp, hp, (nb + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    SHAFA_LAUNCH(ctx, "sf_encode_wide", k_sf_encode_wide, dim3((unsigned)nb), dim3(256), 0, d_sym, dp, d_len, d_code4, d_payload_off,
                 d_z, d_payload, d_err);
    SHAFA_CHECK_LAUNCH(ctx);
    return SHAFA_SUCCESS;
}

int launch_sf_decode_wide(shafa_cuda_ctx *ctx, const uint8_t *d_payload, const uint64_t *h_payload_off, const uint64_t *h_z,
                          const uint64_t *h_n, uint64_t nb, const uint8_t *d_len, const uint64_t *d_code4, uint8_t *d_sym,
                          uint32_t *d_err)
{
    uint64_t *hp = (uint64_t *)pinned_get(ctx, 4 * nb * sizeof(uint64_t));
    uint64_t *dp = (uint64_t *)scratch_get(ctx, SCR_PLAN, 4 * nb * sizeof(uint64_t));
    if (!hp || !dp) return SHAFA_LACK_OF_MEMORY;
    uint64_t out = 0;
    for (uint64_t b = 0; b < nb; b++) {
        if (h_n[b] != 0 && h_z[b] == 0) return set_error(ctx, SHAFA_FILE_UNRECOGNIZABLE, "sf_decode_wide: block %llu has symbols but an empty payload", (unsigned long long)b);
        hp[b] = h_payload_off[b];
        hp[nb + b] = h_z[b];
        hp[2 * nb + b] = h_n[b];
        hp[3 * nb + b] = out;
        out += h_n[b];
    }
    SHAFA_CUDA_TRY(ctx, cudaMemcpyAsync(dp, hp, 4 * nb * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    SHAFA_LAUNCH(ctx, "sf_decode_wide", k_sf_decode_wide, dim3((unsigned)nb), dim3(32), 0, d_payload, dp, dp + nb, dp + 2 * nb,
                 dp + 3 * nb, d_len, d_code4, d_sym, d_err);
    SHAFA_CHECK_LAUNCH(ctx);
    return SHAFA_SUCCESS;
}

}  // namespace shafa
