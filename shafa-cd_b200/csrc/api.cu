// api.cu -- the extern "C" surface of libshafa_cuda (include/shafa_cuda.h):
// context, scratch arena, event-based kernel timing, the per-module device
// entry points and the whole-shard host-buffer calls.
#include <stdarg.h>
#include <new>
#include <stdlib.h>

#include <vector>

#include "internal.h"

using namespace shafa;

// ---------------------------------------------------------------------------------
// context plumbing
// ---------------------------------------------------------------------------------
namespace shafa {

int set_error(shafa_cuda_ctx *ctx, int code, const char *fmt, ...)
{
    if (ctx) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(ctx->err, sizeof ctx->err, fmt, ap);
        va_end(ap);
    }
    return code;
}

int cuda_fail(shafa_cuda_ctx *ctx, cudaError_t e, const char *what)
{
    int code = SHAFA_THREAD_CREATION_FAILED;
    if (e == cudaErrorMemoryAllocation) code = SHAFA_LACK_OF_MEMORY;
    set_error(ctx, code, "CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    cudaGetLastError();  // clear the sticky-less error state
    return code;
}

void *scratch_get(shafa_cuda_ctx *ctx, int slot, size_t bytes)
{
    if (bytes == 0) bytes = 16;
    if (ctx->scratch_cap[slot] >= bytes) return ctx->scratch[slot];
    if (ctx->scratch[slot]) {
        cudaStreamSynchronize(ctx->stream);
        cudaFree(ctx->scratch[slot]);
        ctx->scratch[slot] = nullptr;
        ctx->scratch_cap[slot] = 0;
    }
    size_t cap = bytes + bytes / 8 + 4096;
    cap = (cap + 255) & ~(size_t)255;
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, cap);
    if (e != cudaSuccess) {
        cuda_fail(ctx, e, "cudaMalloc(scratch)");
        return nullptr;
    }
    // Ragged streams are read with 16-byte vector loads that may reach into the padding behind the last valid byte (the
    // bytes are masked off before use): fresh memory is zeroed once so that those loads never see uninitialised data.
    cudaMemsetAsync(p, 0, cap, ctx->stream);
    ctx->scratch[slot] = p;
    ctx->scratch_cap[slot] = cap;
    return p;
}

// Pinned ring for per-call metadata.  A chunk stays valid until the ring wraps;
// wrapping synchronises the stream first so no in-flight copy is overwritten.
void *pinned_get(shafa_cuda_ctx *ctx, size_t bytes)
{
    bytes = (bytes + 63) & ~(size_t)63;
    if (bytes > ctx->pinned_cap / 2) {
        cudaStreamSynchronize(ctx->stream);
        if (ctx->pinned) cudaFreeHost(ctx->pinned);
        ctx->pinned = nullptr;
        size_t cap = bytes * 4 + (1u << 20);
        cudaError_t e = cudaMallocHost((void **)&ctx->pinned, cap);
        if (e != cudaSuccess) {
            ctx->pinned_cap = 0;
            cuda_fail(ctx, e, "cudaMallocHost(ring)");
            return nullptr;
        }
        ctx->pinned_cap = cap;
        ctx->pinned_pos = 0;
    }
    if (ctx->pinned_pos + bytes > ctx->pinned_cap) {
        cudaStreamSynchronize(ctx->stream);
        ctx->pinned_pos = 0;
    }
    void *p = ctx->pinned + ctx->pinned_pos;
    ctx->pinned_pos += bytes;
    return p;
}

shafa_cuda_ctx *worker_ctx(shafa_cuda_ctx *ctx, int i)
{
    if (i < 0 || i >= 4) return nullptr;
    if (!ctx->workers[i]) {
        shafa_cuda_ctx *w = nullptr;
        if (shafa_cuda_create(ctx->device, &w) != SHAFA_SUCCESS) return nullptr;
        ctx->workers[i] = w;
    }
    return ctx->workers[i];
}

uint32_t *misc_ptr(shafa_cuda_ctx *ctx, int which)
{
    uint32_t *base = (uint32_t *)scratch_get(ctx, SCR_MISC, MISC_WORDS * sizeof(uint32_t));
    return base ? base + which : nullptr;
}

static int stat_index(shafa_cuda_ctx *ctx, const char *name)
{
    for (size_t i = 0; i < ctx->stats.size(); i++)
        if (!strcmp(ctx->stats[i].name, name)) return (int)i;
    shafa_kernel_stat s;
    memset(&s, 0, sizeof s);
    strncpy(s.name, name, sizeof s.name - 1);
    ctx->stats.push_back(s);
    return (int)ctx->stats.size() - 1;
}

static cudaEvent_t event_take(shafa_cuda_ctx *ctx)
{
    if (!ctx->event_pool.empty()) {
        cudaEvent_t e = ctx->event_pool.back();
        ctx->event_pool.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}

void prof_begin(shafa_cuda_ctx *ctx, const char *name)
{
    ctx->launches++;
    const int si = stat_index(ctx, name);
    ctx->stats[si].launches++;
    if (!ctx->profile) return;
    PendingTiming p;
    p.stat = si;
    p.a = event_take(ctx);
    p.b = event_take(ctx);
    cudaEventRecord(p.a, ctx->stream);
    ctx->pending.push_back(p);
}

void prof_end(shafa_cuda_ctx *ctx)
{
    if (!ctx->profile || ctx->pending.empty()) return;
    cudaEventRecord(ctx->pending.back().b, ctx->stream);
}

static void prof_resolve(shafa_cuda_ctx *ctx)
{
    if (ctx->pending.empty()) return;
    cudaStreamSynchronize(ctx->stream);
    for (auto &p : ctx->pending) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) ctx->stats[p.stat].total_ms += ms;
        ctx->event_pool.push_back(p.a);
        ctx->event_pool.push_back(p.b);
    }
    ctx->pending.clear();
}

// read the device error word (synchronises) and translate it
static int finish_with_err(shafa_cuda_ctx *ctx, uint32_t *d_err, const char *what)
{
    uint32_t *h = (uint32_t *)pinned_get(ctx, 64);
    if (!h) return SHAFA_LACK_OF_MEMORY;
    SHAFA_CUDA_TRY(ctx, cudaMemcpyAsync(h, d_err, 4, cudaMemcpyDeviceToHost, ctx->stream));
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        cuda_fail(ctx, e, what);
        return SHAFA_THREAD_TERMINATION_FAILED;
    }
    const uint32_t bits = *h;
    ctx->last_err_bits = bits;
    if (!bits) return SHAFA_SUCCESS;
    if (bits & ERR_TIMEOUT) return set_error(ctx, SHAFA_THREAD_TERMINATION_FAILED, "%s: a device-side wait timed out (error bits 0x%x)", what, bits);
    if (bits & ERR_CAPACITY) return set_error(ctx, SHAFA_LACK_OF_MEMORY, "%s: output buffer too small", what);
    return set_error(ctx, SHAFA_FILE_UNRECOGNIZABLE, "%s: malformed input (error bits 0x%x)", what, bits);
}

static int err_reset(shafa_cuda_ctx *ctx, uint32_t **d_err)
{
    *d_err = misc_ptr(ctx, MISC_ERR);
    if (!*d_err) return SHAFA_LACK_OF_MEMORY;
    SHAFA_CUDA_TRY(ctx, cudaMemsetAsync(*d_err, 0, 4, ctx->stream));
    return SHAFA_SUCCESS;
}

}  // namespace shafa

#define CTX_ENTER(ctx)                                                                      \
    do {                                                                                    \
        if (!(ctx)) return SHAFA_OUTSIDE_MODULE;                                            \
        cudaError_t _e = cudaSetDevice((ctx)->device);                                      \
        if (_e != cudaSuccess) return shafa::cuda_fail((ctx), _e, "cudaSetDevice");         \
    } while (0)

extern "C" {

int shafa_cuda_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int shafa_cuda_create(int device, shafa_cuda_ctx **out)
{
    if (!out) return SHAFA_OUTSIDE_MODULE;
    *out = nullptr;
    int n = shafa_cuda_device_count();
    if (device < 0 || device >= n) return SHAFA_THREAD_CREATION_FAILED;  // no GPU: there is no CPU fallback
    shafa_cuda_ctx *ctx = new (std::nothrow) shafa_cuda_ctx();
    if (!ctx) return SHAFA_LACK_OF_MEMORY;
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return SHAFA_THREAD_CREATION_FAILED;
    }
    ctx->stream = ctx->own_stream;
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (ctx->sm_count <= 0) ctx->sm_count = 148;
    *out = ctx;
    return SHAFA_SUCCESS;
}

int shafa_cuda_destroy(shafa_cuda_ctx *ctx)
{
    if (!ctx) return SHAFA_SUCCESS;
    for (int i = 0; i < 4; i++)
        if (ctx->workers[i]) shafa_cuda_destroy(ctx->workers[i]);
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    prof_resolve(ctx);
    for (int i = 0; i < SCR_COUNT; i++)
        if (ctx->scratch[i]) cudaFree(ctx->scratch[i]);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    for (int i = 0; i < 2; i++) {
        if (ctx->bounce[i]) cudaFreeHost(ctx->bounce[i]);
        if (ctx->bounce_ev[i]) cudaEventDestroy(ctx->bounce_ev[i]);
    }
    for (auto e : ctx->event_pool) cudaEventDestroy(e);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return SHAFA_SUCCESS;
}

int shafa_cuda_set_stream(shafa_cuda_ctx *ctx, void *cuda_stream)
{
    if (!ctx) return SHAFA_OUTSIDE_MODULE;
    cudaStreamSynchronize(ctx->stream);
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return SHAFA_SUCCESS;
}

int shafa_cuda_synchronize(shafa_cuda_ctx *ctx)
{
    CTX_ENTER(ctx);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        cuda_fail(ctx, e, "cudaStreamSynchronize");
        return SHAFA_THREAD_TERMINATION_FAILED;
    }
    return SHAFA_SUCCESS;
}

const char *shafa_cuda_last_error(const shafa_cuda_ctx *ctx) { return ctx ? ctx->err : "no context"; }

void *shafa_cuda_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 16, cudaHostAllocPortable) != cudaSuccess) {  // portable: pinned for every device
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void shafa_cuda_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}
void *shafa_cuda_dev_alloc(shafa_cuda_ctx *ctx, size_t bytes)
{
    if (!ctx || cudaSetDevice(ctx->device) != cudaSuccess) return nullptr;
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes ? bytes : 16);
    if (e != cudaSuccess) {
        cuda_fail(ctx, e, "cudaMalloc");
        return nullptr;
    }
    cudaMemsetAsync(p, 0, bytes ? bytes : 16, ctx->stream);  // see scratch_get: padded vector loads behind ragged streams
    return p;
}
void shafa_cuda_dev_free(shafa_cuda_ctx *ctx, void *p)
{
    if (!ctx || !p) return;
    cudaSetDevice(ctx->device);
    cudaFree(p);
}
int shafa_cuda_memcpy_h2d(shafa_cuda_ctx *ctx, void *d_dst, const void *h_src, size_t bytes)
{
    CTX_ENTER(ctx);
    SHAFA_CUDA_TRY(ctx, cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return SHAFA_SUCCESS;
}
int shafa_cuda_memcpy_d2h(shafa_cuda_ctx *ctx, void *h_dst, const void *d_src, size_t bytes)
{
    CTX_ENTER(ctx);
    SHAFA_CUDA_TRY(ctx, cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return SHAFA_SUCCESS;
}

int shafa_cuda_profile_enable(shafa_cuda_ctx *ctx, int on)
{
    if (!ctx) return SHAFA_OUTSIDE_MODULE;
    if (!on) prof_resolve(ctx);
    ctx->profile = on != 0;
    return SHAFA_SUCCESS;
}
int shafa_cuda_profile_reset(shafa_cuda_ctx *ctx)
{
    if (!ctx) return SHAFA_OUTSIDE_MODULE;
    prof_resolve(ctx);
    for (auto &s : ctx->stats) {
        s.launches = 0;
        s.total_ms = 0;
    }
    ctx->launches = 0;
    return SHAFA_SUCCESS;
}
int shafa_cuda_profile_read(shafa_cuda_ctx *ctx, shafa_kernel_stat *out, int cap, int *n)
{
    if (!ctx) return SHAFA_OUTSIDE_MODULE;
    cudaSetDevice(ctx->device);
    prof_resolve(ctx);
    int k = 0;
    for (auto &s : ctx->stats) {
        if (!s.launches) continue;
        if (out && k < cap) out[k] = s;
        k++;
    }
    if (n) *n = k;
    return SHAFA_SUCCESS;
}
uint64_t shafa_cuda_launch_count(const shafa_cuda_ctx *ctx) { return ctx ? ctx->launches : 0; }

// ---------------------------------------------------------------------------------
// module-level device entry points
// ---------------------------------------------------------------------------------

int shafa_cuda_rle_gate(uint64_t n0, uint64_t rle0, int force_rle)
{
    // f.c:252-257: long difference, float32 quotient, compared with the double 0.05
    const long saved = (long)n0 - (long)rle0;
    const float ratio = (float)saved / (float)n0;
    return (ratio < 0.05 && !force_rle) ? 0 : 1;
}

int shafa_cuda_rle_encode_blocks(shafa_cuda_ctx *ctx, const uint8_t *d_in, uint64_t n, uint64_t block_size,
                                 uint8_t *d_rle, uint64_t rle_capacity, uint64_t *h_rle_sizes, uint64_t *d_hist)
{
    CTX_ENTER(ctx);
    uint32_t *d_err;
    int rc = err_reset(ctx, &d_err);
    if (rc) return rc;
    rc = launch_rle_encode(ctx, d_in, n, block_size, d_rle, rle_capacity, h_rle_sizes, d_hist, d_err);
    if (rc) return rc;
    return finish_with_err(ctx, d_err, "rle_encode_blocks");
}

int shafa_cuda_histogram_blocks(shafa_cuda_ctx *ctx, const uint8_t *d_in, const uint64_t *h_offsets, uint64_t n_blocks,
                                uint64_t *d_hist)
{
    CTX_ENTER(ctx);
    return launch_histogram(ctx, d_in, h_offsets, n_blocks, d_hist);
}

int shafa_cuda_sf_build_codes(shafa_cuda_ctx *ctx, const uint64_t *d_hist, uint64_t n_blocks, uint8_t *d_len,
                              uint64_t *d_code)
{
    CTX_ENTER(ctx);
    uint32_t *d_err;
    int rc = err_reset(ctx, &d_err);
    if (rc) return rc;
    rc = launch_sf_build(ctx, d_hist, n_blocks, d_len, d_code, d_err);
    if (rc) return rc;
    return finish_with_err(ctx, d_err, "sf_build_codes");
}

int shafa_cuda_sf_encode_blocks(shafa_cuda_ctx *ctx, const uint8_t *d_sym, const uint64_t *h_offsets, uint64_t nb,
                                const uint8_t *d_len, const uint64_t *d_code, const uint64_t *d_hist,
                                uint8_t *d_payload, uint64_t payload_capacity, uint64_t *h_payload_offsets,
                                uint64_t *h_z)
{
    CTX_ENTER(ctx);
    if (nb == 0) return SHAFA_SUCCESS;
    if (((uintptr_t)d_sym & 15) || ((uintptr_t)d_payload & 15))
        return set_error(ctx, SHAFA_FILE_STREAM_FAILED, "sf_encode: buffers must be 16-byte aligned");
    uint32_t *d_err;
    int rc = err_reset(ctx, &d_err);
    if (rc) return rc;
    if (!d_hist) {
        uint64_t *h = (uint64_t *)scratch_get(ctx, SCR_HIST, nb * 256 * sizeof(uint64_t));
        if (!h) return SHAFA_LACK_OF_MEMORY;
        rc = launch_histogram(ctx, d_sym, h_offsets, nb, h);
        if (rc) return rc;
        d_hist = h;
    }
    // sizes: [bits nb][z nb][off nb][summary 2] u64 + flags nb u32
    uint64_t *sz = (uint64_t *)scratch_get(ctx, SCR_SIZES, (3 * nb + 2) * sizeof(uint64_t) + nb * sizeof(uint32_t));
    if (!sz) return SHAFA_LACK_OF_MEMORY;
    uint64_t *d_bits = sz, *d_z = sz + nb, *d_off = sz + 2 * nb, *d_sum = sz + 3 * nb;
    uint32_t *d_flags = (uint32_t *)(sz + 3 * nb + 2);
    rc = launch_sf_sizes(ctx, d_hist, d_len, nb, d_bits, d_flags, d_z, d_off, d_sum);
    if (rc) return rc;
    uint64_t *h = (uint64_t *)pinned_get(ctx, (2 * nb + 2) * sizeof(uint64_t));
    if (!h) return SHAFA_LACK_OF_MEMORY;
    SHAFA_CUDA_TRY(ctx, cudaMemcpyAsync(h, d_z, (2 * nb + 2) * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    SHAFA_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(h_z, h, nb * sizeof(uint64_t));
    memcpy(h_payload_offsets, h + nb, nb * sizeof(uint64_t));
    const uint64_t total = h[2 * nb], flags = h[2 * nb + 1];
    if (flags & 2) return set_error(ctx, SHAFA_FILE_UNRECOGNIZABLE, "sf_encode: a symbol occurs in the data but has no code");
    if (total > payload_capacity) return set_error(ctx, SHAFA_LACK_OF_MEMORY, "sf_encode: payload needs %llu bytes, capacity %llu", (unsigned long long)total, (unsigned long long)payload_capacity);
    rc = launch_sf_encode(ctx, d_sym, h_offsets, nb, d_len, d_code, d_off, d_z, d_payload, (flags & 1) != 0, d_err);
    if (rc) return rc;
    return finish_with_err(ctx, d_err, "sf_encode_blocks");
}

int shafa_cuda_sf_decode_blocks(shafa_cuda_ctx *ctx, const uint8_t *d_payload, const uint64_t *h_payload_offsets,
                                const uint64_t *h_z, const uint64_t *h_n, uint64_t nb, const uint8_t *d_len,
                                const uint64_t *d_code, uint8_t *d_sym, uint64_t sym_capacity)
{
    CTX_ENTER(ctx);
    if (nb == 0) return SHAFA_SUCCESS;
    if ((uintptr_t)d_payload & 15) return set_error(ctx, SHAFA_FILE_STREAM_FAILED, "sf_decode: payload must be 16-byte aligned");
    uint64_t total = 0;
    for (uint64_t b = 0; b < nb; b++) total += h_n[b];
    if (total > sym_capacity) return set_error(ctx, SHAFA_LACK_OF_MEMORY, "sf_decode: output needs %llu bytes", (unsigned long long)total);
    uint32_t *d_err;
    int rc = err_reset(ctx, &d_err);
    if (rc) return rc;
    rc = launch_sf_decode(ctx, d_payload, h_payload_offsets, h_z, h_n, nb, d_len, d_code, d_sym, d_err, false, false);
    if (rc) return rc;
    rc = finish_with_err(ctx, d_err, "sf_decode_blocks");
    if (rc == SHAFA_FILE_UNRECOGNIZABLE && (ctx->last_err_bits & (ERR_SPECULATION | ERR_STITCH))) {
        // The single-pass decoder guessed that a sequence's end does not depend on where its first code starts,
        // checked the guess, and found it wrong somewhere (codes that do not self-synchronise): its output is not
        // trusted.  Decode again with the exact path.
        if ((rc = err_reset(ctx, &d_err))) return rc;
        // (the same for the exact path's own optimism: two stitch passes without host polling were not enough)
        rc = launch_sf_decode(ctx, d_payload, h_payload_offsets, h_z, h_n, nb, d_len, d_code, d_sym, d_err, true, true);
        if (rc) return rc;
        rc = finish_with_err(ctx, d_err, "sf_decode_blocks (exact path)");
    }
    return rc;
}

int shafa_cuda_rle_decode_blocks(shafa_cuda_ctx *ctx, const uint8_t *d_rle, const uint64_t *h_offsets, uint64_t nb,
                                 uint8_t *d_out, uint64_t out_capacity, uint64_t *h_final_sizes)
{
    CTX_ENTER(ctx);
    uint32_t *d_err;
    int rc = err_reset(ctx, &d_err);
    if (rc) return rc;
    rc = launch_rle_decode(ctx, d_rle, h_offsets, nb, d_out, out_capacity, h_final_sizes, d_err);
    if (rc) return rc;
    return finish_with_err(ctx, d_err, "rle_decode_blocks");
}

// ---------------------------------------------------------------------------------
// whole-shard calls on host buffers
// ---------------------------------------------------------------------------------

void shafa_cuda_encoded_free(shafa_cuda_encoded *e)
{
    if (!e) return;
    free(e->block_sizes);
    free(e->sym_sizes);
    free(e->hist);
    free(e->hist_orig);
    free(e->code_len);
    free(e->code_bits);
    free(e->payload_offsets);
    free(e->z);
    if (e->owns_rle == 2)
        free(e->rle);
    else if (e->owns_rle && e->rle)
        cudaFreeHost(e->rle);
    if (e->owns_payload == 2)
        free(e->payload);
    else if (e->owns_payload && e->payload)
        cudaFreeHost(e->payload);
    memset(e, 0, sizeof *e);
}

static inline uint64_t round16(uint64_t x) { return (x + 15) & ~15ull; }

// mode: -1 = probe block 0 (whole file), 0 / 1 = N / R decided elsewhere (shard without the file's block 0)
static int encode_host_impl(shafa_cuda_ctx *ctx, const uint8_t *h_in, uint64_t n, uint64_t block_size, int force_rle,
                            int force_freq, int mode, shafa_cuda_encoded *out)
{
    CTX_ENTER(ctx);
    if (!out || !h_in) return SHAFA_OUTSIDE_MODULE;
    if (block_size == 0) block_size = 524288;                     // utils/file.c:62 (the reference's default)
    if (block_size < 512) block_size = 512;                       // utils/file.c:63
    if (n >= (1ull << 48)) return SHAFA_LACK_OF_MEMORY;           // keeps every nb * 256 * 8 product far from wrapping
    if (block_size > 67108864ull) return SHAFA_FILE_STREAM_FAILED;
    if (mode < 0 && n < 1024) return SHAFA_FILE_TOO_SMALL;        // f.c:220,366 (a file-level rule)
    if (n == 0) return SHAFA_FILE_TOO_SMALL;
    const uint64_t nb = (n + block_size - 1) / block_size;        // utils/file.c:80-85
    uint8_t *user_rle = out->rle, *user_payload = out->payload;
    const uint64_t user_rle_cap = out->rle_capacity, user_payload_cap = out->payload_capacity;
    memset(out, 0, sizeof *out);
    out->n = n;
    out->block_size = block_size;
    out->n_blocks = nb;
    out->block_sizes = (uint64_t *)malloc(nb * sizeof(uint64_t));
    out->sym_sizes = (uint64_t *)malloc(nb * sizeof(uint64_t));
    out->hist = (uint64_t *)malloc(nb * 256 * sizeof(uint64_t));
    out->code_len = (uint8_t *)malloc(nb * 256);
    out->code_bits = (uint64_t *)malloc(nb * 256 * sizeof(uint64_t));
    out->payload_offsets = (uint64_t *)malloc(nb * sizeof(uint64_t));
    out->z = (uint64_t *)malloc(nb * sizeof(uint64_t));
    uint64_t *offsets = (uint64_t *)malloc((nb + 1) * sizeof(uint64_t));
    int rc = SHAFA_SUCCESS;
#define FAIL(code)      \
    do {                \
        rc = (code);    \
        goto done;      \
    } while (0)
    if (!out->block_sizes || !out->sym_sizes || !out->hist || !out->code_len || !out->code_bits ||
        !out->payload_offsets || !out->z || !offsets)
        FAIL(SHAFA_LACK_OF_MEMORY);
    {
        for (uint64_t b = 0; b < nb; b++) out->block_sizes[b] = (b == nb - 1) ? n - b * block_size : block_size;

        uint8_t *d_in = (uint8_t *)scratch_get(ctx, SCR_STAGE_A, round16(n) + 64);
        uint64_t *d_hist = (uint64_t *)scratch_get(ctx, SCR_STAGE_C, nb * 256 * sizeof(uint64_t));
        uint8_t *d_len = (uint8_t *)scratch_get(ctx, SCR_STAGE_D, nb * 256 * (sizeof(uint64_t) + 1));
        if (!d_in || !d_hist || !d_len) FAIL(SHAFA_LACK_OF_MEMORY);
        uint64_t *d_code = (uint64_t *)d_len;  // codes first (8-byte aligned), lengths after
        d_len = (uint8_t *)(d_code + nb * 256);
        if ((rc = shafa_cuda_memcpy_h2d(ctx, d_in, h_in, n))) goto done;

        // ---- F: block-0 probe decides R/N for the whole file (f.c:246-258) ----
        if (mode < 0) {
            uint64_t rle0 = 0;
            if ((rc = shafa_cuda_rle_encode_blocks(ctx, d_in, out->block_sizes[0], block_size, nullptr, 0, &rle0, nullptr))) goto done;
            out->rle_mode = shafa_cuda_rle_gate(out->block_sizes[0], rle0, force_rle);
        } else {
            out->rle_mode = mode != 0;
        }

        const uint8_t *d_sym = d_in;
        if (out->rle_mode) {
            const uint64_t cap = 2 * n + 3 * nb + 64;  // f.c:244 per block
            uint8_t *d_rle = (uint8_t *)scratch_get(ctx, SCR_STAGE_B, cap);
            if (!d_rle) FAIL(SHAFA_LACK_OF_MEMORY);
            if ((rc = shafa_cuda_rle_encode_blocks(ctx, d_in, n, block_size, d_rle, cap, out->sym_sizes, d_hist))) goto done;
            uint64_t tot = 0;
            for (uint64_t b = 0; b < nb; b++) tot += out->sym_sizes[b];
            out->rle_total = tot;
            if (user_rle && user_rle_cap >= tot) {
                out->rle = user_rle;
                out->rle_capacity = user_rle_cap;
            } else {
                if (cudaMallocHost((void **)&out->rle, tot + 16) != cudaSuccess) {
                    cudaGetLastError();
                    out->rle = nullptr;
                    FAIL(SHAFA_LACK_OF_MEMORY);
                }
                out->owns_rle = 1;
                out->rle_capacity = tot + 16;
            }
            if ((rc = shafa_cuda_memcpy_d2h(ctx, out->rle, d_rle, tot))) goto done;
            d_sym = d_rle;
            if (force_freq) {  // f.c:322-331: the original's histograms too
                out->hist_orig = (uint64_t *)malloc(nb * 256 * sizeof(uint64_t));
                uint64_t *d_ho = (uint64_t *)scratch_get(ctx, SCR_STAGE_E, nb * 256 * sizeof(uint64_t));
                if (!out->hist_orig || !d_ho) FAIL(SHAFA_LACK_OF_MEMORY);
                offsets[0] = 0;
                for (uint64_t b = 0; b < nb; b++) offsets[b + 1] = offsets[b] + out->block_sizes[b];
                if ((rc = shafa_cuda_histogram_blocks(ctx, d_in, offsets, nb, d_ho))) goto done;
                if ((rc = shafa_cuda_memcpy_d2h(ctx, out->hist_orig, d_ho, nb * 256 * sizeof(uint64_t)))) goto done;
            }
        } else {
            memcpy(out->sym_sizes, out->block_sizes, nb * sizeof(uint64_t));
        }
        offsets[0] = 0;
        for (uint64_t b = 0; b < nb; b++) offsets[b + 1] = offsets[b] + out->sym_sizes[b];
        if (!out->rle_mode) {
            if ((rc = shafa_cuda_histogram_blocks(ctx, d_in, offsets, nb, d_hist))) goto done;
        }
        if ((rc = shafa_cuda_memcpy_d2h(ctx, out->hist, d_hist, nb * 256 * sizeof(uint64_t)))) goto done;

        // ---- T ----
        if ((rc = shafa_cuda_sf_build_codes(ctx, d_hist, nb, d_len, d_code))) goto done;
        if ((rc = shafa_cuda_memcpy_d2h(ctx, out->code_bits, d_code, nb * 256 * sizeof(uint64_t)))) goto done;
        if ((rc = shafa_cuda_memcpy_d2h(ctx, out->code_len, d_len, nb * 256))) goto done;

        // ---- C ----  Fano codes average <= H + 1 <= 9 bits/symbol
        const uint64_t pcap = offsets[nb] + offsets[nb] / 8 + 32 * nb + 4096;
        uint8_t *d_payload = (uint8_t *)scratch_get(ctx, SCR_STAGE_F, pcap);
        if (!d_payload) FAIL(SHAFA_LACK_OF_MEMORY);
        if ((rc = shafa_cuda_sf_encode_blocks(ctx, d_sym, offsets, nb, d_len, d_code, d_hist, d_payload, pcap,
                                              out->payload_offsets, out->z)))
            goto done;
        const uint64_t ptotal = nb ? out->payload_offsets[nb - 1] + round16(out->z[nb - 1]) : 0;
        out->payload_total = ptotal;
        if (user_payload && user_payload_cap >= ptotal) {
            out->payload = user_payload;
            out->payload_capacity = user_payload_cap;
        } else {
            if (cudaMallocHost((void **)&out->payload, ptotal + 16) != cudaSuccess) {
                cudaGetLastError();
                out->payload = nullptr;
                FAIL(SHAFA_LACK_OF_MEMORY);
            }
            out->owns_payload = 1;
            out->payload_capacity = ptotal + 16;
        }
        if ((rc = shafa_cuda_memcpy_d2h(ctx, out->payload, d_payload, ptotal))) goto done;
        rc = shafa_cuda_synchronize(ctx);
    }
done:
#undef FAIL
    free(offsets);
    if (rc) {
        // keep caller-owned buffers out of the free
        if (!out->owns_rle) out->rle = nullptr;
        if (!out->owns_payload) out->payload = nullptr;
        shafa_cuda_encoded_free(out);
    }
    return rc;
}

int shafa_cuda_encode_host(shafa_cuda_ctx *ctx, const uint8_t *h_in, uint64_t n, uint64_t block_size, int force_rle,
                           int force_freq, shafa_cuda_encoded *out)
{
    if (out && out->payload && out->payload_capacity && out->rle && out->rle_capacity && n >= (64ull << 20) && ctx &&
        ctx->stream == ctx->own_stream)
        return encode_host_pipelined(&ctx, 1, h_in, n, block_size, force_rle, force_freq, -1, out);
    return encode_host_impl(ctx, h_in, n, block_size, force_rle, force_freq, -1, out);
}

int shafa_cuda_encode_shard_host(shafa_cuda_ctx *ctx, const uint8_t *h_in, uint64_t n, uint64_t block_size, int rle_mode,
                                 int force_freq, shafa_cuda_encoded *out)
{
    if (out && out->payload && out->payload_capacity && out->rle && out->rle_capacity && n >= (64ull << 20) && ctx &&
        ctx->stream == ctx->own_stream)
        return encode_host_pipelined(&ctx, 1, h_in, n, block_size, 0, force_freq, rle_mode ? 1 : 0, out);
    return encode_host_impl(ctx, h_in, n, block_size, 0, force_freq, rle_mode ? 1 : 0, out);
}

// ---- code tables with codes of 65..255 bits (sf_wide.cu) ---------------------------------------------------------------
int shafa_cuda_sf_encode_wide_host(shafa_cuda_ctx *ctx, const uint8_t *h_sym, const uint64_t *h_offsets, uint64_t nb,
                                   const uint8_t *h_len, const uint64_t *h_code4, uint8_t *h_payload, uint64_t payload_capacity,
                                   uint64_t *h_payload_offsets, uint64_t *h_z, uint64_t *h_payload_total)
{
    CTX_ENTER(ctx);
    if (h_payload_total) *h_payload_total = 0;
    if (nb == 0) return SHAFA_SUCCESS;
    if (nb > (1ull << 32)) return SHAFA_LACK_OF_MEMORY;
    const uint64_t lo = h_offsets[0], total = h_offsets[nb] - lo;
    std::vector<uint64_t> rel(nb + 1);
    for (uint64_t b = 0; b <= nb; b++) rel[b] = h_offsets[b] - lo;
    uint32_t *d_err;
    int rc = err_reset(ctx, &d_err);
    if (rc) return rc;
    uint8_t *d_sym = (uint8_t *)scratch_get(ctx, SCR_STAGE_A, round16(total) + 64);
    uint64_t *d_code4 = (uint64_t *)scratch_get(ctx, SCR_STAGE_D, nb * 256 * (4 * sizeof(uint64_t) + 1));
    uint64_t *d_hist = (uint64_t *)scratch_get(ctx, SCR_HIST, nb * 256 * sizeof(uint64_t));
    uint64_t *sz = (uint64_t *)scratch_get(ctx, SCR_SIZES, (3 * nb + 2) * sizeof(uint64_t) + nb * sizeof(uint32_t));
    if (!d_sym || !d_code4 || !d_hist || !sz) return SHAFA_LACK_OF_MEMORY;
    uint8_t *d_len = (uint8_t *)(d_code4 + nb * 256 * 4);
    uint64_t *d_bits = sz, *d_z = sz + nb, *d_off = sz + 2 * nb, *d_sum = sz + 3 * nb;
    uint32_t *d_flags = (uint32_t *)(sz + 3 * nb + 2);
    if ((rc = shafa_cuda_memcpy_h2d(ctx, d_sym, h_sym + lo, total))) return rc;
    if ((rc = shafa_cuda_memcpy_h2d(ctx, d_code4, h_code4, nb * 256 * 4 * sizeof(uint64_t)))) return rc;
    if ((rc = shafa_cuda_memcpy_h2d(ctx, d_len, h_len, nb * 256))) return rc;
    if ((rc = launch_histogram(ctx, d_sym, rel.data(), nb, d_hist))) return rc;
    if ((rc = launch_sf_sizes(ctx, d_hist, d_len, nb, d_bits, d_flags, d_z, d_off, d_sum))) return rc;
    uint64_t *h = (uint64_t *)pinned_get(ctx, (2 * nb + 2) * sizeof(uint64_t));
    if (!h) return SHAFA_LACK_OF_MEMORY;
    SHAFA_CUDA_TRY(ctx, cudaMemcpyAsync(h, d_z, (2 * nb + 2) * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    SHAFA_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(h_z, h, nb * sizeof(uint64_t));
    memcpy(h_payload_offsets, h + nb, nb * sizeof(uint64_t));
    const uint64_t ptotal = h[2 * nb], flags = h[2 * nb + 1];
    if (flags & 2) return set_error(ctx, SHAFA_FILE_UNRECOGNIZABLE, "sf_encode_wide: a symbol occurs in the data but has no code");
    if (h_payload_total) *h_payload_total = ptotal;
    if (ptotal > payload_capacity) return set_error(ctx, SHAFA_LACK_OF_MEMORY, "sf_encode_wide: payload needs %llu bytes, capacity %llu", (unsigned long long)ptotal, (unsigned long long)payload_capacity);
    uint8_t *d_payload = (uint8_t *)scratch_get(ctx, SCR_STAGE_F, round16(ptotal) + 64);
    if (!d_payload) return SHAFA_LACK_OF_MEMORY;
    if ((rc = launch_sf_encode_wide(ctx, d_sym, rel.data(), nb, d_len, d_code4, d_off, d_z, d_payload, d_err))) return rc;
    if ((rc = finish_with_err(ctx, d_err, "sf_encode_wide"))) return rc;
    if ((rc = shafa_cuda_memcpy_d2h(ctx, h_payload, d_payload, ptotal))) return rc;
    return shafa_cuda_synchronize(ctx);
}

int shafa_cuda_sf_decode_wide_host(shafa_cuda_ctx *ctx, const uint8_t *h_payload, const uint64_t *h_payload_offsets,
                                   const uint64_t *h_z, const uint64_t *h_n, uint64_t nb, const uint8_t *h_len,
                                   const uint64_t *h_code4, uint8_t *h_sym, uint64_t sym_capacity)
{
    CTX_ENTER(ctx);
    if (nb == 0) return SHAFA_SUCCESS;
    if (nb > (1ull << 32)) return SHAFA_LACK_OF_MEMORY;
    uint64_t extent = 0, stotal = 0;
    for (uint64_t b = 0; b < nb; b++) {
        const uint64_t e = h_payload_offsets[b] + h_z[b];
        if (e < h_payload_offsets[b]) return SHAFA_FILE_UNRECOGNIZABLE;
        if (e > extent) extent = e;
        stotal += h_n[b];
    }
    if (stotal > sym_capacity) return set_error(ctx, SHAFA_LACK_OF_MEMORY, "sf_decode_wide: output needs %llu bytes", (unsigned long long)stotal);
    uint32_t *d_err;
    int rc = err_reset(ctx, &d_err);
    if (rc) return rc;
    uint8_t *d_payload = (uint8_t *)scratch_get(ctx, SCR_STAGE_A, round16(extent) + 64);
    uint8_t *d_sym = (uint8_t *)scratch_get(ctx, SCR_STAGE_B, round16(stotal) + 64);
    uint64_t *d_code4 = (uint64_t *)scratch_get(ctx, SCR_STAGE_D, nb * 256 * (4 * sizeof(uint64_t) + 1));
    if (!d_payload || !d_sym || !d_code4) return SHAFA_LACK_OF_MEMORY;
    uint8_t *d_len = (uint8_t *)(d_code4 + nb * 256 * 4);
    if ((rc = shafa_cuda_memcpy_h2d(ctx, d_payload, h_payload, extent))) return rc;
    if ((rc = shafa_cuda_memcpy_h2d(ctx, d_code4, h_code4, nb * 256 * 4 * sizeof(uint64_t)))) return rc;
    if ((rc = shafa_cuda_memcpy_h2d(ctx, d_len, h_len, nb * 256))) return rc;
    if ((rc = launch_sf_decode_wide(ctx, d_payload, h_payload_offsets, h_z, h_n, nb, d_len, d_code4, d_sym, d_err))) return rc;
    if ((rc = finish_with_err(ctx, d_err, "sf_decode_wide"))) return rc;
    if ((rc = shafa_cuda_memcpy_d2h(ctx, h_sym, d_sym, stotal))) return rc;
    return shafa_cuda_synchronize(ctx);
}

int shafa_cuda_rle_decode_host(shafa_cuda_ctx *ctx, const uint8_t *h_rle, const uint64_t *h_offsets, uint64_t nb,
                               uint8_t *h_out, uint64_t out_capacity, uint64_t *h_out_size, uint64_t *h_final_sizes)
{
    CTX_ENTER(ctx);
    if (h_out_size) *h_out_size = 0;
    if (nb == 0) return SHAFA_SUCCESS;
    const uint64_t total = h_offsets[nb];
    uint8_t *d_rle = (uint8_t *)scratch_get(ctx, SCR_STAGE_A, round16(total) + 64);
    uint8_t *d_out = (uint8_t *)scratch_get(ctx, SCR_STAGE_B, round16(out_capacity) + 64);
    uint64_t *fin = h_final_sizes ? h_final_sizes : (uint64_t *)malloc(nb * sizeof(uint64_t));
    if (!d_rle || !d_out || !fin) {
        if (!h_final_sizes) free(fin);
        return SHAFA_LACK_OF_MEMORY;
    }
    memset(fin, 0, nb * sizeof(uint64_t));
    int rc = shafa_cuda_memcpy_h2d(ctx, d_rle, h_rle, total);
    if (!rc) rc = shafa_cuda_rle_decode_blocks(ctx, d_rle, h_offsets, nb, d_out, out_capacity, fin);
    uint64_t produced = 0;
    if (!rc || rc == SHAFA_LACK_OF_MEMORY)  // the sizes are exact in both cases (the caller may retry with a larger buffer)
        for (uint64_t b = 0; b < nb; b++) produced += fin[b];
    if (h_out_size) *h_out_size = produced;
    if (!rc) rc = shafa_cuda_memcpy_d2h(ctx, h_out, d_out, produced);
    if (!rc) rc = shafa_cuda_synchronize(ctx);
    if (!h_final_sizes) free(fin);
    return rc;
}

int shafa_cuda_decode_host(shafa_cuda_ctx *ctx, const uint8_t *h_payload, const uint64_t *h_payload_offsets,
                           const uint64_t *h_z, const uint64_t *h_n, uint64_t nb, const uint8_t *h_len,
                           const uint64_t *h_code, int rle_expand, uint8_t *h_out, uint64_t out_capacity,
                           uint64_t *h_out_size, uint64_t *h_final_sizes)
{
    CTX_ENTER(ctx);
    if (h_out_size) *h_out_size = 0;
    if (nb == 0) return SHAFA_SUCCESS;
    {
        uint64_t sym_total = 0;
        for (uint64_t b = 0; b < nb; b++) sym_total += h_n[b];
        if (nb >= 8 && sym_total >= (64ull << 20) && ctx->stream == ctx->own_stream)
            return decode_host_pipelined(&ctx, 1, h_payload, h_payload_offsets, h_z, h_n, nb, h_len, h_code, nullptr, rle_expand, h_out,
                                         out_capacity, h_out_size, h_final_sizes);
    }
    uint64_t ptotal = 0, stotal = 0;
    for (uint64_t b = 0; b < nb; b++) {
        const uint64_t e = h_payload_offsets[b] + round16(h_z[b]);
        if (e > ptotal) ptotal = e;
        stotal += h_n[b];
    }
    uint8_t *d_payload = (uint8_t *)scratch_get(ctx, SCR_STAGE_A, ptotal + 64);
    uint8_t *d_sym = (uint8_t *)scratch_get(ctx, SCR_STAGE_B, round16(stotal) + 64);
    uint8_t *d_tab = (uint8_t *)scratch_get(ctx, SCR_STAGE_D, nb * 256 * (sizeof(uint64_t) + 1));
    if (!d_payload || !d_sym || !d_tab) return SHAFA_LACK_OF_MEMORY;
    uint64_t *d_code = (uint64_t *)d_tab;
    uint8_t *d_len = (uint8_t *)(d_code + nb * 256);
    int rc;
    // the last slot may be shorter in the caller's buffer than its 16-byte rounding
    uint64_t copy_bytes = 0;
    for (uint64_t b = 0; b < nb; b++) {
        const uint64_t e = h_payload_offsets[b] + h_z[b];
        if (e > copy_bytes) copy_bytes = e;
    }
    if ((rc = shafa_cuda_memcpy_h2d(ctx, d_payload, h_payload, copy_bytes))) return rc;
    if ((rc = shafa_cuda_memcpy_h2d(ctx, d_code, h_code, nb * 256 * sizeof(uint64_t)))) return rc;
    if ((rc = shafa_cuda_memcpy_h2d(ctx, d_len, h_len, nb * 256))) return rc;
    if ((rc = shafa_cuda_sf_decode_blocks(ctx, d_payload, h_payload_offsets, h_z, h_n, nb, d_len, d_code, d_sym,
                                          round16(stotal) + 64)))
        return rc;
    if (!rle_expand) {
        if (stotal > out_capacity) return set_error(ctx, SHAFA_LACK_OF_MEMORY, "decode: output needs %llu bytes", (unsigned long long)stotal);
        if (h_final_sizes) memcpy(h_final_sizes, h_n, nb * sizeof(uint64_t));
        if ((rc = shafa_cuda_memcpy_d2h(ctx, h_out, d_sym, stotal))) return rc;
        if (h_out_size) *h_out_size = stotal;
        return shafa_cuda_synchronize(ctx);
    }
    uint64_t *offsets = (uint64_t *)malloc((nb + 1) * sizeof(uint64_t));
    uint64_t *fin = h_final_sizes ? h_final_sizes : (uint64_t *)malloc(nb * sizeof(uint64_t));
    uint8_t *d_out = (uint8_t *)scratch_get(ctx, SCR_STAGE_F, round16(out_capacity) + 64);
    if (!offsets || !fin || !d_out) {
        free(offsets);
        if (!h_final_sizes) free(fin);
        return SHAFA_LACK_OF_MEMORY;
    }
    offsets[0] = 0;
    for (uint64_t b = 0; b < nb; b++) offsets[b + 1] = offsets[b] + h_n[b];
    rc = shafa_cuda_rle_decode_blocks(ctx, d_sym, offsets, nb, d_out, out_capacity, fin);
    uint64_t produced = 0;
    for (uint64_t b = 0; b < nb; b++) produced += fin[b];
    if (h_out_size) *h_out_size = produced;
    if (!rc) rc = shafa_cuda_memcpy_d2h(ctx, h_out, d_out, produced);
    if (!rc) rc = shafa_cuda_synchronize(ctx);
    free(offsets);
    if (!h_final_sizes) free(fin);
    return rc;
}

}  // extern "C"
