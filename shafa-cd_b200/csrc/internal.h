// internal.h -- host-side plumbing shared by the .cu files of libshafa_cuda:
// the context (stream, scratch arena, event-based kernel timing) and the
// launchers each kernel file exports to api.cu.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/shafa_cuda.h"

namespace shafa {

enum ScratchSlot {
    SCR_PLAN = 0,   // per-block offsets / tile maps
    SCR_DESC,       // look-back descriptors
    SCR_DESC2,      // second chain
    SCR_MISC,       // counters, flags, small results
    SCR_HIST,       // internal histograms
    SCR_SIZES,      // per-block sizes / bit totals
    SCR_DEC_LUT,    // decode tables
    SCR_DEC_SUB,    // decode sub-sequence info
    SCR_DEC_SEQ,    // decode sequence info
    SCR_DEC_FB,     // decode sync: per-CTA tables of the all-starts fallback
    SCR_DEC_REC,    // fused decoder: per-sequence records
    SCR_DEC_CLS,    // decode sync: sequence numbering of the two classes, lean table slots
    SCR_DEC_LUT16,  // decode sync: 16-bit multi-code tables of the lean blocks
    SCR_STAGE_A,    // host API staging (input)
    SCR_STAGE_B,
    SCR_STAGE_C,
    SCR_STAGE_D,
    SCR_STAGE_E,
    SCR_STAGE_F,
    SCR_COUNT
};

struct PendingTiming {
    int stat;
    cudaEvent_t a, b;
};

}  // namespace shafa

struct shafa_cuda_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    char err[512] = {0};
    uint32_t last_err_bits = 0;  // device error word of the last finished call (shafa::ErrBits)
    void *scratch[shafa::SCR_COUNT] = {nullptr};
    size_t scratch_cap[shafa::SCR_COUNT] = {0};
    uint8_t *pinned = nullptr;  // pinned ring for per-call metadata
    size_t pinned_cap = 0;
    size_t pinned_pos = 0;
    bool profile = false;
    uint64_t launches = 0;
    std::vector<shafa_kernel_stat> stats;
    std::vector<shafa::PendingTiming> pending;
    std::vector<cudaEvent_t> event_pool;
    // worker contexts (own stream + scratch) of the pipelined host-buffer calls; owned by this context
    shafa_cuda_ctx *workers[4] = {nullptr, nullptr, nullptr, nullptr};
    // per-context (= per-device) cache of kernel attributes / occupancy, see shafa::OccId
    int occ[16] = {0};
    // pinned bounce buffers for copies from / to PAGEABLE host memory (host_pipeline.cu): 2 x bounce_cap bytes
    uint8_t *bounce[2] = {nullptr, nullptr};
    cudaEvent_t bounce_ev[2] = {nullptr, nullptr};
    size_t bounce_cap = 0;
};

namespace shafa {

int set_error(shafa_cuda_ctx *ctx, int code, const char *fmt, ...);
int cuda_fail(shafa_cuda_ctx *ctx, cudaError_t e, const char *what);
// grow-only device scratch; returns nullptr (and sets ctx->err) on failure
void *scratch_get(shafa_cuda_ctx *ctx, int slot, size_t bytes);
void *pinned_get(shafa_cuda_ctx *ctx, size_t bytes);
void prof_begin(shafa_cuda_ctx *ctx, const char *name);
void prof_end(shafa_cuda_ctx *ctx);

#define SHAFA_CUDA_TRY(ctx, expr)                                         \
    do {                                                                  \
        cudaError_t _e = (expr);                                          \
        if (_e != cudaSuccess) return shafa::cuda_fail((ctx), _e, #expr); \
    } while (0)

#define SHAFA_LAUNCH(ctx, name, kern, grid, block, smem, ...)               \
    do {                                                                    \
        shafa::prof_begin((ctx), (name));                                   \
        kern<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);      \
        shafa::prof_end((ctx));                                             \
    } while (0)

#define SHAFA_CHECK_LAUNCH(ctx)                                                        \
    do {                                                                               \
        cudaError_t _e = cudaGetLastError();                                           \
        if (_e != cudaSuccess) return shafa::cuda_fail((ctx), _e, "kernel launch");    \
    } while (0)

// ---- launchers (one per kernel file) ------------------------------------------
// Asynchronous on ctx->stream unless they return per-block sizes to the host.
// d_err is a device word collecting error bits:
enum ErrBits : uint32_t {
    ERR_TOTAL_TOO_BIG = 1,   // block total >= 2^31 (t.c int arithmetic)
    ERR_CODE_TOO_LONG = 2,   // code longer than 64 bits
    ERR_NO_CODE = 4,         // symbol occurs but has no code
    ERR_PAYLOAD_SHORT = 8,   // payload holds fewer symbols than announced
    ERR_BAD_WINDOW = 16,     // bit pattern that is no code
    ERR_CAPACITY = 32,       // output buffer too small
    ERR_RLE_TRUNCATED = 64,  // RLE block ends inside a triple
    ERR_SPECULATION = 128,   // fused SF decoder: a sequence's end depended on its start (decode again, exact path)
    ERR_STITCH = 512,        // exact SF decoder: sequence ends still moved in the second stitch pass (decode again, polling)
    ERR_TIMEOUT = 256        // fused SF decoder: a bounded device-side wait gave up (never seen; keeps a bug from hanging the GPU)
};
enum OccId { OCC_SF_ENCODE = 0, OCC_RLE_ENC = 1 /* +variant 0..3 */, OCC_RLE_DEC = 5, OCC_DEC_SYNC = 6, OCC_DEC_WRITE = 7, OCC_HIST = 8, OCC_DEC_FUSED = 9, OCC_DEC_LEAN = 10 };
enum MiscSlot { MISC_COUNTER = 0, MISC_FLAG = 16, MISC_ERR = 32, MISC_WORDS = 64 };  // uint32 indices into SCR_MISC
uint32_t *misc_ptr(shafa_cuda_ctx *ctx, int which);
shafa_cuda_ctx *worker_ctx(shafa_cuda_ctx *ctx, int i);  // lazily created child context i (0..3) on the same device

// host_pipeline.cu: chunked, overlapped versions of the whole-shard host calls over the worker contexts of one or
// several devices (ctxs[0] owns the file's block 0)
int encode_host_pipelined(shafa_cuda_ctx *const *ctxs, int n_ctx, const uint8_t *h_in, uint64_t n, uint64_t block_size, int force_rle,
                          int force_freq, int mode, shafa_cuda_encoded *out);
int decode_host_pipelined(shafa_cuda_ctx *const *ctxs, int n_ctx, const uint8_t *h_payload, const uint64_t *h_payload_offsets,
                          const uint64_t *h_z, const uint64_t *h_n, uint64_t nb, const uint8_t *h_len,
                          const uint64_t *h_code, const uint8_t *h_sym, int rle_expand, uint8_t *h_out, uint64_t out_capacity,
                          uint64_t *h_out_size, uint64_t *h_final_sizes);

// hist.cu
int launch_histogram(shafa_cuda_ctx *ctx, const uint8_t *d_in, const uint64_t *h_off, uint64_t nb, uint64_t *d_hist);

// sf_build.cu
int launch_sf_build(shafa_cuda_ctx *ctx, const uint64_t *d_hist, uint64_t nb, uint8_t *d_len, uint64_t *d_code,
                    uint32_t *d_err);

// sf_encode.cu
int launch_sf_sizes(shafa_cuda_ctx *ctx, const uint64_t *d_hist, const uint8_t *d_len, uint64_t nb, uint64_t *d_bits,
                    uint32_t *d_flags, uint64_t *d_z, uint64_t *d_payload_off, uint64_t *d_summary);
int launch_sf_encode(shafa_cuda_ctx *ctx, const uint8_t *d_sym, const uint64_t *h_off, uint64_t nb,
                     const uint8_t *d_len, const uint64_t *d_code, const uint64_t *d_payload_off, const uint64_t *d_z,
                     uint8_t *d_payload, bool any_long, uint32_t *d_err);

// sf_decode.cu
int launch_sf_decode(shafa_cuda_ctx *ctx, const uint8_t *d_payload, const uint64_t *h_payload_off,
                     const uint64_t *h_z, const uint64_t *h_n, uint64_t nb, const uint8_t *d_len,
                     const uint64_t *d_code, uint8_t *d_sym, uint32_t *d_err, bool force_exact, bool careful_stitch);

// sf_wide.cu: tables with codes of 65..255 bits (four 64-bit words per code, word 0 least significant)
int launch_sf_encode_wide(shafa_cuda_ctx *ctx, const uint8_t *d_sym, const uint64_t *h_off, uint64_t nb, const uint8_t *d_len,
                          const uint64_t *d_code4, const uint64_t *d_payload_off, const uint64_t *d_z, uint8_t *d_payload,
                          uint32_t *d_err);
int launch_sf_decode_wide(shafa_cuda_ctx *ctx, const uint8_t *d_payload, const uint64_t *h_payload_off, const uint64_t *h_z,
                          const uint64_t *h_n, uint64_t nb, const uint8_t *d_len, const uint64_t *d_code4, uint8_t *d_sym,
                          uint32_t *d_err);

// rle_encode.cu (synchronises: returns per-block sizes)
int launch_rle_encode(shafa_cuda_ctx *ctx, const uint8_t *d_in, uint64_t n, uint64_t block_size, uint8_t *d_rle,
                      uint64_t rle_capacity, uint64_t *h_rle_sizes, uint64_t *d_hist, uint32_t *d_err);

// rle_decode.cu (synchronises: returns per-block sizes)
int launch_rle_decode(shafa_cuda_ctx *ctx, const uint8_t *d_rle, const uint64_t *h_off, uint64_t nb, uint8_t *d_out,
                      uint64_t out_capacity, uint64_t *h_final_sizes, uint32_t *d_err);

}  // namespace shafa
