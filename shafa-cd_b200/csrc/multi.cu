// multi.cu -- ONE file sharded over N GPUs by ONE host process.
//
// The reference runs one pthread per block and serialises the writes with a join chain
// (utils/multithread.c:70-87).  Blocks are independent (RLE stops at the block edge f.c:38, one
// histogram / code table per block f.c:310 t.c:305-347, byte-aligned bitstream per block c.c:80), so
// here the blocks of the file are cut into chunks of whole blocks and every chunk goes to whichever
// worker stream of whichever GPU is free: H2D copy -> kernels -> D2H copy, results landing compacted
// in the caller's buffers in block order.  There is NO collective on the data path -- nothing is
// reduced or gathered between GPUs, so NCCL/NVLink are not used; the only file-wide scalar is block
// 0's R/N decision (f.c:246-258), taken on the first device before the other chunks start.
#include <stdlib.h>
#include <sys/mman.h>

#include <new>

#include "internal.h"

using namespace shafa;

struct shafa_cuda_multi {
    std::vector<shafa_cuda_ctx *> ctx;
    char err[600] = {0};
};

static int multi_fail(shafa_cuda_multi *m, int rc, const char *what)
{
    if (m && rc) snprintf(m->err, sizeof m->err, "%s: %s", what, m->ctx.empty() ? "no device context" : shafa_cuda_last_error(m->ctx[0]));
    return rc;
}

extern "C" {

int shafa_cuda_multi_create(const int *devices, int n_devices, shafa_cuda_multi **out)
{
    if (!out) return SHAFA_OUTSIDE_MODULE;
    *out = nullptr;
    if (!devices || n_devices < 1 || n_devices > 64) return SHAFA_OUTSIDE_MODULE;
    shafa_cuda_multi *m = new (std::nothrow) shafa_cuda_multi();
    if (!m) return SHAFA_LACK_OF_MEMORY;
    for (int i = 0; i < n_devices; i++) {
        shafa_cuda_ctx *c = nullptr;
        const int rc = shafa_cuda_create(devices[i], &c);  // the same device may be listed twice (two contexts on it)
        if (rc) {
            for (shafa_cuda_ctx *p : m->ctx) shafa_cuda_destroy(p);
            delete m;
            return rc;
        }
        m->ctx.push_back(c);
    }
    *out = m;
    return SHAFA_SUCCESS;
}

int shafa_cuda_multi_destroy(shafa_cuda_multi *m)
{
    if (!m) return SHAFA_SUCCESS;
    for (shafa_cuda_ctx *p : m->ctx) shafa_cuda_destroy(p);
    delete m;
    return SHAFA_SUCCESS;
}

int shafa_cuda_multi_device_count(const shafa_cuda_multi *m) { return m ? (int)m->ctx.size() : 0; }

shafa_cuda_ctx *shafa_cuda_multi_ctx(shafa_cuda_multi *m, int i) { return (m && i >= 0 && i < (int)m->ctx.size()) ? m->ctx[i] : nullptr; }

const char *shafa_cuda_multi_last_error(const shafa_cuda_multi *m) { return m ? m->err : "no multi-GPU context"; }

uint64_t shafa_cuda_multi_launch_count(const shafa_cuda_multi *m)
{
    uint64_t n = 0;
    if (m)
        for (shafa_cuda_ctx *c : m->ctx) {
            n += c->launches;
            for (int i = 0; i < 4; i++)
                if (c->workers[i]) n += c->workers[i]->launches;
        }
    return n;
}

int shafa_cuda_encode_host_multi(shafa_cuda_multi *m, const uint8_t *h_in, uint64_t n, uint64_t block_size, int force_rle,
                                 int force_freq, shafa_cuda_encoded *out)
{
    if (!m || m->ctx.empty() || !out) return SHAFA_OUTSIDE_MODULE;
    // output buffers: the caller's (ideally pinned, shafa_cuda_host_alloc) or pageable worst-case ones from here
    bool own_rle = false, own_pay = false;
    if (block_size == 0) block_size = 524288;  // utils/file.c:62
    if (block_size < 512) block_size = 512;
    const uint64_t nb = block_size ? (n + block_size - 1) / block_size : 0;
    if (!out->rle || !out->rle_capacity) {
        out->rle_capacity = 2 * n + 3 * nb + 64;  // f.c:244 per block (untouched pages cost nothing)
        out->rle = nullptr;
        if (posix_memalign((void **)&out->rle, 2u << 20, out->rle_capacity) != 0) return SHAFA_LACK_OF_MEMORY;
        madvise(out->rle, out->rle_capacity, MADV_HUGEPAGE);  // GiB-sized, written once: 512x fewer first-touch faults
        own_rle = true;
    }
    if (!out->payload || !out->payload_capacity) {
        out->payload_capacity = 2 * n + n / 8 + 32 * nb + 4096;  // Fano codes average <= H + 1 <= 9 bits per symbol of <= 2n + 3 symbols
        out->payload = nullptr;
        if (posix_memalign((void **)&out->payload, 2u << 20, out->payload_capacity) != 0) {
            if (own_rle) free(out->rle);
            return SHAFA_LACK_OF_MEMORY;
        }
        madvise(out->payload, out->payload_capacity, MADV_HUGEPAGE);
        own_pay = true;
    }
    uint8_t *rle = out->rle, *pay = out->payload;
    const int rc = encode_host_pipelined(m->ctx.data(), (int)m->ctx.size(), h_in, n, block_size, force_rle, force_freq, -1, out);
    if (rc) {
        if (own_rle) free(rle);
        if (own_pay) free(pay);
        return multi_fail(m, rc, "encode_host_multi");
    }
    if (own_rle) {
        if (out->rle_mode)
            out->owns_rle = 2;  // released with free() by shafa_cuda_encoded_free
        else
            free(rle);
    }
    if (own_pay) out->owns_payload = 2;
    return SHAFA_SUCCESS;
}

int shafa_cuda_decode_host_multi(shafa_cuda_multi *m, const uint8_t *h_payload, const uint64_t *h_payload_offsets, const uint64_t *h_z,
                                 const uint64_t *h_n, uint64_t n_blocks, const uint8_t *h_len, const uint64_t *h_code, int rle_expand,
                                 uint8_t *h_out, uint64_t out_capacity, uint64_t *h_out_size, uint64_t *h_final_sizes)
{
    if (!m || m->ctx.empty() || !h_payload) return SHAFA_OUTSIDE_MODULE;
    if (h_out_size) *h_out_size = 0;
    if (n_blocks == 0) return SHAFA_SUCCESS;
    return multi_fail(m,
                      decode_host_pipelined(m->ctx.data(), (int)m->ctx.size(), h_payload, h_payload_offsets, h_z, h_n, n_blocks, h_len, h_code,
                                            nullptr, rle_expand, h_out, out_capacity, h_out_size, h_final_sizes),
                      "decode_host_multi");
}

int shafa_cuda_rle_decode_host_multi(shafa_cuda_multi *m, const uint8_t *h_rle, const uint64_t *h_offsets, uint64_t n_blocks, uint8_t *h_out,
                                     uint64_t out_capacity, uint64_t *h_out_size, uint64_t *h_final_sizes)
{
    if (!m || m->ctx.empty() || !h_rle || !h_offsets) return SHAFA_OUTSIDE_MODULE;
    if (h_out_size) *h_out_size = 0;
    if (n_blocks == 0) return SHAFA_SUCCESS;
    std::vector<uint64_t> sizes(n_blocks);
    for (uint64_t b = 0; b < n_blocks; b++) {
        if (h_offsets[b + 1] < h_offsets[b]) return SHAFA_FILE_STREAM_FAILED;
        sizes[b] = h_offsets[b + 1] - h_offsets[b];
    }
    return multi_fail(m,
                      decode_host_pipelined(m->ctx.data(), (int)m->ctx.size(), nullptr, nullptr, nullptr, sizes.data(), n_blocks, nullptr, nullptr,
                                            h_rle + h_offsets[0], 1, h_out, out_capacity, h_out_size, h_final_sizes),
                      "rle_decode_host_multi");
}

}  // extern "C"
