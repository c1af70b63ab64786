// host_pipeline.cu -- chunked, overlapped versions of the whole-shard host-buffer calls.
//
// The reference overlaps per-block work with ordered output through its pthread join chain
// (utils/multithread.c:70-87).  Here the blocks of a shard are cut into chunks; worker threads,
// each with its own context (stream + scratch), run chunk after chunk through
//   H2D copy -> kernels -> D2H copy,
// so the copies of one chunk overlap the kernels and the opposite-direction copies of the others
// (PCIe is full duplex).  Results land directly at their final place in the caller's buffers:
// a chunk learns where its .rle bytes / payload slots / decoded bytes start from its predecessor
// through a small "base offset" chain, which is the only ordering between chunks (the analogue of
// the reference's write-in-block-order rule).
#include <stdlib.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include "internal.h"

namespace shafa {

namespace {

constexpr int PIPE_MAX_WORKERS = 4;  // = number of worker contexts a context can own

// tunables (environment overrides are for experiments; defaults were picked on a B200 / PCIe Gen5 host)
int pipe_workers()
{
    const char *e = getenv("SHAFA_PIPE_WORKERS");
    int w = e ? atoi(e) : 4;
    return w < 1 ? 1 : (w > PIPE_MAX_WORKERS ? PIPE_MAX_WORKERS : w);
}
uint64_t pipe_chunk_bytes()
{
    const char *e = getenv("SHAFA_PIPE_CHUNK_MB");
    const long mb = e ? atol(e) : 64;
    return (uint64_t)(mb < 1 ? 1 : mb) << 20;
}

// worker contexts of all participating devices, device-major.  One device: pipe_workers() of them; several
// devices (one file sharded over N GPUs): at most 16 host threads in total, at least two per
// device so that each GPU's copies overlap its kernels.
bool gather_workers(shafa_cuda_ctx *const *ctxs, int n_ctx, std::vector<shafa_cuda_ctx *> &out)
{
    int per = pipe_workers();
    if (n_ctx > 1) {
        const int cap = 16 / n_ctx < 2 ? 2 : 16 / n_ctx;
        if (per > cap) per = cap;
    }
    for (int c = 0; c < n_ctx; c++)
        for (int i = 0; i < per; i++) {
            shafa_cuda_ctx *w = worker_ctx(ctxs[c], i);
            if (!w) return false;
            out.push_back(w);
        }
    return true;
}

// ---- copies from / to PAGEABLE host memory ------------------------------------------------------------------
// cudaMemcpyAsync on pageable memory goes through the driver's own staging at a few GB/s and blocks the calling
// thread.  A caller that cannot pin its buffers (the `shafa` CLI reads a file once: pinning a GiB costs more than it
// saves) gets the same effect from two small pinned bounce buffers per worker: memcpy into one while the other is
// on the wire.  Several workers run side by side, so the host memcpy bandwidth adds up.
constexpr size_t BOUNCE_BYTES = 8u << 20;

bool is_pageable(const void *p)
{
    if (!p) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

bool bounce_ready(shafa_cuda_ctx *w)
{
    if (w->bounce_cap) return true;
    for (int i = 0; i < 2; i++) {
        if (cudaHostAlloc((void **)&w->bounce[i], BOUNCE_BYTES, cudaHostAllocDefault) != cudaSuccess ||
            cudaEventCreateWithFlags(&w->bounce_ev[i], cudaEventDisableTiming) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
    }
    w->bounce_cap = BOUNCE_BYTES;
    return true;
}

int copy_h2d(shafa_cuda_ctx *w, void *d_dst, const uint8_t *h_src, size_t bytes, bool pageable)
{
    if (!pageable || bytes < (1u << 20) || !bounce_ready(w)) return shafa_cuda_memcpy_h2d(w, d_dst, h_src, bytes);
    int k = 0;
    for (size_t off = 0; off < bytes; off += BOUNCE_BYTES, k ^= 1) {
        const size_t piece = bytes - off < BOUNCE_BYTES ? bytes - off : BOUNCE_BYTES;
        SHAFA_CUDA_TRY(w, cudaEventSynchronize(w->bounce_ev[k]));  // the copy that last used this buffer is done
        memcpy(w->bounce[k], h_src + off, piece);
        SHAFA_CUDA_TRY(w, cudaMemcpyAsync((uint8_t *)d_dst + off, w->bounce[k], piece, cudaMemcpyHostToDevice, w->stream));
        SHAFA_CUDA_TRY(w, cudaEventRecord(w->bounce_ev[k], w->stream));
    }
    return SHAFA_SUCCESS;
}

int copy_d2h(shafa_cuda_ctx *w, uint8_t *h_dst, const void *d_src, size_t bytes, bool pageable)
{
    if (!pageable || bytes < (1u << 20) || !bounce_ready(w)) return shafa_cuda_memcpy_d2h(w, h_dst, d_src, bytes);
    // piece i travels into buffer i & 1 while piece i - 1 is copied out of the other one
    size_t prev_off = 0, prev_piece = 0;
    int k = 0;
    for (size_t off = 0; off < bytes; off += BOUNCE_BYTES, k ^= 1) {
        const size_t piece = bytes - off < BOUNCE_BYTES ? bytes - off : BOUNCE_BYTES;
        SHAFA_CUDA_TRY(w, cudaMemcpyAsync(w->bounce[k], (const uint8_t *)d_src + off, piece, cudaMemcpyDeviceToHost, w->stream));
        SHAFA_CUDA_TRY(w, cudaEventRecord(w->bounce_ev[k], w->stream));
        if (prev_piece) {
            SHAFA_CUDA_TRY(w, cudaEventSynchronize(w->bounce_ev[k ^ 1]));
            memcpy(h_dst + prev_off, w->bounce[k ^ 1], prev_piece);
        }
        prev_off = off;
        prev_piece = piece;
    }
    if (prev_piece) {
        SHAFA_CUDA_TRY(w, cudaEventSynchronize(w->bounce_ev[k ^ 1]));
        memcpy(h_dst + prev_off, w->bounce[k ^ 1], prev_piece);
    }
    return SHAFA_SUCCESS;
}

struct Chain {
    std::mutex m;
    std::condition_variable cv;
    std::vector<uint64_t> base;
    std::vector<char> ready;
    bool dead = false;
    explicit Chain(size_t n) : base(n + 1, 0), ready(n + 1, 0) { ready[0] = 1; }
    bool wait(size_t c, uint64_t *v)
    {
        std::unique_lock<std::mutex> lk(m);
        cv.wait(lk, [&] { return dead || ready[c]; });
        if (!ready[c]) return false;
        *v = base[c];
        return true;
    }
    void publish(size_t c, uint64_t v)
    {
        {
            std::lock_guard<std::mutex> lk(m);
            base[c] = v;
            ready[c] = 1;
        }
        cv.notify_all();
    }
    void abort()
    {
        {
            std::lock_guard<std::mutex> lk(m);
            dead = true;
        }
        cv.notify_all();
    }
};

inline uint64_t round16(uint64_t x) { return (x + 15) & ~15ull; }

struct Chunks {
    uint64_t per;    // blocks per chunk
    uint64_t count;  // number of chunks
};

Chunks plan_chunks(uint64_t nb, uint64_t bytes, size_t n_workers)
{
    const uint64_t chunk = pipe_chunk_bytes();
    uint64_t want = (bytes + chunk - 1) / chunk;
    if (want < 2 * (uint64_t)n_workers) want = 2 * (uint64_t)n_workers;  // enough chunks to keep both copy directions busy
    if (want > nb) want = nb;
    if (want < 1) want = 1;
    Chunks c;
    c.per = (nb + want - 1) / want;
    c.count = (nb + c.per - 1) / c.per;
    return c;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------
// encode
// ---------------------------------------------------------------------------------------------------
int encode_host_pipelined(shafa_cuda_ctx *const *ctxs, int n_ctx, const uint8_t *h_in, uint64_t n, uint64_t block_size, int force_rle,
                          int force_freq, int mode, shafa_cuda_encoded *out)
{
    if (!ctxs || n_ctx < 1 || !ctxs[0] || !out || !h_in) return SHAFA_OUTSIDE_MODULE;
    shafa_cuda_ctx *const ctx = ctxs[0];  // block 0's owner: the R/N decision of the whole file is taken here (f.c:246-258)
    if (cudaSetDevice(ctx->device) != cudaSuccess) return SHAFA_THREAD_CREATION_FAILED;
    if (block_size == 0) block_size = 524288;                     // utils/file.c:62 (the reference's default)
    if (block_size < 512) block_size = 512;                       // utils/file.c:63
    if (n >= (1ull << 48)) return SHAFA_LACK_OF_MEMORY;           // keeps every nb * 256 * 8 product far from wrapping
    if (block_size > 67108864ull) return SHAFA_FILE_STREAM_FAILED;
    if (mode < 0 && n < 1024) return SHAFA_FILE_TOO_SMALL;  // f.c:220,366
    if (n == 0) return SHAFA_FILE_TOO_SMALL;
    const uint64_t nb = (n + block_size - 1) / block_size;  // utils/file.c:80-85
    uint8_t *const user_rle = out->rle, *const user_payload = out->payload;
    const uint64_t rle_cap = out->rle_capacity, pay_cap = out->payload_capacity;
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
    if (force_freq) out->hist_orig = (uint64_t *)malloc(nb * 256 * sizeof(uint64_t));
    int rc = SHAFA_SUCCESS;
    if (!out->block_sizes || !out->sym_sizes || !out->hist || !out->code_len || !out->code_bits || !out->payload_offsets ||
        !out->z || (force_freq && !out->hist_orig))
        rc = SHAFA_LACK_OF_MEMORY;
    if (!rc) {
        for (uint64_t b = 0; b < nb; b++) out->block_sizes[b] = (b == nb - 1) ? n - b * block_size : block_size;
        // ---- block 0 decides R/N for the whole file (f.c:246-258) ----
        if (mode < 0) {
            const uint64_t n0 = out->block_sizes[0];
            uint8_t *d0 = (uint8_t *)scratch_get(ctx, SCR_STAGE_A, round16(n0) + 64);
            uint64_t rle0 = 0;
            if (!d0)
                rc = SHAFA_LACK_OF_MEMORY;
            else if (!(rc = shafa_cuda_memcpy_h2d(ctx, d0, h_in, n0)) &&
                     !(rc = shafa_cuda_rle_encode_blocks(ctx, d0, n0, block_size, nullptr, 0, &rle0, nullptr)))
                mode = shafa_cuda_rle_gate(n0, rle0, force_rle);
        }
    }
    if (rc) {
        shafa_cuda_encoded_free(out);
        return rc;
    }
    out->rle_mode = mode != 0;
    if (!out->rle_mode && out->hist_orig) {  // N mode: the one .freq IS the original's (f.c:322-331)
        free(out->hist_orig);
        out->hist_orig = nullptr;
    }
    out->rle = out->rle_mode ? user_rle : nullptr;
    out->rle_capacity = rle_cap;
    out->payload = user_payload;
    out->payload_capacity = pay_cap;

    std::vector<shafa_cuda_ctx *> workers;
    if (!gather_workers(ctxs, n_ctx, workers)) {
        out->rle = nullptr;
        out->payload = nullptr;
        shafa_cuda_encoded_free(out);
        return set_error(ctx, SHAFA_LACK_OF_MEMORY, "pipelined encode: worker contexts could not be created");
    }
    const Chunks ch = plan_chunks(nb, n, workers.size());
    Chain rle_chain(ch.count), pay_chain(ch.count);
    std::atomic<uint64_t> next{0};
    std::atomic<int> err{0};
    const bool rle_mode = out->rle_mode, want_orig = out->hist_orig != nullptr;
    const bool in_pageable = is_pageable(h_in), rle_pageable = is_pageable(out->rle), pay_pageable = is_pageable(out->payload);

    std::mutex err_m;
    std::string err_text;
    auto fail_from = [&](shafa_cuda_ctx *w, int code) {
        int expected = 0;
        if (err.compare_exchange_strong(expected, code)) {
            std::lock_guard<std::mutex> lk(err_m);
            err_text = w ? shafa_cuda_last_error(w) : "worker context could not be created";  // the FIRST failing worker's text
        }
        rle_chain.abort();
        pay_chain.abort();
    };

    std::vector<uint32_t> chunks_done(workers.size(), 0);
    auto work = [&](int widx) {
        shafa_cuda_ctx *w = workers[widx];
        if (cudaSetDevice(w->device) != cudaSuccess) return fail_from(nullptr, SHAFA_THREAD_CREATION_FAILED);
        auto fail = [&](int code) { fail_from(w, code); };
        for (;;) {
            const uint64_t c = next.fetch_add(1);
            if (c >= ch.count || err.load()) return;
            chunks_done[widx]++;
            const uint64_t b0 = c * ch.per, b1 = (b0 + ch.per < nb) ? b0 + ch.per : nb, cb = b1 - b0;
            const uint64_t lo = b0 * block_size, hi = (b1 * block_size < n) ? b1 * block_size : n, cn = hi - lo;
            int r;
            uint8_t *d_in = (uint8_t *)scratch_get(w, SCR_STAGE_A, round16(cn) + 64);
            uint64_t *d_hist = (uint64_t *)scratch_get(w, SCR_STAGE_C, cb * 256 * sizeof(uint64_t));
            uint64_t *d_code = (uint64_t *)scratch_get(w, SCR_STAGE_D, cb * 256 * (sizeof(uint64_t) + 1));
            if (!d_in || !d_hist || !d_code) return fail(SHAFA_LACK_OF_MEMORY);
            uint8_t *d_len = (uint8_t *)(d_code + cb * 256);
            if ((r = copy_h2d(w, d_in, h_in + lo, cn, in_pageable))) return fail(r);
            std::vector<uint64_t> offsets(cb + 1, 0);
            const uint8_t *d_sym = d_in;
            if (rle_mode) {
                const uint64_t cap = 2 * cn + 3 * cb + 64;  // f.c:244 per block
                uint8_t *d_rle = (uint8_t *)scratch_get(w, SCR_STAGE_B, cap);
                if (!d_rle) return fail(SHAFA_LACK_OF_MEMORY);
                if ((r = shafa_cuda_rle_encode_blocks(w, d_in, cn, block_size, d_rle, cap, out->sym_sizes + b0, d_hist))) return fail(r);
                uint64_t tot = 0;
                for (uint64_t b = 0; b < cb; b++) tot += out->sym_sizes[b0 + b];
                uint64_t base;
                if (!rle_chain.wait(c, &base)) return;
                if (base + tot > rle_cap) return fail(SHAFA_LACK_OF_MEMORY);
                rle_chain.publish(c + 1, base + tot);
                if ((r = copy_d2h(w, out->rle + base, d_rle, tot, rle_pageable))) return fail(r);
                d_sym = d_rle;
                if (want_orig) {  // f.c:322-331: the original's histograms too
                    uint64_t *d_ho = (uint64_t *)scratch_get(w, SCR_STAGE_E, cb * 256 * sizeof(uint64_t));
                    if (!d_ho) return fail(SHAFA_LACK_OF_MEMORY);
                    for (uint64_t b = 0; b < cb; b++) offsets[b + 1] = offsets[b] + out->block_sizes[b0 + b];
                    if ((r = shafa_cuda_histogram_blocks(w, d_in, offsets.data(), cb, d_ho))) return fail(r);
                    if ((r = shafa_cuda_memcpy_d2h(w, out->hist_orig + b0 * 256, d_ho, cb * 256 * sizeof(uint64_t)))) return fail(r);
                }
            } else {
                memcpy(out->sym_sizes + b0, out->block_sizes + b0, cb * sizeof(uint64_t));
            }
            for (uint64_t b = 0; b < cb; b++) offsets[b + 1] = offsets[b] + out->sym_sizes[b0 + b];
            if (!rle_mode && (r = shafa_cuda_histogram_blocks(w, d_in, offsets.data(), cb, d_hist))) return fail(r);
            if ((r = shafa_cuda_memcpy_d2h(w, out->hist + b0 * 256, d_hist, cb * 256 * sizeof(uint64_t)))) return fail(r);
            // ---- T ----
            if ((r = shafa_cuda_sf_build_codes(w, d_hist, cb, d_len, d_code))) return fail(r);
            if ((r = shafa_cuda_memcpy_d2h(w, out->code_bits + b0 * 256, d_code, cb * 256 * sizeof(uint64_t)))) return fail(r);
            if ((r = shafa_cuda_memcpy_d2h(w, out->code_len + b0 * 256, d_len, cb * 256))) return fail(r);
            // ---- C ----  Fano codes average <= H + 1 <= 9 bits/symbol
            const uint64_t pcap = offsets[cb] + offsets[cb] / 8 + 32 * cb + 4096;
            uint8_t *d_payload = (uint8_t *)scratch_get(w, SCR_STAGE_F, pcap);
            if (!d_payload) return fail(SHAFA_LACK_OF_MEMORY);
            if ((r = shafa_cuda_sf_encode_blocks(w, d_sym, offsets.data(), cb, d_len, d_code, d_hist, d_payload, pcap,
                                                 out->payload_offsets + b0, out->z + b0)))
                return fail(r);
            const uint64_t ptot = out->payload_offsets[b1 - 1] + round16(out->z[b1 - 1]);
            uint64_t pbase;
            if (!pay_chain.wait(c, &pbase)) return;
            if (pbase + ptot > pay_cap) return fail(SHAFA_LACK_OF_MEMORY);
            pay_chain.publish(c + 1, pbase + ptot);
            for (uint64_t b = b0; b < b1; b++) out->payload_offsets[b] += pbase;
            if ((r = copy_d2h(w, out->payload + pbase, d_payload, ptot, pay_pageable))) return fail(r);
            if ((r = shafa_cuda_synchronize(w))) return fail(r);
        }
    };

    {
        // chunks go to whichever worker is free, devices interleaved (worker i of every device before worker i + 1)
        std::vector<std::thread> threads;
        const size_t nw = ch.count < workers.size() ? (size_t)ch.count : workers.size();
        const size_t per = workers.size() / (size_t)n_ctx;
        for (size_t i = 0; i < nw; i++) threads.emplace_back(work, (int)((i % (size_t)n_ctx) * per + i / (size_t)n_ctx));
        for (auto &t : threads) t.join();
    }
    cudaSetDevice(ctx->device);
    rc = err.load();
    if (rc) {
        // copies into the caller's buffers may still be in flight on the streams of workers that bailed out
        for (shafa_cuda_ctx *w : workers) {
            cudaSetDevice(w->device);
            cudaStreamSynchronize(w->stream);
        }
        cudaSetDevice(ctx->device);
        set_error(ctx, rc, "pipelined encode: %s", err_text.c_str());
        out->rle = nullptr;  // caller-owned buffers stay with the caller
        out->payload = nullptr;
        shafa_cuda_encoded_free(out);
        return rc;
    }
    if (getenv("SHAFA_VERBOSE")) {
        fprintf(stderr, "pipelined encode: %llu chunks of %llu block(s); per worker (device:chunks):", (unsigned long long)ch.count, (unsigned long long)ch.per);
        for (size_t i = 0; i < workers.size(); i++) fprintf(stderr, " %d:%u", workers[i]->device, chunks_done[i]);
        fprintf(stderr, "\n");
    }
    out->rle_total = rle_mode ? rle_chain.base[ch.count] : 0;
    out->payload_total = pay_chain.base[ch.count];
    out->owns_rle = out->owns_payload = 0;
    return SHAFA_SUCCESS;
}

// ---------------------------------------------------------------------------------------------------
// decode
// ---------------------------------------------------------------------------------------------------
// h_payload == nullptr: "-d r" -- the symbol stream itself is given (h_sym, blocks at the running sums of h_n) and only
// expanded (rle_block_decompressor, d.c:116-197)
int decode_host_pipelined(shafa_cuda_ctx *const *ctxs, int n_ctx, const uint8_t *h_payload, const uint64_t *h_payload_offsets,
                          const uint64_t *h_z, const uint64_t *h_n, uint64_t nb, const uint8_t *h_len,
                          const uint64_t *h_code, const uint8_t *h_sym, int rle_expand, uint8_t *h_out, uint64_t out_capacity,
                          uint64_t *h_out_size, uint64_t *h_final_sizes)
{
    if (!ctxs || n_ctx < 1 || !ctxs[0]) return SHAFA_OUTSIDE_MODULE;
    shafa_cuda_ctx *const ctx = ctxs[0];
    if (cudaSetDevice(ctx->device) != cudaSuccess) return SHAFA_THREAD_CREATION_FAILED;
    if (h_out_size) *h_out_size = 0;
    uint64_t sym_total = 0;
    std::vector<uint64_t> sym_off(nb + 1, 0);
    for (uint64_t b = 0; b < nb; b++) {
        sym_total += h_n[b];
        sym_off[b + 1] = sym_total;
    }
    std::vector<shafa_cuda_ctx *> workers;
    if (!gather_workers(ctxs, n_ctx, workers)) return set_error(ctx, SHAFA_LACK_OF_MEMORY, "pipelined decode: worker contexts could not be created");
    const Chunks ch = plan_chunks(nb, sym_total, workers.size());
    const bool in_pageable = is_pageable(h_payload ? h_payload : h_sym), out_pageable = is_pageable(h_out);
    Chain out_chain(ch.count);
    std::atomic<uint64_t> next{0};
    std::atomic<int> err{0};
    std::vector<uint64_t> fin_local;
    uint64_t *fin = h_final_sizes;
    if (!fin) {
        fin_local.assign(nb, 0);
        fin = fin_local.data();
    }
    std::mutex err_m;
    std::string err_text;
    auto fail_from = [&](shafa_cuda_ctx *w, int code) {
        int expected = 0;
        if (err.compare_exchange_strong(expected, code)) {
            std::lock_guard<std::mutex> lk(err_m);
            err_text = w ? shafa_cuda_last_error(w) : "worker context could not be created";
        }
        out_chain.abort();
    };
    auto work = [&](int widx) {
        shafa_cuda_ctx *w = workers[widx];
        if (cudaSetDevice(w->device) != cudaSuccess) return fail_from(nullptr, SHAFA_THREAD_CREATION_FAILED);
        auto fail = [&](int code) { fail_from(w, code); };
        for (;;) {
            const uint64_t c = next.fetch_add(1);
            if (c >= ch.count || err.load()) return;
            const uint64_t b0 = c * ch.per, b1 = (b0 + ch.per < nb) ? b0 + ch.per : nb, cb = b1 - b0;
            int r;
            if (!h_payload) {  // RLE expansion only: the chunk's slice of the symbol stream goes up as it is
                const uint64_t slo = sym_off[b0], stot = sym_off[b1] - sym_off[b0];
                std::vector<uint64_t> offsets(cb + 1, 0);
                for (uint64_t b = 0; b < cb; b++) offsets[b + 1] = offsets[b] + h_n[b0 + b];
                uint8_t *d_sym = (uint8_t *)scratch_get(w, SCR_STAGE_B, round16(stot) + 64);
                if (!d_sym) return fail(SHAFA_LACK_OF_MEMORY);
                if ((r = copy_h2d(w, d_sym, h_sym + slo, stot, in_pageable))) return fail(r);
                uint64_t cap = 0;
                for (uint64_t b = 0; b < cb; b++) {
                    const uint64_t s = h_n[b0 + b];
                    cap += (s <= 65536 ? 65536 : s <= 655360 ? 655360 : s <= 8388608 ? 8388608 : 67108864) + 1024;  // d.c:129-136
                }
                uint8_t *d_out = (uint8_t *)scratch_get(w, SCR_STAGE_F, round16(cap) + 64);
                if (!d_out) return fail(SHAFA_LACK_OF_MEMORY);
                r = shafa_cuda_rle_decode_blocks(w, d_sym, offsets.data(), cb, d_out, cap, fin + b0);
                if (r == SHAFA_LACK_OF_MEMORY) {
                    cap = 0;
                    for (uint64_t b = 0; b < cb; b++) cap += fin[b0 + b];
                    d_out = (uint8_t *)scratch_get(w, SCR_STAGE_F, round16(cap) + 64);
                    if (!d_out) return fail(SHAFA_LACK_OF_MEMORY);
                    r = shafa_cuda_rle_decode_blocks(w, d_sym, offsets.data(), cb, d_out, cap, fin + b0);
                }
                if (r) return fail(r);
                uint64_t produced = 0;
                for (uint64_t b = 0; b < cb; b++) produced += fin[b0 + b];
                uint64_t base;
                if (!out_chain.wait(c, &base)) return;
                out_chain.publish(c + 1, base + produced);
                if (base + produced <= out_capacity) {
                    if ((r = copy_d2h(w, h_out + base, (const uint8_t *)d_out, produced, out_pageable))) return fail(r);
                    if ((r = shafa_cuda_synchronize(w))) return fail(r);
                }
                continue;
            }
            // payload slice of this chunk: slots are laid out in block order by the encoder / the .shaf reader;
            // be general and take min/max over the chunk's blocks
            uint64_t plo = ~0ull, phi = 0, stot = 0;
            for (uint64_t b = b0; b < b1; b++) {
                if (h_payload_offsets[b] < plo) plo = h_payload_offsets[b];
                const uint64_t e = h_payload_offsets[b] + h_z[b];
                if (e > phi) phi = e;
                stot += h_n[b];
            }
            if (phi < plo) phi = plo;
            std::vector<uint64_t> poff(cb), offsets(cb + 1, 0);
            for (uint64_t b = 0; b < cb; b++) {
                poff[b] = h_payload_offsets[b0 + b] - plo;
                offsets[b + 1] = offsets[b] + h_n[b0 + b];
            }
            uint8_t *d_payload = (uint8_t *)scratch_get(w, SCR_STAGE_A, round16(phi - plo) + 64);
            uint8_t *d_sym = (uint8_t *)scratch_get(w, SCR_STAGE_B, round16(stot) + 64);
            uint64_t *d_code = (uint64_t *)scratch_get(w, SCR_STAGE_D, cb * 256 * (sizeof(uint64_t) + 1));
            if (!d_payload || !d_sym || !d_code) return fail(SHAFA_LACK_OF_MEMORY);
            uint8_t *d_len = (uint8_t *)(d_code + cb * 256);
            if ((r = copy_h2d(w, d_payload, h_payload + plo, phi - plo, in_pageable))) return fail(r);
            if ((r = shafa_cuda_memcpy_h2d(w, d_code, h_code + b0 * 256, cb * 256 * sizeof(uint64_t)))) return fail(r);
            if ((r = shafa_cuda_memcpy_h2d(w, d_len, h_len + b0 * 256, cb * 256))) return fail(r);
            if ((r = shafa_cuda_sf_decode_blocks(w, d_payload, poff.data(), h_z + b0, h_n + b0, cb, d_len, d_code, d_sym,
                                                 round16(stot) + 64)))
                return fail(r);
            uint64_t produced = stot;
            const uint8_t *d_result = d_sym;
            if (rle_expand) {
                // expansion of this chunk: sizes are not known in advance -> the reference's own first guess per block
                // (d.c:129-136), then the exact size if that was too small
                uint64_t cap = 0;
                for (uint64_t b = 0; b < cb; b++) {
                    const uint64_t s = h_n[b0 + b];
                    cap += (s <= 65536 ? 65536 : s <= 655360 ? 655360 : s <= 8388608 ? 8388608 : 67108864) + 1024;
                }
                uint8_t *d_out = (uint8_t *)scratch_get(w, SCR_STAGE_F, round16(cap) + 64);
                if (!d_out) return fail(SHAFA_LACK_OF_MEMORY);
                r = shafa_cuda_rle_decode_blocks(w, d_sym, offsets.data(), cb, d_out, cap, fin + b0);
                if (r == SHAFA_LACK_OF_MEMORY) {
                    cap = 0;
                    for (uint64_t b = 0; b < cb; b++) cap += fin[b0 + b];
                    d_out = (uint8_t *)scratch_get(w, SCR_STAGE_F, round16(cap) + 64);
                    if (!d_out) return fail(SHAFA_LACK_OF_MEMORY);
                    r = shafa_cuda_rle_decode_blocks(w, d_sym, offsets.data(), cb, d_out, cap, fin + b0);
                }
                if (r) return fail(r);
                produced = 0;
                for (uint64_t b = 0; b < cb; b++) produced += fin[b0 + b];
                d_result = d_out;
            } else {
                for (uint64_t b = 0; b < cb; b++) fin[b0 + b] = h_n[b0 + b];
            }
            uint64_t base;
            if (!out_chain.wait(c, &base)) return;
            out_chain.publish(c + 1, base + produced);
            if (base + produced <= out_capacity) {
                if ((r = copy_d2h(w, h_out + base, d_result, produced, out_pageable))) return fail(r);
                if ((r = shafa_cuda_synchronize(w))) return fail(r);
            }  // else: keep going so that the total size is reported (the caller may retry with a larger buffer)
        }
    };
    {
        std::vector<std::thread> threads;
        const size_t nw = ch.count < workers.size() ? (size_t)ch.count : workers.size();
        const size_t per = workers.size() / (size_t)n_ctx;
        for (size_t i = 0; i < nw; i++) threads.emplace_back(work, (int)((i % (size_t)n_ctx) * per + i / (size_t)n_ctx));
        for (auto &t : threads) t.join();
    }
    cudaSetDevice(ctx->device);
    const int rc = err.load();
    if (rc) {
        for (shafa_cuda_ctx *w : workers) {
            cudaSetDevice(w->device);
            cudaStreamSynchronize(w->stream);
        }
        cudaSetDevice(ctx->device);
        set_error(ctx, rc, "pipelined decode: %s", err_text.c_str());
        return rc;
    }
    const uint64_t total = out_chain.base[ch.count];
    if (h_out_size) *h_out_size = total;
    if (total > out_capacity) return set_error(ctx, SHAFA_LACK_OF_MEMORY, "decode: output needs %llu bytes", (unsigned long long)total);
    return SHAFA_SUCCESS;
}

}  // namespace shafa
