/*
 * shafa_cuda.h -- C ABI of libshafa_cuda, the B200 (sm_100a) implementation of
 * Shafa's per-block hot path.
 *
 * The reference keeps all of this work in static functions inside its four
 * modules and runs it one block per call / one pthread per block.  This
 * library replaces exactly those functions, batched: one call covers all blocks
 * of a file (or of one GPU's shard).  Text containers (.freq/.cod), "@n@"
 * framing and file I/O stay on the host side (shafa-cd_b200/host/).
 *
 *   reference function (under /root/reference/src/modules)      entry point here
 *   ---------------------------------------------------------   ----------------------------
 *   block_compression        f.c:29-55                           shafa_cuda_rle_encode_blocks
 *   RLE gate                 f.c:250-258                         shafa_cuda_rle_gate
 *   make_freq                f.c:63-79                           shafa_cuda_histogram_blocks
 *                                                                (+ fused in rle_encode_blocks)
 *   insert_sort/not_null/best_Division/sf_codes  t.c:74-210      shafa_cuda_sf_build_codes
 *   compress_to_buffer + binary_coding  c.c:52-83,91-237         shafa_cuda_sf_encode_blocks
 *   create_tree + shafa_block_decompressor  d.c:422-551          shafa_cuda_sf_decode_blocks
 *   rle_block_decompressor   d.c:116-197                         shafa_cuda_rle_decode_blocks
 *   multithread_create/wait  utils/multithread.c:126-223         (replaced by the batch calls;
 *                                                                 in-order output is the
 *                                                                 caller's plain loop)
 *
 * Conventions
 *   - every function returns a reference `_modules_error` number
 *     (utils/errors.h:5-16); 0 is success.  CUDA failures are folded into that
 *     enum (out of memory -> 2, launch failure -> 7, sync failure -> 8) and
 *     shafa_cuda_last_error() has the text.
 *   - `d_` pointers are device memory of the context's GPU, `h_` pointers are
 *     host memory.  Bulk data is device memory; per-block metadata is host
 *     memory.  The *_host functions take host memory only and do the copies.
 *   - device byte streams (d_in, d_sym, d_rle, d_payload) must be 16-byte
 *     aligned and readable up to the next 16-byte boundary past their end.
 *   - no C++ types, no exceptions, no torch types cross this boundary.
 *   - there is NO CPU fallback: without a CUDA device every entry point fails.
 */
#ifndef SHAFA_CUDA_H
#define SHAFA_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SHAFA_NSYM 256

/* utils/errors.h:5-16 */
enum {
    SHAFA_SUCCESS = 0,
    SHAFA_OUTSIDE_MODULE = 1,
    SHAFA_LACK_OF_MEMORY = 2,
    SHAFA_FILE_INACCESSIBLE = 3,
    SHAFA_FILE_UNRECOGNIZABLE = 4,
    SHAFA_FILE_STREAM_FAILED = 5,
    SHAFA_FILE_TOO_SMALL = 6,
    SHAFA_THREAD_CREATION_FAILED = 7,   /* also: kernel launch failed */
    SHAFA_THREAD_TERMINATION_FAILED = 8 /* also: stream synchronisation failed */
};

typedef struct shafa_cuda_ctx shafa_cuda_ctx;

/* ---- context ---------------------------------------------------------- */

int shafa_cuda_device_count(void);
/* one context per GPU; owns a stream and a scratch arena */
int shafa_cuda_create(int device, shafa_cuda_ctx **ctx);
int shafa_cuda_destroy(shafa_cuda_ctx *ctx);
/* run on the caller's cudaStream_t (e.g. torch's current stream); NULL restores the own stream */
int shafa_cuda_set_stream(shafa_cuda_ctx *ctx, void *cuda_stream);
int shafa_cuda_synchronize(shafa_cuda_ctx *ctx);
const char *shafa_cuda_last_error(const shafa_cuda_ctx *ctx);

/* memory helpers for a C host (pinned host memory, device memory, async copies on the ctx stream) */
void *shafa_cuda_host_alloc(size_t bytes);
void shafa_cuda_host_free(void *p);
void *shafa_cuda_dev_alloc(shafa_cuda_ctx *ctx, size_t bytes);
void shafa_cuda_dev_free(shafa_cuda_ctx *ctx, void *p);
int shafa_cuda_memcpy_h2d(shafa_cuda_ctx *ctx, void *d_dst, const void *h_src, size_t bytes);
int shafa_cuda_memcpy_d2h(shafa_cuda_ctx *ctx, void *h_dst, const void *d_src, size_t bytes);

/* kernel-level timing (CUDA events on the launch stream) and launch counting */
typedef struct {
    char name[48];
    uint64_t launches;
    double total_ms;
} shafa_kernel_stat;
int shafa_cuda_profile_enable(shafa_cuda_ctx *ctx, int on);
int shafa_cuda_profile_reset(shafa_cuda_ctx *ctx);
/* resolves pending events (synchronises), fills up to cap entries, *n = entries available */
int shafa_cuda_profile_read(shafa_cuda_ctx *ctx, shafa_kernel_stat *out, int cap, int *n);
uint64_t shafa_cuda_launch_count(const shafa_cuda_ctx *ctx);

/* ---- module F ---------------------------------------------------------- */

/* f.c:250-258: 1 = R mode (keep RLE), 0 = N mode.  float32 arithmetic. Host only. */
int shafa_cuda_rle_gate(uint64_t n0, uint64_t rle0, int force_rle);

/* f.c:29-55 for every block of d_in[0..n): blocks are block_size bytes (multiple
 * of 16), the last one shorter.  Output is the blocks' RLE bytes concatenated
 * (= the .rle file, f.c:307).  d_rle may be NULL (sizes only, e.g. the block-0
 * probe).  d_hist (B*256 u64, may be NULL) receives the histogram of the
 * EMITTED bytes per block (f.c:310), fused.  h_rle_sizes[B] is always filled.
 * SHAFA_LACK_OF_MEMORY if the output does not fit rle_capacity (worst case
 * 2*n + 3*B, f.c:244). */
int shafa_cuda_rle_encode_blocks(shafa_cuda_ctx *ctx, const uint8_t *d_in, uint64_t n,
                                 uint64_t block_size, uint8_t *d_rle, uint64_t rle_capacity,
                                 uint64_t *h_rle_sizes, uint64_t *d_hist);

/* f.c:63-79 for B ragged blocks: block b is d_in[h_offsets[b] .. h_offsets[b+1]).
 * d_hist: B*256 u64. */
int shafa_cuda_histogram_blocks(shafa_cuda_ctx *ctx, const uint8_t *d_in, const uint64_t *h_offsets,
                                uint64_t n_blocks, uint64_t *d_hist);

/* ---- module T ---------------------------------------------------------- */

/* t.c:74-210: per block, Shannon-Fano code of every symbol from its histogram.
 * d_len[b*256+s] = code length in bits (0 = symbol absent), d_code = the code,
 * right aligned.  SHAFA_FILE_UNRECOGNIZABLE if a block total is >= 2^31 (the
 * reference's int arithmetic, t.c:133-141) -- codes are then <= 52 bits. */
int shafa_cuda_sf_build_codes(shafa_cuda_ctx *ctx, const uint64_t *d_hist, uint64_t n_blocks,
                              uint8_t *d_len, uint64_t *d_code);

/* ---- module C ---------------------------------------------------------- */

/* c.c:52-83: per block, the codes of its symbols concatenated MSB-first, zero
 * padded to a byte; z_b = ceil(bits/8).  Block b's payload starts at
 * d_payload + h_payload_offsets[b] (16-byte aligned slots, filled in here).
 * d_hist may be NULL (computed internally).  Codes longer than 64 bits ->
 * SHAFA_FILE_UNRECOGNIZABLE at table level (see shafa_cuda_codes_from_text in
 * the host layer).  A symbol present in the data without a code ->
 * SHAFA_FILE_UNRECOGNIZABLE.  Not enough payload_capacity -> SHAFA_LACK_OF_MEMORY
 * with h_z filled. */
int shafa_cuda_sf_encode_blocks(shafa_cuda_ctx *ctx, const uint8_t *d_sym, const uint64_t *h_offsets,
                                uint64_t n_blocks, const uint8_t *d_len, const uint64_t *d_code,
                                const uint64_t *d_hist, uint8_t *d_payload, uint64_t payload_capacity,
                                uint64_t *h_payload_offsets, uint64_t *h_z);

/* ---- module D ---------------------------------------------------------- */

/* d.c:422-551: decode h_n[b] symbols of block b from d_payload+h_payload_offsets[b]
 * (z = h_z[b] bytes, 16-byte aligned start).  Output blocks are concatenated:
 * block b lands at d_sym + sum(h_n[0..b)).  A bit pattern that is no code, or a
 * payload too short for h_n[b] symbols -> SHAFA_FILE_UNRECOGNIZABLE. */
int shafa_cuda_sf_decode_blocks(shafa_cuda_ctx *ctx, const uint8_t *d_payload,
                                const uint64_t *h_payload_offsets, const uint64_t *h_z,
                                const uint64_t *h_n, uint64_t n_blocks, const uint8_t *d_len,
                                const uint64_t *d_code, uint8_t *d_sym, uint64_t sym_capacity);

/* d.c:116-197: expand B ragged RLE blocks; outputs concatenated in d_out.
 * h_final_sizes[B] always filled.  A block expanding past 64 MiB + 1 KiB or
 * truncated inside a triple -> SHAFA_FILE_UNRECOGNIZABLE (d.c:165-168).  Total
 * larger than out_capacity -> SHAFA_LACK_OF_MEMORY (nothing written past it). */
int shafa_cuda_rle_decode_blocks(shafa_cuda_ctx *ctx, const uint8_t *d_rle, const uint64_t *h_offsets,
                                 uint64_t n_blocks, uint8_t *d_out, uint64_t out_capacity,
                                 uint64_t *h_final_sizes);

/* ---- whole-shard calls on HOST buffers (what the host modules / e2e path use) ---- */

/* Everything modules F, T and C write for one input, as host arrays.
 * Zero the struct before the call.  The two bulk buffers may be preset by the
 * caller (ideally pinned memory from shafa_cuda_host_alloc, so the device->host
 * copies run at full PCIe speed): set rle/rle_capacity and/or
 * payload/payload_capacity; otherwise the library allocates pinned memory.
 * Everything the library allocated is released by shafa_cuda_encoded_free. */
typedef struct {
    int rle_mode;               /* 1: symbol stream = rle bytes, 0: = the input   */
    uint64_t n_blocks;
    uint64_t n;                 /* input bytes                                    */
    uint64_t block_size;
    uint64_t *block_sizes;      /* [B] original block sizes                       */
    uint8_t *rle;               /* R mode: concatenated RLE blocks (.rle)         */
    uint64_t rle_capacity;      /* in: capacity of a caller-owned rle buffer      */
    uint64_t rle_total;
    uint64_t *sym_sizes;        /* [B] symbol-stream block sizes (rle or orig)    */
    uint64_t *hist;             /* [B*256] histogram of the symbol stream (.freq / .rle.freq) */
    uint64_t *hist_orig;        /* [B*256] original histograms when force_freq in R mode, else NULL */
    uint8_t *code_len;          /* [B*256]                                        */
    uint64_t *code_bits;        /* [B*256]                                        */
    uint8_t *payload;           /* SF payloads, block b at payload_offsets[b]     */
    uint64_t payload_capacity;  /* in: capacity of a caller-owned payload buffer  */
    uint64_t payload_total;
    uint64_t *payload_offsets;  /* [B]                                            */
    uint64_t *z;                /* [B] payload bytes per block (.shaf "@z@")      */
    int owns_rle, owns_payload; /* set by the library (1: pinned, 2: malloc'd)    */
} shafa_cuda_encoded;

/* F -> T -> C for one host buffer (H2D, kernels, D2H inside). */
int shafa_cuda_encode_host(shafa_cuda_ctx *ctx, const uint8_t *h_in, uint64_t n, uint64_t block_size,
                           int force_rle, int force_freq, shafa_cuda_encoded *out);
/* The same for a shard of a file that does NOT contain the file's block 0 (multi-GPU: blocks are
 * independent, the only file-wide decision is block 0's R/N gate, f.c:246-258): rle_mode is given by
 * the owner of block 0 instead of being probed.  h_in must start on a block boundary of the file. */
int shafa_cuda_encode_shard_host(shafa_cuda_ctx *ctx, const uint8_t *h_in, uint64_t n, uint64_t block_size,
                                 int rle_mode, int force_freq, shafa_cuda_encoded *out);
void shafa_cuda_encoded_free(shafa_cuda_encoded *e);

/* D for host buffers: SF decode of B blocks (tables given as len/code arrays),
 * optionally followed by RLE expansion.  h_out must hold out_capacity bytes;
 * *h_out_size receives the bytes produced; h_final_sizes[B] (may be NULL) the
 * per-block output sizes. */
int shafa_cuda_decode_host(shafa_cuda_ctx *ctx, const uint8_t *h_payload,
                           const uint64_t *h_payload_offsets, const uint64_t *h_z,
                           const uint64_t *h_n, uint64_t n_blocks, const uint8_t *h_len,
                           const uint64_t *h_code, int rle_expand, uint8_t *h_out,
                           uint64_t out_capacity, uint64_t *h_out_size, uint64_t *h_final_sizes);

/* ---- code tables with codes longer than 64 bits ------------------------------
 * Module T cannot produce them, but a .cod file may carry them: the reference's packer keeps up to 32
 * bytes per code (c.c:18, 115-160) and its decoder builds its tree from the strings whatever their
 * length (d.c:422-504).  Here a code is FOUR 64-bit words (h_code4[B][256][4], word 0 the least
 * significant) of 1..255 bits (h_len); host buffers in and out, plain block-parallel kernels
 * (csrc/sf_wide.cu) -- the path exists for parity with every file the reference accepts, not for speed.
 *
 * c.c:52-83 (binary_coding): block b is h_sym[h_offsets[b] .. h_offsets[b+1]).  Payload slots are laid out
 * like shafa_cuda_sf_encode_blocks does (16-byte aligned, h_payload_offsets / h_z filled);
 * *h_payload_total = bytes used of h_payload.  Errors: a symbol that occurs without a code ->
 * SHAFA_FILE_UNRECOGNIZABLE; capacity -> SHAFA_LACK_OF_MEMORY (with *h_payload_total set). */
int shafa_cuda_sf_encode_wide_host(shafa_cuda_ctx *ctx, const uint8_t *h_sym, const uint64_t *h_offsets,
                                   uint64_t n_blocks, const uint8_t *h_len, const uint64_t *h_code4,
                                   uint8_t *h_payload, uint64_t payload_capacity,
                                   uint64_t *h_payload_offsets, uint64_t *h_z, uint64_t *h_payload_total);
/* d.c:514-551 (shafa_block_decompressor): h_n[b] symbols of block b into h_sym, blocks back to back.  A bit
 * pattern that is no code, or a payload that ends inside a code -> SHAFA_FILE_UNRECOGNIZABLE (the
 * reference dereferences NULL there, d.c:533-536).  RLE expansion, if wanted, is
 * shafa_cuda_rle_decode_host on the result. */
int shafa_cuda_sf_decode_wide_host(shafa_cuda_ctx *ctx, const uint8_t *h_payload,
                                   const uint64_t *h_payload_offsets, const uint64_t *h_z,
                                   const uint64_t *h_n, uint64_t n_blocks, const uint8_t *h_len,
                                   const uint64_t *h_code4, uint8_t *h_sym, uint64_t sym_capacity);

/* RLE-only decode (rle_decompress, d.c:232-388) for host buffers. */
int shafa_cuda_rle_decode_host(shafa_cuda_ctx *ctx, const uint8_t *h_rle, const uint64_t *h_offsets,
                               uint64_t n_blocks, uint8_t *h_out, uint64_t out_capacity,
                               uint64_t *h_out_size, uint64_t *h_final_sizes);

/* ---- one file sharded over N GPUs by one host process ----------------------------
 *
 * Replaces the reference's "one pthread per block, writes serialised by a join chain"
 * (utils/multithread.c:70-87,126-194) across GPUs: the blocks of ONE file are cut into chunks of
 * whole blocks, every chunk goes to whichever worker stream of whichever device is free, and the
 * results land compacted, in block order, in the caller's host buffers.  No collective: blocks are
 * independent, the only file-wide scalar is block 0's R/N decision (f.c:246-258), taken on
 * devices[0].  Outputs are byte-identical to the single-device calls for any device list. */
typedef struct shafa_cuda_multi shafa_cuda_multi;
int shafa_cuda_multi_create(const int *devices, int n_devices, shafa_cuda_multi **out);
int shafa_cuda_multi_destroy(shafa_cuda_multi *m);
int shafa_cuda_multi_device_count(const shafa_cuda_multi *m);
/* context i of the set (borrowed: e.g. for shafa_cuda_sf_build_codes on small tables) */
shafa_cuda_ctx *shafa_cuda_multi_ctx(shafa_cuda_multi *m, int i);
const char *shafa_cuda_multi_last_error(const shafa_cuda_multi *m);
uint64_t shafa_cuda_multi_launch_count(const shafa_cuda_multi *m);
/* as shafa_cuda_encode_host.  out->rle / out->payload may be preset (pinned memory from
 * shafa_cuda_host_alloc is pinned for every device); otherwise pageable buffers are allocated here
 * (owns_* = 2: released with free() by shafa_cuda_encoded_free). */
int shafa_cuda_encode_host_multi(shafa_cuda_multi *m, const uint8_t *h_in, uint64_t n, uint64_t block_size, int force_rle,
                                 int force_freq, shafa_cuda_encoded *out);
/* as shafa_cuda_decode_host / shafa_cuda_rle_decode_host */
int shafa_cuda_decode_host_multi(shafa_cuda_multi *m, const uint8_t *h_payload, const uint64_t *h_payload_offsets,
                                 const uint64_t *h_z, const uint64_t *h_n, uint64_t n_blocks, const uint8_t *h_len,
                                 const uint64_t *h_code, int rle_expand, uint8_t *h_out, uint64_t out_capacity,
                                 uint64_t *h_out_size, uint64_t *h_final_sizes);
int shafa_cuda_rle_decode_host_multi(shafa_cuda_multi *m, const uint8_t *h_rle, const uint64_t *h_offsets, uint64_t n_blocks,
                                     uint8_t *h_out, uint64_t out_capacity, uint64_t *h_out_size, uint64_t *h_final_sizes);

#ifdef __cplusplus
}
#endif
#endif /* SHAFA_CUDA_H */
