/* util.h -- host-side helpers of the `shafa` program: paths, whole-file I/O into
 * pinned memory, a cursor over the text containers, the shared GPU context. */
#ifndef SHAFA_HOST_UTIL_H
#define SHAFA_HOST_UTIL_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

#include "modules.h"
#include "shafa_cuda.h"

#define RLE_EXT ".rle"
#define FREQ_EXT ".freq"
#define CODES_EXT ".cod"
#define SHAFA_EXT ".shaf"

enum { KIB_1 = 1024, KIB_64 = 65536, KIB_640 = 655360, MIB_8 = 8388608, MIB_64 = 67108864 };

/* extension helpers (same contracts as utils/extensions.h:19-35) */
bool check_ext(const char *path, const char *ext);
char *add_ext(const char *path, const char *ext);
char *rm_ext(const char *path);

/* The process-wide GPU set: devices from $SHAFA_DEVICES ("0-7", "0,2,5"; one file is then sharded over them by
 * block chunks, no collective), else $SHAFA_DEVICE, else device 0.  NULL if there is no usable GPU (no CPU fallback). */
shafa_cuda_multi *gpu_multi(void);
/* its first context (the per-module device calls run there) */
shafa_cuda_ctx *gpu(void);
/* start creating the contexts on a background thread: CUDA start-up then overlaps the reading of the input file */
void gpu_prefetch(void);
void gpu_release(void);
/* libshafa_cuda status -> _modules_error (same numbering) */
_modules_error gpu_status(int rc);

/* whole file -> host buffer (64 zeroed spare bytes after the end); pageable unless SHAFA_CLI_PINNED is set */
typedef struct {
    uint8_t *data;
    uint64_t size;
    bool pinned; /* data came from shafa_cuda_host_alloc (else from malloc / posix_memalign) */
} blob;
_modules_error blob_read(const char *path, blob *b);
/* large one-shot host buffers of the CLI: pageable by default, pinned with SHAFA_CLI_PINNED=1 */
void *cli_host_alloc(size_t bytes);
void cli_host_free(void *p);
void blob_free(blob *b);

/* cursor over "@...@..." text */
typedef struct {
    const uint8_t *p, *end;
} cursor;
bool cur_lit(cursor *c, char ch);
bool cur_u64(cursor *c, uint64_t *v);
/* span of bytes up to the next '@' (not copied); false if empty */
bool cur_field(cursor *c, const uint8_t **start, size_t *len);

/* one .cod block: 256 ';'-separated '0'/'1' strings (c.c:115-177, d.c:479-495) -> length + right-aligned bits.
 * Codes longer than 64 bits do not fit this form (_FILE_UNRECOGNIZABLE): callers look at cod_max_code_bits first and
 * take the wide form for such tables. */
_modules_error cod_parse_fields(const uint8_t *p, size_t n, uint8_t *len, uint64_t *code);
/* longest run of '0'/'1' characters in the block = its longest code */
size_t cod_max_code_bits(const uint8_t *p, size_t n);
/* the same for tables with codes of up to 255 bits (the reference's packer keeps 32 bytes per code, c.c:18): four
 * 64-bit words per code, word 0 the least significant (the library's shafa_cuda_sf_*_wide_host form) */
#define COD_WIDE_MAX_BITS 255
_modules_error cod_parse_fields_wide(const uint8_t *p, size_t n, uint8_t *len, uint64_t *code4);

/* Block counts come from untrusted "@N@<count>" headers and size every table that follows.  A .freq / .cod
 * block is at least "@d@" plus 255 ';' = 258 bytes, so a count above file_size / 258 is a lie
 * (_FILE_UNRECOGNIZABLE) -- checked BEFORE anything is allocated from it. */
bool block_count_plausible(uint64_t nb, uint64_t container_bytes);

/* module summaries (stdout), shared with the fused path */
void f_summary(uint64_t nb, const uint64_t *sizes, uint64_t n, const uint64_t *rle_sizes, double ms, const char *p_rle,
               const char *p_freq, const char *p_rle_freq);
void t_summary(uint64_t nb, const uint64_t *sizes, double ms, const char *p_cod);
void c_summary(uint64_t nb, const uint64_t *sizes, const uint64_t *z, double ms, const char *p_shaf);
/* fused.c: F, T and C of one invocation (shafa.c:293-298) in one pass over the data */
_modules_error shafa_encode_fused(char **path, bool force_rle, bool force_freq, unsigned long block_size);

/* container writers (containers.c) */
_modules_error write_freq_file(const char *path, char mode, uint64_t nb, const uint64_t *sizes, const uint64_t *hist);
_modules_error write_cod_file(const char *path, char mode, uint64_t nb, const uint64_t *sizes, const uint8_t *len, const uint64_t *code);
_modules_error write_shaf_file(const char *path, uint64_t nb, const uint64_t *z, const uint8_t *payload, const uint64_t *poff);
_modules_error write_all(const char *path, const uint8_t *data, uint64_t n);

double now_ms(void);

#endif
