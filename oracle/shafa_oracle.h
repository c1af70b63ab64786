/*
 * shafa_oracle.h -- CPU restatement of the Shafa F/T/C/D per-block algorithms.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing on the product path (libshafa_cuda, the
 * `shafa` host binary, the python package) may include, link or call this.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs use it, and only as the checker / the reported CPU number.
 *
 * Parity status: PINNED.  Every function is checked (tests/test_oracle_*.py)
 * against fixtures produced by the unmodified reference binary built from
 * /root/reference/src by oracle/Makefile (target _ref), see
 * tests/golden/make_golden.py.
 *
 * All citations are relative to /root/reference/src/modules.
 */
#ifndef SHAFA_ORACLE_H
#define SHAFA_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NSYM 256
#define ORC_MAX_CODE 256 /* longest code string incl. NUL (t.c:161, c.c:18) */

/* error numbers = reference _modules_error (utils/errors.h:5-16) */
enum {
    ORC_OK = 0,
    ORC_LACK_OF_MEMORY = 2,
    ORC_FILE_UNRECOGNIZABLE = 4,
    ORC_FILE_STREAM_FAILED = 5,
    ORC_FILE_TOO_SMALL = 6
};

/* growable byte buffer returned by the whole-file functions */
typedef struct {
    uint8_t *data;
    uint64_t size;
    uint64_t cap;
} orc_buf;

void orc_buf_free(orc_buf *b);

/* ---------------- per-block primitives ---------------- */

/* f.c:29-55  greedy RLE of one block; out must hold 2*n+3 bytes. Returns size. */
uint64_t orc_rle_encode_block(const uint8_t *in, uint64_t n, uint8_t *out);

/* f.c:250-257  block-0 gate: 1 = keep RLE (R mode), 0 = N mode. float32 maths. */
int orc_rle_gate(uint64_t n0, uint64_t rle0, int force_rle);

/* f.c:63-79 */
void orc_histogram(const uint8_t *in, uint64_t n, uint64_t hist[ORC_NSYM]);

/* f.c:89-111  one block's 256 run-elided decimal fields (no "@n@", no "@0").
 * out must hold >= 21*256+256 bytes. Returns chars written (no NUL counted). */
size_t orc_freq_fields(const uint64_t hist[ORC_NSYM], char *out);

/* t.c:27-65  inverse of orc_freq_fields on a NUL-terminated field string. */
int orc_freq_parse(const char *text, uint64_t hist[ORC_NSYM]);

/* t.c:74-210,341-347  Shannon-Fano codes as '0'/'1' strings indexed by SYMBOL
 * (absent symbols -> ""). */
void orc_sf_codes(const uint64_t hist[ORC_NSYM], char codes[ORC_NSYM][ORC_MAX_CODE]);

/* same, packed: len[s] in bits, code[s] right-aligned in 64 bits.
 * Returns max length, or -1 if some code is longer than 64 bits. */
int orc_sf_codes_packed(const uint64_t hist[ORC_NSYM], uint8_t len[ORC_NSYM], uint64_t code[ORC_NSYM]);

/* t.c:350-361  256 code strings joined by ';' (no "@n@"). Returns chars written. */
size_t orc_cod_fields(const char codes[ORC_NSYM][ORC_MAX_CODE], char *out);

/* c.c:115-171 (parse side) split a NUL-terminated ';'-joined field string. */
int orc_cod_parse(const char *text, char codes[ORC_NSYM][ORC_MAX_CODE]);

/* c.c:52-83  MSB-first concatenation of codes, zero padded; out must hold
 * ceil(sum(len)/8) bytes. Returns z. */
uint64_t orc_sf_encode_block(const uint8_t *in, uint64_t n,
                             const char codes[ORC_NSYM][ORC_MAX_CODE], uint8_t *out);

/* exact payload size without encoding */
uint64_t orc_sf_encoded_size(const uint64_t hist[ORC_NSYM], const char codes[ORC_NSYM][ORC_MAX_CODE]);

/* d.c:422-551  prefix-tree walk, stops after n symbols. Returns ORC_OK or
 * ORC_FILE_UNRECOGNIZABLE where the reference would walk off the tree/payload. */
int orc_sf_decode_block(const uint8_t *payload, uint64_t z, uint64_t n,
                        const char codes[ORC_NSYM][ORC_MAX_CODE], uint8_t *out);

/* d.c:144-186  expand one RLE block. cap = capacity of out.  *final = bytes the
 * block expands to (computed even when it does not fit).  Returns ORC_OK,
 * ORC_FILE_UNRECOGNIZABLE when the block expands past 64 MiB + 1 KiB
 * (d.c:154-168) or is truncated inside a triple, ORC_LACK_OF_MEMORY if cap is
 * too small. */
int orc_rle_decode_block(const uint8_t *in, uint64_t n, uint8_t *out, uint64_t cap, uint64_t *final);

/* ---------------- whole-file container functions (in memory) ---------------- */

typedef struct {
    int rle_mode;         /* 1 = R (".rle" produced), 0 = N */
    uint64_t n_blocks;
    orc_buf rle;          /* ".rle" bytes (R mode)                      f.c:307 */
    orc_buf rle_freq;     /* ".rle.freq" text (R mode)                  f.c:289,312-314 */
    orc_buf freq;         /* ".freq" text (N mode or force_freq)        f.c:294,327-329 */
    uint64_t *rle_sizes;  /* per block, R mode */
} orc_f_result;

/* f.c:180-412 */
int orc_module_f(const uint8_t *in, uint64_t n, uint64_t block_size, int force_rle,
                 int force_freq, orc_f_result *res);
void orc_f_result_free(orc_f_result *res);

/* t.c:246-445  ".freq" text -> ".cod" text */
int orc_module_t(const uint8_t *freq_text, uint64_t freq_len, orc_buf *cod);

/* c.c:306-472  symbol stream + ".cod" text -> ".shaf" bytes */
int orc_module_c(const uint8_t *sym, uint64_t n, const uint8_t *cod_text, uint64_t cod_len, orc_buf *shaf);

/* d.c:628-834  ".shaf" + ".cod" -> symbol stream (and RLE-expanded if rle!=0) */
int orc_module_d_shaf(const uint8_t *shaf, uint64_t shaf_len, const uint8_t *cod_text,
                      uint64_t cod_len, int rle, orc_buf *out);

/* d.c:232-388  ".rle" + ".rle.freq" -> original */
int orc_module_d_rle(const uint8_t *rle, uint64_t rle_len, const uint8_t *freq_text,
                     uint64_t freq_len, orc_buf *out);

#ifdef __cplusplus
}
#endif
#endif
