/*
 * shafa_oracle.c -- CPU restatement of the Shafa F/T/C/D block algorithms.
 *
 * TEST INFRASTRUCTURE ONLY (see shafa_oracle.h).  It is the checker for the
 * CUDA path and the "port" CPU baseline; it is never linked into the product.
 * Parity status: PINNED against the unmodified reference binary
 * (oracle/Makefile target _ref; fixtures in tests/golden/).
 *
 * Citations are relative to /root/reference/src/modules.
 */
#include "shafa_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ */
/* small growable buffer                                               */
/* ------------------------------------------------------------------ */

static int buf_reserve(orc_buf *b, uint64_t extra)
{
    if (b->size + extra <= b->cap) return 1;
    uint64_t cap = b->cap ? b->cap : 4096;
    while (cap < b->size + extra) cap += cap / 2 + 4096;
    uint8_t *p = (uint8_t *)realloc(b->data, cap);
    if (!p) return 0;
    b->data = p;
    b->cap = cap;
    return 1;
}

static int buf_put(orc_buf *b, const void *src, uint64_t n)
{
    if (!buf_reserve(b, n)) return 0;
    memcpy(b->data + b->size, src, n);
    b->size += n;
    return 1;
}

static int buf_printf_u64(orc_buf *b, const char *pre, uint64_t v, const char *post)
{
    char tmp[64];
    int k = snprintf(tmp, sizeof tmp, "%s%llu%s", pre, (unsigned long long)v, post);
    return buf_put(b, tmp, (uint64_t)k);
}

void orc_buf_free(orc_buf *b)
{
    free(b->data);
    b->data = NULL;
    b->size = b->cap = 0;
}

/* ------------------------------------------------------------------ */
/* Module F                                                            */
/* ------------------------------------------------------------------ */

/* f.c:29-55.  At position i take the equal-byte run starting there, capped at
 * 255 and at the block end (f.c:38).  A run of >= 4, or any run of 0x00, is
 * written as {0, sym, len} and skipped whole (f.c:40-46); otherwise ONE
 * literal is written and the scan moves one byte (f.c:47-52). */
uint64_t orc_rle_encode_block(const uint8_t *in, uint64_t n, uint8_t *out)
{
    uint64_t i = 0, o = 0;
    while (i < n) {
        const uint8_t sym = in[i];
        const uint64_t limit = (n - i < 255) ? (n - i) : 255;
        uint64_t run = 1;
        while (run < limit && in[i + run] == sym) run++;
        if (run >= 4 || sym == 0) {
            out[o++] = 0;
            out[o++] = sym;
            out[o++] = (uint8_t)run;
            i += run;
        } else {
            out[o++] = sym;
            i += 1;
        }
    }
    return o;
}

/* f.c:250-257.  compression is a signed long difference, the ratio a float32
 * quotient of two float32 conversions, compared (after promotion) with the
 * double literal 0.05. */
int orc_rle_gate(uint64_t n0, uint64_t rle0, int force_rle)
{
    const long saved = (long)n0 - (long)rle0;
    const float ratio = (float)saved / (float)n0;
    if (ratio < 0.05 && !force_rle) return 0;
    return 1;
}

/* f.c:63-79 */
void orc_histogram(const uint8_t *in, uint64_t n, uint64_t hist[ORC_NSYM])
{
    memset(hist, 0, sizeof(uint64_t) * ORC_NSYM);
    for (uint64_t i = 0; i < n; i++) hist[in[i]]++;
}

/* f.c:94-110.  The count is printed when it differs from the previous
 * symbol's count (always for symbol 0); a ';' follows every symbol but 255. */
size_t orc_freq_fields(const uint64_t hist[ORC_NSYM], char *out)
{
    char *p = out;
    for (int s = 0; s < ORC_NSYM; s++) {
        if (s == 0 || hist[s] != hist[s - 1])
            p += sprintf(p, "%llu", (unsigned long long)hist[s]);
        if (s != ORC_NSYM - 1) *p++ = ';';
    }
    *p = '\0';
    return (size_t)(p - out);
}

/* t.c:27-65.  256 ';'-separated fields; an empty field repeats the previous
 * value (t.c:50-53); the first field must be a number (t.c:31). */
int orc_freq_parse(const char *text, uint64_t hist[ORC_NSYM])
{
    const char *p = text;
    for (int s = 0; s < ORC_NSYM; s++) {
        if (*p >= '0' && *p <= '9') {
            uint64_t v = 0;
            while (*p >= '0' && *p <= '9') v = v * 10 + (uint64_t)(*p++ - '0');
            hist[s] = v;
        } else {
            if (s == 0) return ORC_FILE_UNRECOGNIZABLE;
            hist[s] = hist[s - 1];
        }
        if (s != ORC_NSYM - 1) {
            if (*p != ';') return ORC_FILE_UNRECOGNIZABLE;
            p++;
        }
    }
    return (*p == '\0') ? ORC_OK : ORC_FILE_UNRECOGNIZABLE;
}

/* ------------------------------------------------------------------ */
/* Module T                                                            */
/* ------------------------------------------------------------------ */

/* t.c:130-152.  Walk d = first, first+1, ... accumulating the left weight and
 * keep going only while |2*left - total| strictly decreases; the split is the
 * last d that improved.  Same int arithmetic range as the reference. */
static int sf_split(const uint64_t *w, int first, int last)
{
    long long total = 0;
    for (int i = first; i <= last; i++) total += (long long)w[i];
    long long best = total, left = 0;
    int d = first;
    for (;;) {
        left += (long long)w[d];
        long long dif = 2 * left - total;
        if (dif < 0) dif = -dif;
        if (dif < best) {
            best = dif;
            d++;
        } else
            break;
    }
    return d - 1;
}

/* t.c:74-104 (stable descending order, symbol -> rank), t.c:202-210 (last
 * non-zero rank), t.c:182-195 (recursive split; left gets '0', right '1'). */
void orc_sf_codes(const uint64_t hist[ORC_NSYM], char codes[ORC_NSYM][ORC_MAX_CODE])
{
    int order[ORC_NSYM]; /* rank -> symbol */
    uint64_t w[ORC_NSYM];
    char rcode[ORC_NSYM][ORC_MAX_CODE];
    int rlen[ORC_NSYM];

    /* stable insertion by (count desc); equal counts keep symbol order because
     * an element only passes strictly smaller ones (t.c:87) */
    for (int s = 0; s < ORC_NSYM; s++) {
        int j = s;
        while (j > 0 && hist[order[j - 1]] < hist[s]) {
            order[j] = order[j - 1];
            j--;
        }
        order[j] = s;
    }
    for (int r = 0; r < ORC_NSYM; r++) {
        w[r] = hist[order[r]];
        rcode[r][0] = '\0';
        rlen[r] = 0;
    }
    memset(codes, 0, (size_t)ORC_NSYM * ORC_MAX_CODE);

    int nn = ORC_NSYM - 1;
    while (nn > 0 && w[nn] == 0) nn--;
    /* (an all-zero histogram never reaches T: blocks are non-empty) */

    /* explicit stack instead of recursion; order of expansion is irrelevant
     * because ranges are disjoint */
    int stack[2 * ORC_NSYM][2], sp = 0;
    stack[sp][0] = 0;
    stack[sp][1] = nn;
    sp++;
    while (sp) {
        sp--;
        const int lo = stack[sp][0], hi = stack[sp][1];
        if (lo == hi) continue; /* t.c:185 */
        const int d = sf_split(w, lo, hi);
        for (int r = lo; r <= hi; r++) {
            rcode[r][rlen[r]++] = (r <= d) ? '0' : '1';
            rcode[r][rlen[r]] = '\0';
        }
        stack[sp][0] = lo, stack[sp][1] = d, sp++;
        stack[sp][0] = d + 1, stack[sp][1] = hi, sp++;
    }
    /* t.c:355: symbol s prints codes[positions[s]] */
    for (int r = 0; r < ORC_NSYM; r++) memcpy(codes[order[r]], rcode[r], (size_t)rlen[r] + 1);
}

int orc_sf_codes_packed(const uint64_t hist[ORC_NSYM], uint8_t len[ORC_NSYM], uint64_t code[ORC_NSYM])
{
    static _Thread_local char codes[ORC_NSYM][ORC_MAX_CODE];
    int maxlen = 0, too_long = 0;
    orc_sf_codes(hist, codes);
    for (int s = 0; s < ORC_NSYM; s++) {
        size_t l = strlen(codes[s]);
        uint64_t v = 0;
        if (l > 64) too_long = 1;
        for (size_t i = 0; i < l && i < 64; i++) v = (v << 1) | (uint64_t)(codes[s][i] == '1');
        len[s] = (uint8_t)l;
        code[s] = v;
        if ((int)l > maxlen) maxlen = (int)l;
    }
    return too_long ? -1 : maxlen;
}

/* t.c:353-361 */
size_t orc_cod_fields(const char codes[ORC_NSYM][ORC_MAX_CODE], char *out)
{
    char *p = out;
    for (int s = 0; s < ORC_NSYM; s++) {
        size_t l = strlen(codes[s]);
        memcpy(p, codes[s], l);
        p += l;
        if (s != ORC_NSYM - 1) *p++ = ';';
    }
    *p = '\0';
    return (size_t)(p - out);
}

/* c.c:115-177 / d.c:479-495: fields are '0'/'1' runs split by ';'. */
int orc_cod_parse(const char *text, char codes[ORC_NSYM][ORC_MAX_CODE])
{
    const char *p = text;
    memset(codes, 0, (size_t)ORC_NSYM * ORC_MAX_CODE);
    for (int s = 0; s < ORC_NSYM; s++) {
        int l = 0;
        while (*p == '0' || *p == '1') {
            if (l >= ORC_MAX_CODE - 1) return ORC_FILE_UNRECOGNIZABLE;
            codes[s][l++] = *p++;
        }
        if (s != ORC_NSYM - 1) {
            if (*p != ';') return ORC_FILE_UNRECOGNIZABLE; /* c.c:127-131,165-169 */
            p++;
        }
    }
    return (*p == '\0') ? ORC_OK : ORC_FILE_UNRECOGNIZABLE; /* c.c:174-177 */
}

/* ------------------------------------------------------------------ */
/* Module C                                                            */
/* ------------------------------------------------------------------ */

/* c.c:52-83 builds the stream by OR-ing pre-shifted code bytes; the result is
 * simply the codes of the symbols concatenated MSB-first, the last byte zero
 * padded, z = ceil(bits/8) (c.c:80). */
uint64_t orc_sf_encode_block(const uint8_t *in, uint64_t n,
                             const char codes[ORC_NSYM][ORC_MAX_CODE], uint8_t *out)
{
    uint16_t clen[ORC_NSYM];
    for (int s = 0; s < ORC_NSYM; s++) clen[s] = (uint16_t)strlen(codes[s]);
    uint64_t bit = 0;
    uint8_t acc = 0;
    for (uint64_t i = 0; i < n; i++) {
        const char *c = codes[in[i]];
        const unsigned l = clen[in[i]];
        for (unsigned k = 0; k < l; k++) {
            acc = (uint8_t)((acc << 1) | (c[k] == '1'));
            if ((++bit & 7) == 0) {
                out[(bit >> 3) - 1] = acc;
                acc = 0;
            }
        }
    }
    if (bit & 7) {
        out[bit >> 3] = (uint8_t)(acc << (8 - (bit & 7)));
        return (bit >> 3) + 1;
    }
    return bit >> 3;
}

uint64_t orc_sf_encoded_size(const uint64_t hist[ORC_NSYM], const char codes[ORC_NSYM][ORC_MAX_CODE])
{
    uint64_t bits = 0;
    for (int s = 0; s < ORC_NSYM; s++) bits += hist[s] * (uint64_t)strlen(codes[s]);
    return (bits + 7) >> 3;
}

/* ------------------------------------------------------------------ */
/* Module D                                                            */
/* ------------------------------------------------------------------ */

/* d.c:394-504 builds a pointer trie; here it is a flat node array.
 * d.c:530-548: one bit per step MSB-first, emit on reaching a leaf, stop after
 * n symbols.  Stepping to a missing child (the reference dereferences NULL,
 * d.c:533-536) or past the payload is reported as FILE_UNRECOGNIZABLE. */
int orc_sf_decode_block(const uint8_t *payload, uint64_t z, uint64_t n,
                        const char codes[ORC_NSYM][ORC_MAX_CODE], uint8_t *out)
{
    enum { MAXN = 2 * ORC_NSYM * ORC_MAX_CODE };
    int32_t(*child)[2] = (int32_t(*)[2])malloc(sizeof(int32_t[2]) * MAXN);
    int16_t *leaf = (int16_t *)malloc(sizeof(int16_t) * MAXN);
    if (!child || !leaf) {
        free(child);
        free(leaf);
        return ORC_LACK_OF_MEMORY;
    }
    int nodes = 1;
    child[0][0] = child[0][1] = -1;
    leaf[0] = -1;
    for (int s = 0; s < ORC_NSYM; s++) {
        const char *c = codes[s];
        if (!*c) continue;
        int cur = 0;
        for (; *c; c++) {
            int b = (*c == '1');
            if (child[cur][b] < 0) {
                child[nodes][0] = child[nodes][1] = -1;
                leaf[nodes] = -1;
                child[cur][b] = nodes++;
            }
            cur = child[cur][b];
        }
        leaf[cur] = (int16_t)s;
    }
    int rc = ORC_OK;
    uint64_t produced = 0, bit = 0;
    int cur = 0;
    while (produced < n) {
        if ((bit >> 3) >= z) {
            rc = ORC_FILE_UNRECOGNIZABLE;
            break;
        }
        int b = (payload[bit >> 3] >> (7 - (bit & 7))) & 1;
        bit++;
        cur = child[cur][b];
        if (cur < 0) {
            rc = ORC_FILE_UNRECOGNIZABLE;
            break;
        }
        if (child[cur][0] < 0 && child[cur][1] < 0) { /* d.c:536 */
            out[produced++] = (uint8_t)leaf[cur];
            cur = 0;
        }
    }
    free(child);
    free(leaf);
    return rc;
}

/* d.c:144-186.  0x00 opens {0, sym, count}; count 0 yields ONE copy of sym
 * (d.c:179-184); anything else is a literal.  The reference grows its buffer
 * stepwise up to 64 MiB + 1 KiB and fails beyond (d.c:154-168). */
int orc_rle_decode_block(const uint8_t *in, uint64_t n, uint8_t *out, uint64_t cap, uint64_t *final)
{
    const uint64_t hard_cap = 67108864ull + 1024ull;
    uint64_t o = 0;
    int rc = ORC_OK;
    for (uint64_t i = 0; i < n; i++) {
        uint8_t sym = in[i];
        unsigned reps = 0;
        if (sym == 0) {
            if (i + 2 >= n) { /* reference reads past the block here */
                rc = ORC_FILE_UNRECOGNIZABLE;
                break;
            }
            sym = in[++i];
            reps = in[++i];
        }
        const uint64_t emit = reps ? reps : 1;
        if (o + reps > hard_cap) {
            rc = ORC_FILE_UNRECOGNIZABLE;
            break;
        }
        if (out && o + emit <= cap) memset(out + o, sym, emit);
        o += emit;
    }
    *final = o;
    if (rc == ORC_OK && out && o > cap) rc = ORC_LACK_OF_MEMORY;
    return rc;
}

/* ------------------------------------------------------------------ */
/* whole-file drivers                                                  */
/* ------------------------------------------------------------------ */

void orc_f_result_free(orc_f_result *res)
{
    orc_buf_free(&res->rle);
    orc_buf_free(&res->rle_freq);
    orc_buf_free(&res->freq);
    free(res->rle_sizes);
    res->rle_sizes = NULL;
}

static int put_freq_block(orc_buf *dst, uint64_t n, const uint64_t hist[ORC_NSYM], int last)
{
    char fields[ORC_NSYM * 22];
    size_t k = orc_freq_fields(hist, fields);
    if (!buf_printf_u64(dst, "@", n, "@")) return 0; /* f.c:312,327 */
    if (!buf_put(dst, fields, k)) return 0;
    if (last && !buf_put(dst, "@0", 2)) return 0; /* f.c:112-116 */
    return 1;
}

/* f.c:180-412 */
int orc_module_f(const uint8_t *in, uint64_t n, uint64_t block_size, int force_rle,
                 int force_freq, orc_f_result *res)
{
    memset(res, 0, sizeof *res);
    if (block_size < 512) block_size = 512; /* utils/file.c:61-65 */
    if (block_size > 67108864ull) return ORC_FILE_STREAM_FAILED;
    if (n < 1024) return ORC_FILE_TOO_SMALL; /* f.c:220,366 */
    const uint64_t B = (n + block_size - 1) / block_size; /* utils/file.c:80-85 */
    res->n_blocks = B;
    res->rle_sizes = (uint64_t *)calloc(B, sizeof(uint64_t));
    uint8_t *scratch = (uint8_t *)malloc(2 * block_size + 3); /* f.c:244 */
    if (!res->rle_sizes || !scratch) {
        free(scratch);
        return ORC_LACK_OF_MEMORY;
    }
    int use_rle = 1, ok = 1;
    uint64_t hist[ORC_NSYM];
    for (uint64_t b = 0; b < B && ok; b++) {
        const uint8_t *blk = in + b * block_size;
        const uint64_t len = (b == B - 1) ? n - b * block_size : block_size;
        uint64_t rsz = 0;
        if (use_rle) {
            rsz = orc_rle_encode_block(blk, len, scratch);
            if (b == 0) use_rle = orc_rle_gate(len, rsz, force_rle); /* f.c:250-258 */
        }
        if (b == 0) {
            if (use_rle) ok &= buf_printf_u64(&res->rle_freq, "@R@", B, ""); /* f.c:289 */
            if (!use_rle || force_freq) ok &= buf_printf_u64(&res->freq, "@N@", B, ""); /* f.c:294 */
        }
        if (use_rle) {
            res->rle_sizes[b] = rsz;
            ok &= buf_put(&res->rle, scratch, rsz); /* f.c:307 */
            orc_histogram(scratch, rsz, hist);      /* f.c:310 */
            ok &= put_freq_block(&res->rle_freq, rsz, hist, b == B - 1);
        }
        if (!use_rle || force_freq) {
            orc_histogram(blk, len, hist); /* f.c:325 */
            ok &= put_freq_block(&res->freq, len, hist, b == B - 1);
        }
    }
    free(scratch);
    res->rle_mode = use_rle;
    if (!ok) {
        orc_f_result_free(res);
        return ORC_LACK_OF_MEMORY;
    }
    return ORC_OK;
}

/* tiny cursor over a text container */
typedef struct {
    const uint8_t *p, *end;
} cursor;

static int cur_lit(cursor *c, char ch)
{
    if (c->p < c->end && *c->p == (uint8_t)ch) {
        c->p++;
        return 1;
    }
    return 0;
}

static int cur_u64(cursor *c, uint64_t *v)
{
    if (c->p >= c->end || *c->p < '0' || *c->p > '9') return 0;
    uint64_t x = 0;
    while (c->p < c->end && *c->p >= '0' && *c->p <= '9') x = x * 10 + (uint64_t)(*c->p++ - '0');
    *v = x;
    return 1;
}

/* copy chars up to the next '@' (or end) into dst (NUL terminated) */
static int cur_field(cursor *c, char *dst, size_t cap)
{
    size_t k = 0;
    while (c->p < c->end && *c->p != '@') {
        if (k + 1 >= cap) return 0;
        dst[k++] = (char)*c->p++;
    }
    dst[k] = '\0';
    return k > 0;
}

/* t.c:246-445 */
int orc_module_t(const uint8_t *freq_text, uint64_t freq_len, orc_buf *cod)
{
    cursor c = {freq_text, freq_text + freq_len};
    uint64_t B = 0;
    memset(cod, 0, sizeof *cod);
    if (!cur_lit(&c, '@') || c.p >= c.end) return ORC_FILE_UNRECOGNIZABLE;
    const char mode = (char)*c.p++;
    if (!cur_lit(&c, '@') || !cur_u64(&c, &B)) return ORC_FILE_UNRECOGNIZABLE; /* t.c:278 */
    if (mode != 'R' && mode != 'N') return ORC_FILE_UNRECOGNIZABLE;            /* t.c:281 */
    char hdr[48];
    int k = snprintf(hdr, sizeof hdr, "@%c@%llu", mode, (unsigned long long)B); /* t.c:302 */
    if (!buf_put(cod, hdr, (uint64_t)k)) return ORC_LACK_OF_MEMORY;

    char *fields = (char *)malloc(2560 + 8);
    char(*codes)[ORC_MAX_CODE] = (char(*)[ORC_MAX_CODE])malloc((size_t)ORC_NSYM * ORC_MAX_CODE);
    char *text = (char *)malloc((size_t)ORC_NSYM * ORC_MAX_CODE);
    int rc = ORC_OK;
    if (!fields || !codes || !text) rc = ORC_LACK_OF_MEMORY;
    for (uint64_t b = 0; b < B && rc == ORC_OK; b++) {
        uint64_t n, hist[ORC_NSYM];
        if (!cur_lit(&c, '@') || !cur_u64(&c, &n) || !cur_lit(&c, '@') ||
            !cur_field(&c, fields, 2560)) { /* t.c:320,332 */
            rc = ORC_FILE_STREAM_FAILED;
            break;
        }
        rc = orc_freq_parse(fields, hist);
        if (rc) break;
        orc_sf_codes(hist, codes);
        size_t tl = orc_cod_fields(codes, text);
        if (!buf_printf_u64(cod, "@", n, "@") || !buf_put(cod, text, tl)) rc = ORC_LACK_OF_MEMORY;
    }
    if (rc == ORC_OK && !buf_put(cod, "@0", 2)) rc = ORC_LACK_OF_MEMORY; /* t.c:395-396 */
    free(fields);
    free(codes);
    free(text);
    if (rc) orc_buf_free(cod);
    return rc;
}

#define COD_FIELD_CAP 33152 /* c.c:362, d.c:712 */

/* c.c:306-472 */
int orc_module_c(const uint8_t *sym, uint64_t n, const uint8_t *cod_text, uint64_t cod_len, orc_buf *shaf)
{
    cursor c = {cod_text, cod_text + cod_len};
    uint64_t B = 0, pos = 0;
    memset(shaf, 0, sizeof *shaf);
    if (!cur_lit(&c, '@') || c.p >= c.end) return ORC_FILE_UNRECOGNIZABLE;
    c.p++; /* mode char is skipped, c.c:333 */
    if (!cur_lit(&c, '@') || !cur_u64(&c, &B)) return ORC_FILE_UNRECOGNIZABLE;
    if (!buf_printf_u64(shaf, "@", B, "")) return ORC_LACK_OF_MEMORY; /* c.c:351 */

    char *text = (char *)malloc(COD_FIELD_CAP);
    char(*codes)[ORC_MAX_CODE] = (char(*)[ORC_MAX_CODE])malloc((size_t)ORC_NSYM * ORC_MAX_CODE);
    int rc = (!text || !codes) ? ORC_LACK_OF_MEMORY : ORC_OK;
    for (uint64_t b = 0; b < B && rc == ORC_OK; b++) {
        uint64_t bn;
        if (!cur_lit(&c, '@') || !cur_u64(&c, &bn) || !cur_lit(&c, '@') ||
            !cur_field(&c, text, COD_FIELD_CAP)) { /* c.c:369 */
            rc = ORC_FILE_STREAM_FAILED;
            break;
        }
        if (pos + bn > n) { /* c.c:392 short read */
            rc = ORC_FILE_STREAM_FAILED;
            break;
        }
        rc = orc_cod_parse(text, codes);
        if (rc) break;
        uint64_t hist[ORC_NSYM];
        orc_histogram(sym + pos, bn, hist);
        const uint64_t z = orc_sf_encoded_size(hist, codes);
        if (!buf_printf_u64(shaf, "@", z, "@") || !buf_reserve(shaf, z + 1)) { /* c.c:256 */
            rc = ORC_LACK_OF_MEMORY;
            break;
        }
        const uint64_t z2 = orc_sf_encode_block(sym + pos, bn, codes, shaf->data + shaf->size);
        if (z2 != z) {
            rc = ORC_FILE_UNRECOGNIZABLE;
            break;
        }
        shaf->size += z;
        pos += bn;
    }
    free(text);
    free(codes);
    if (rc) orc_buf_free(shaf);
    return rc;
}

/* d.c:628-834 */
int orc_module_d_shaf(const uint8_t *shaf, uint64_t shaf_len, const uint8_t *cod_text,
                      uint64_t cod_len, int rle, orc_buf *out)
{
    cursor s = {shaf, shaf + shaf_len}, c = {cod_text, cod_text + cod_len};
    uint64_t B = 0, Bc = 0;
    memset(out, 0, sizeof *out);
    if (!cur_lit(&s, '@') || !cur_u64(&s, &B)) return ORC_FILE_STREAM_FAILED; /* d.c:673 */
    if (!cur_lit(&c, '@') || c.p >= c.end) return ORC_FILE_STREAM_FAILED;
    const char mode = (char)*c.p++;
    if (!cur_lit(&c, '@') || !cur_u64(&c, &Bc)) return ORC_FILE_STREAM_FAILED; /* d.c:676 */
    if (!((mode == 'N' && !rle) || mode == 'R')) return ORC_FILE_UNRECOGNIZABLE; /* d.c:678 */
    B = Bc; /* the reference overwrites length with the .cod value */

    char *text = (char *)malloc(COD_FIELD_CAP);
    char(*codes)[ORC_MAX_CODE] = (char(*)[ORC_MAX_CODE])malloc((size_t)ORC_NSYM * ORC_MAX_CODE);
    int rc = (!text || !codes) ? ORC_LACK_OF_MEMORY : ORC_OK;
    uint8_t *symbuf = NULL;
    uint64_t symcap = 0;
    for (uint64_t b = 0; b < B && rc == ORC_OK; b++) {
        uint64_t z, n;
        if (!cur_lit(&s, '@') || !cur_u64(&s, &z) || !cur_lit(&s, '@') ||
            (uint64_t)(s.end - s.p) < z) { /* d.c:697,706 */
            rc = ORC_FILE_STREAM_FAILED;
            break;
        }
        const uint8_t *payload = s.p;
        s.p += z;
        if (!cur_lit(&c, '@') || !cur_u64(&c, &n) || !cur_lit(&c, '@') ||
            !cur_field(&c, text, COD_FIELD_CAP)) { /* d.c:709,716 */
            rc = ORC_FILE_STREAM_FAILED;
            break;
        }
        rc = orc_cod_parse(text, codes);
        if (rc) break;
        if (n > symcap) {
            free(symbuf);
            symbuf = (uint8_t *)malloc(n + 1);
            symcap = n;
            if (!symbuf) {
                rc = ORC_LACK_OF_MEMORY;
                break;
            }
        }
        rc = orc_sf_decode_block(payload, z, n, codes, symbuf);
        if (rc) break;
        if (!rle) {
            if (!buf_put(out, symbuf, n)) rc = ORC_LACK_OF_MEMORY;
        } else {
            uint64_t fin = 0;
            rc = orc_rle_decode_block(symbuf, n, NULL, 0, &fin);
            if (rc) break;
            if (!buf_reserve(out, fin + 1)) {
                rc = ORC_LACK_OF_MEMORY;
                break;
            }
            rc = orc_rle_decode_block(symbuf, n, out->data + out->size, fin, &fin);
            out->size += fin;
        }
    }
    free(text);
    free(codes);
    free(symbuf);
    if (rc) orc_buf_free(out);
    return rc;
}

/* d.c:232-388 */
int orc_module_d_rle(const uint8_t *rle, uint64_t rle_len, const uint8_t *freq_text,
                     uint64_t freq_len, orc_buf *out)
{
    cursor c = {freq_text, freq_text + freq_len};
    uint64_t B = 0, pos = 0;
    memset(out, 0, sizeof *out);
    if (!cur_lit(&c, '@') || c.p >= c.end) return ORC_FILE_STREAM_FAILED;
    const char mode = (char)*c.p++;
    if (!cur_lit(&c, '@') || !cur_u64(&c, &B)) return ORC_FILE_STREAM_FAILED; /* d.c:268 */
    if (mode != 'R') return ORC_FILE_UNRECOGNIZABLE;                          /* d.c:270 */
    int rc = ORC_OK;
    for (uint64_t b = 0; b < B && rc == ORC_OK; b++) {
        uint64_t n, fin = 0;
        if (!cur_lit(&c, '@') || !cur_u64(&c, &n) || !cur_lit(&c, '@')) { /* d.c:278 */
            rc = ORC_FILE_STREAM_FAILED;
            break;
        }
        while (c.p < c.end && *c.p != '@') c.p++;
        if (pos + n > rle_len) { /* d.c:84 short read */
            rc = ORC_FILE_STREAM_FAILED;
            break;
        }
        rc = orc_rle_decode_block(rle + pos, n, NULL, 0, &fin);
        if (rc) break;
        if (!buf_reserve(out, fin + 1)) {
            rc = ORC_LACK_OF_MEMORY;
            break;
        }
        rc = orc_rle_decode_block(rle + pos, n, out->data + out->size, fin, &fin);
        out->size += fin;
        pos += n;
    }
    if (rc) orc_buf_free(out);
    return rc;
}
