/*
 * containers.c -- writers of the text / framed containers, shared by the per-module entry points (mod_*.c) and
 * the fused F -> T -> C path (fused.c).  Formats are the reference's, byte for byte:
 *   .freq  f.c:89-119,289-294,312-314     .cod  t.c:302,350-361,395-396     .shaf  c.c:351,256-258
 */
#include <stdlib.h>
#include <string.h>

#include "util.h"

/* f.c:89-119 body: a count is printed when it differs from the previous symbol's */
static size_t freq_fields(const uint64_t *h, char *out)
{
    char *p = out;
    for (int s = 0; s < 256; s++) {
        if (s == 0 || h[s] != h[s - 1]) p += sprintf(p, "%llu", (unsigned long long)h[s]);
        if (s != 255) *p++ = ';';
    }
    return (size_t)(p - out);
}

_modules_error write_freq_file(const char *path, char mode, uint64_t nb, const uint64_t *sizes, const uint64_t *hist)
{
    FILE *f = fopen(path, "wb");
    if (!f) return _FILE_INACCESSIBLE;
    char fields[256 * 22];
    bool ok = fprintf(f, "@%c@%llu", mode, (unsigned long long)nb) >= 4;
    for (uint64_t b = 0; b < nb && ok; b++) {
        const size_t k = freq_fields(hist + b * 256, fields);
        ok = fprintf(f, "@%llu@", (unsigned long long)sizes[b]) >= 2 && fwrite(fields, 1, k, f) == k;
    }
    ok = ok && fputs("@0", f) >= 0;
    if (fclose(f) != 0) ok = false;
    return ok ? _SUCCESS : _FILE_STREAM_FAILED;
}

_modules_error write_cod_file(const char *path, char mode, uint64_t nb, const uint64_t *sizes, const uint8_t *len, const uint64_t *code)
{
    FILE *f = fopen(path, "wb");
    if (!f) return _FILE_INACCESSIBLE;
    char *text = malloc(256 * 65 + 256);
    bool ok = text && fprintf(f, "@%c@%llu", mode, (unsigned long long)nb) >= 3; /* t.c:302 */
    for (uint64_t b = 0; b < nb && ok; b++) {
        char *p = text;
        for (int s = 0; s < 256; s++) { /* t.c:353-361: codes in symbol order joined by ';' */
            const unsigned l = len[b * 256 + s];
            const uint64_t v = code[b * 256 + s];
            for (unsigned i = 0; i < l; i++) *p++ = ((v >> (l - 1 - i)) & 1) ? '1' : '0';
            if (s != 255) *p++ = ';';
        }
        ok = fprintf(f, "@%llu@", (unsigned long long)sizes[b]) >= 2 && fwrite(text, 1, (size_t)(p - text), f) == (size_t)(p - text);
    }
    ok = ok && fputs("@0", f) >= 0; /* t.c:395-396 */
    free(text);
    if (fclose(f) != 0) ok = false;
    return ok ? _SUCCESS : _FILE_STREAM_FAILED;
}

_modules_error write_shaf_file(const char *path, uint64_t nb, const uint64_t *z, const uint8_t *payload, const uint64_t *poff)
{
    FILE *f = fopen(path, "wb");
    if (!f) return _FILE_INACCESSIBLE;
    setvbuf(f, NULL, _IOFBF, 1 << 20);
    bool ok = fprintf(f, "@%llu", (unsigned long long)nb) >= 2; /* c.c:351 */
    for (uint64_t b = 0; b < nb && ok; b++)                      /* c.c:256-258, in block order */
        ok = fprintf(f, "@%llu@", (unsigned long long)z[b]) >= 2 && fwrite(payload + poff[b], 1, z[b], f) == z[b];
    if (fclose(f) != 0) ok = false;
    return ok ? _SUCCESS : _FILE_STREAM_FAILED;
}

_modules_error write_all(const char *path, const uint8_t *data, uint64_t n)
{
    FILE *f = fopen(path, "wb");
    if (!f) return _FILE_INACCESSIBLE;
    const bool ok = fwrite(data, 1, n, f) == n;
    return (fclose(f) != 0 || !ok) ? _FILE_STREAM_FAILED : _SUCCESS;
}
