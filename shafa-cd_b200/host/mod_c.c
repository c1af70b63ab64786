/*
 * mod_c.c -- module C: Shannon-Fano bit packing.
 * Interface and files: reference c.c:306-472 (shafa_compress).  The reference
 * spawns one pthread per block (compress_to_buffer / write_shafa through
 * multithread_create, c.c:411) and serialises the writes with a join chain; here
 * all blocks are encoded by one batched CUDA call and the ordered write is a
 * plain loop.
 */
#include <stdlib.h>
#include <string.h>

#include "util.h"

void c_summary(uint64_t nb, const uint64_t *sizes, const uint64_t *z, double ms, const char *p_shaf)
{
    printf("Module: C (Symbol codes' codification) [libshafa_cuda]\nNumber of blocks: %llu\n", (unsigned long long)nb);
    for (uint64_t b = 0; b < nb; b++)
        printf("Size before/after & compression rate (Block %llu): %llu/%llu -> %d%%\n", (unsigned long long)b,
               (unsigned long long)sizes[b], (unsigned long long)z[b], (int)(((float)z[b] / sizes[b]) * 100));
    printf("Module runtime (milliseconds): %f\nGenerated file %s\n", ms, p_shaf);
}

_modules_error shafa_compress(char **const path)
{
    const double t0 = now_ms();
    char *p_cod = add_ext(*path, CODES_EXT), *p_shaf = add_ext(*path, SHAFA_EXT);
    if (!p_cod || !p_shaf) {
        free(p_cod);
        free(p_shaf);
        return _LACK_OF_MEMORY;
    }
    gpu_prefetch(); /* CUDA start-up overlaps the reads */
    blob cod = {0}, sym = {0};
    _modules_error err = blob_read(p_cod, &cod);
    free(p_cod);
    uint64_t nb = 0, *sizes = NULL, *offsets = NULL, *code = NULL, *z = NULL, *poff = NULL, *code4 = NULL;
    uint8_t *len = NULL, *h_payload = NULL;
    const uint8_t **fld = NULL; /* the blocks' code fields, kept for the wide parser */
    size_t *fll = NULL;
    bool wide = false;          /* some code has more than 64 bits (c.c:18 allows 32 bytes): the plain kernels of sf_wide.cu */
    void *d_sym = NULL, *d_len = NULL, *d_code = NULL, *d_payload = NULL;
    shafa_cuda_ctx *g = NULL;
    if (err) goto out;
    cursor c = {cod.data, cod.data + cod.size};
    if (!cur_lit(&c, '@') || c.p >= c.end) {
        err = _FILE_UNRECOGNIZABLE;
        goto out;
    }
    c.p++; /* mode char not used here (c.c:333) */
    if (!cur_lit(&c, '@') || !cur_u64(&c, &nb)) {
        err = _FILE_UNRECOGNIZABLE;
        goto out;
    }
    if (!block_count_plausible(nb, cod.size)) {
        err = _FILE_UNRECOGNIZABLE;
        goto out;
    }
    if ((err = blob_read(*path, &sym))) goto out;
    g = gpu();
    if (!g) {
        err = _THREAD_CREATION_FAILED;
        goto out;
    }
    sizes = malloc((nb + 1) * sizeof *sizes);
    offsets = malloc((nb + 1) * sizeof *offsets);
    code = malloc((nb * 256 + 1) * sizeof *code);
    len = malloc(nb * 256 + 1);
    z = malloc((nb + 1) * sizeof *z);
    poff = malloc((nb + 1) * sizeof *poff);
    fld = malloc((nb + 1) * sizeof *fld);
    fll = malloc((nb + 1) * sizeof *fll);
    if (!sizes || !offsets || !code || !len || !z || !poff || !fld || !fll) {
        err = _LACK_OF_MEMORY;
        goto out;
    }
    offsets[0] = 0;
    for (uint64_t b = 0; b < nb; b++) {
        const uint8_t *f;
        size_t fl;
        if (!cur_lit(&c, '@') || !cur_u64(&c, &sizes[b]) || !cur_lit(&c, '@') || !cur_field(&c, &f, &fl)) { /* c.c:369 */
            err = _FILE_STREAM_FAILED;
            goto out;
        }
        fld[b] = f;
        fll[b] = fl;
        if (cod_max_code_bits(f, fl) > 64) wide = true;
        if (!wide && (err = cod_parse_fields(f, fl, len + b * 256, code + b * 256))) goto out;
        offsets[b + 1] = offsets[b] + sizes[b];
        if (sizes[b] > sym.size || offsets[b + 1] > sym.size) { /* c.c:392 short read (and no wrap-around) */
            err = _FILE_STREAM_FAILED;
            goto out;
        }
    }
    if (nb && wide) {
        /* a table with codes of 65..255 bits: every block through shafa_cuda_sf_encode_wide_host */
        code4 = malloc(nb * 256 * 4 * sizeof *code4);
        if (!code4) {
            err = _LACK_OF_MEMORY;
            goto out;
        }
        for (uint64_t b = 0; b < nb; b++)
            if ((err = cod_parse_fields_wide(fld[b], fll[b], len + b * 256, code4 + b * 1024))) goto out;
        uint64_t cap = offsets[nb] + offsets[nb] / 8 + 32 * nb + 4096, ptotal = 0;
        for (int attempt = 0; attempt < 2; attempt++) {
            h_payload = cli_host_alloc(cap + 16);
            if (!h_payload) {
                err = _LACK_OF_MEMORY;
                goto out;
            }
            err = gpu_status(shafa_cuda_sf_encode_wide_host(g, sym.data, offsets, nb, len, code4, h_payload, cap, poff, z, &ptotal));
            if (err != _LACK_OF_MEMORY || attempt || ptotal <= cap) break;
            cli_host_free(h_payload); /* the exact total is known now */
            h_payload = NULL;
            cap = ptotal;
        }
        if (err) goto out;
    } else if (nb) {
        const uint64_t total = offsets[nb];
        /* exact sizes come back from the library; capacity: <= 64 bits per symbol is the hard bound, 9/8 the practical one */
        uint64_t cap = total + total / 8 + 32 * nb + 4096;
        d_sym = shafa_cuda_dev_alloc(g, total + 64);
        d_len = shafa_cuda_dev_alloc(g, nb * 256);
        d_code = shafa_cuda_dev_alloc(g, nb * 256 * 8);
        if (!d_sym || !d_len || !d_code) {
            err = _LACK_OF_MEMORY;
            goto out;
        }
        if ((err = gpu_status(shafa_cuda_memcpy_h2d(g, d_sym, sym.data, total)))) goto out;
        if ((err = gpu_status(shafa_cuda_memcpy_h2d(g, d_len, len, nb * 256)))) goto out;
        if ((err = gpu_status(shafa_cuda_memcpy_h2d(g, d_code, code, nb * 256 * 8)))) goto out;
        for (int attempt = 0; attempt < 2; attempt++) {
            d_payload = shafa_cuda_dev_alloc(g, cap);
            if (!d_payload) {
                err = _LACK_OF_MEMORY;
                goto out;
            }
            err = gpu_status(shafa_cuda_sf_encode_blocks(g, d_sym, offsets, nb, d_len, d_code, NULL, d_payload, cap, poff, z));
            if (err != _LACK_OF_MEMORY || attempt) break;
            /* z[] is exact now: retry with the real total (codes that expand the data) */
            shafa_cuda_dev_free(g, d_payload);
            d_payload = NULL;
            cap = 4096;
            for (uint64_t b = 0; b < nb; b++) cap += (z[b] + 15) & ~15ull;
        }
        if (err) goto out;
        const uint64_t ptotal = poff[nb - 1] + ((z[nb - 1] + 15) & ~15ull);
        h_payload = cli_host_alloc(ptotal + 16);
        if (!h_payload) {
            err = _LACK_OF_MEMORY;
            goto out;
        }
        if ((err = gpu_status(shafa_cuda_memcpy_d2h(g, h_payload, d_payload, ptotal)))) goto out;
        if ((err = gpu_status(shafa_cuda_synchronize(g)))) goto out;
    }
    err = write_shaf_file(p_shaf, nb, z, h_payload, poff);
out:
    if (g) {
        if (d_sym) shafa_cuda_dev_free(g, d_sym);
        if (d_len) shafa_cuda_dev_free(g, d_len);
        if (d_code) shafa_cuda_dev_free(g, d_code);
        if (d_payload) shafa_cuda_dev_free(g, d_payload);
    }
    if (h_payload) cli_host_free(h_payload);
    if (!err) {
        c_summary(nb, sizes, z, now_ms() - t0, p_shaf);
        free(*path); /* c.c:460-461 */
        *path = p_shaf;
        p_shaf = NULL;
    }
    blob_free(&cod);
    blob_free(&sym);
    free(p_shaf);
    free(sizes);
    free(offsets);
    free(code);
    free(code4);
    free(fld);
    free(fll);
    free(len);
    free(z);
    free(poff);
    return err;
}
