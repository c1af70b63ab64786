/*
 * mod_d.c -- module D: Shannon-Fano decoding and/or RLE expansion.
 * Interfaces and files: reference d.c:628-834 (shafa_decompress) and d.c:232-388
 * (rle_decompress).  The reference runs process_shafa_decomp /
 * rle_block_decompressor in one pthread per block and serialises the writes
 * (multithread_create, d.c:342,735); here all blocks go through one batched CUDA
 * call each and the ordered write is a single fwrite of the concatenated output.
 */
#include <stdlib.h>
#include <string.h>

#include "util.h"

/* d.c:129-136: the reference's first guess for an expanded block, by size class */
static uint64_t rle_guess(uint64_t rle_block)
{
    if (rle_block <= KIB_64) return KIB_64 + KIB_1;
    if (rle_block <= KIB_640) return KIB_640 + KIB_1;
    if (rle_block <= MIB_8) return MIB_8 + KIB_1;
    return MIB_64 + KIB_1;
}

static void summary(const char *what, double ms, const uint64_t *before, const uint64_t *after, uint64_t nb, const char *out_path)
{
    printf("Module: D (%s decoding) [libshafa_cuda]\n", what);
    for (uint64_t b = 0; b < nb; b++)
        printf("Size before/after generating file (block %llu): %llu/%llu\n", (unsigned long long)(b + 1), (unsigned long long)before[b],
               (unsigned long long)after[b]);
    printf("Module runtime (in milliseconds): %f\nGenerated file %s\n", ms, out_path);
}

_modules_error rle_decompress(char **const path)
{
    const double t0 = now_ms();
    char *p_freq = add_ext(*path, FREQ_EXT), *p_out = rm_ext(*path);
    if (!p_freq || !p_out) {
        free(p_freq);
        free(p_out);
        return _LACK_OF_MEMORY;
    }
    blob rle = {0}, freq = {0};
    uint64_t nb = 0, *offsets = NULL, *sizes = NULL, *fin = NULL, out_size = 0;
    uint8_t *h_out = NULL;
    shafa_cuda_multi *m = NULL;
    gpu_prefetch(); /* CUDA start-up overlaps the reads */
    _modules_error err = blob_read(*path, &rle);
    if (!err) err = blob_read(p_freq, &freq);
    free(p_freq);
    if (err) goto out;
    cursor c = {freq.data, freq.data + freq.size};
    if (!cur_lit(&c, '@') || c.p >= c.end) {
        err = _FILE_STREAM_FAILED;
        goto out;
    }
    const char mode = (char)*c.p++;
    if (!cur_lit(&c, '@') || !cur_u64(&c, &nb)) { /* d.c:268 */
        err = _FILE_STREAM_FAILED;
        goto out;
    }
    if (mode != 'R' || !block_count_plausible(nb, freq.size)) { /* d.c:270 */
        err = _FILE_UNRECOGNIZABLE;
        goto out;
    }
    offsets = malloc((nb + 1) * sizeof *offsets);
    sizes = malloc((nb + 1) * sizeof *sizes);
    fin = malloc((nb + 1) * sizeof *fin);
    if (!offsets || !sizes || !fin) {
        err = _LACK_OF_MEMORY;
        goto out;
    }
    offsets[0] = 0;
    for (uint64_t b = 0; b < nb; b++) { /* d.c:278: sizes only, the counts are skipped */
        const uint8_t *f;
        size_t fl;
        if (!cur_lit(&c, '@') || !cur_u64(&c, &sizes[b]) || !cur_lit(&c, '@')) {
            err = _FILE_STREAM_FAILED;
            goto out;
        }
        cur_field(&c, &f, &fl);
        offsets[b + 1] = offsets[b] + sizes[b];
        if (sizes[b] > rle.size || offsets[b + 1] > rle.size) { /* d.c:84 short read (and no wrap-around) */
            err = _FILE_STREAM_FAILED;
            goto out;
        }
    }
    m = gpu_multi();
    if (!m) {
        err = _THREAD_CREATION_FAILED;
        goto out;
    }
    if (nb) {
        /* all blocks through the chunked multi-stream (multi-GPU) expansion; the reference's first guess per block
         * (d.c:129-136) sizes the output, the exact total comes back if that was too small (d.c:154-170 grows) */
        uint64_t cap = 0;
        for (uint64_t b = 0; b < nb; b++) cap += rle_guess(sizes[b]);
        for (int attempt = 0; attempt < 2; attempt++) {
            h_out = cli_host_alloc(cap + 64);
            if (!h_out) {
                err = _LACK_OF_MEMORY;
                goto out;
            }
            err = gpu_status(shafa_cuda_rle_decode_host_multi(m, rle.data, offsets, nb, h_out, cap, &out_size, fin));
            if (err != _LACK_OF_MEMORY || attempt || out_size <= cap) break;
            cli_host_free(h_out);
            h_out = NULL;
            cap = out_size;
        }
        if (err) goto out;
    }
    err = write_all(p_out, h_out, out_size); /* d.c:218 */
out:
    if (h_out) cli_host_free(h_out);
    if (!err) {
        summary("RLE", now_ms() - t0, sizes, fin, nb, p_out);
        free(*path); /* d.c:378-379 */
        *path = p_out;
        p_out = NULL;
    }
    blob_free(&rle);
    blob_free(&freq);
    free(p_out);
    free(offsets);
    free(sizes);
    free(fin);
    return err;
}

_modules_error shafa_decompress(char **const path, const bool decompress_rle)
{
    const double t0 = now_ms();
    char *p_base = rm_ext(*path); /* "<x>.shaf" -> "<x>" (d.c:652) */
    char *p_out = NULL, *p_cod = NULL;
    if (p_base) {
        p_out = decompress_rle ? rm_ext(p_base) : add_ext(p_base, ""); /* d.c:655-661 */
        p_cod = add_ext(p_base, CODES_EXT);
    }
    free(p_base);
    if (!p_out || !p_cod) {
        free(p_out);
        free(p_cod);
        return _LACK_OF_MEMORY;
    }
    blob shaf = {0}, cod = {0};
    uint64_t nb = 0, nb_shaf = 0, *z = NULL, *poff = NULL, *n = NULL, *offsets = NULL, *fin = NULL, *code = NULL, out_size = 0;
    uint64_t *code4 = NULL;
    uint8_t *len = NULL, *h_payload = NULL, *h_out = NULL, *h_sym = NULL;
    const uint8_t **fld = NULL; /* the blocks' code fields, kept for the wide parser */
    size_t *fll = NULL;
    bool wide = false;          /* some code has more than 64 bits (d.c:422-504 takes any length): sf_wide.cu */
    shafa_cuda_multi *m = NULL;
    gpu_prefetch(); /* CUDA start-up overlaps the reads */
    _modules_error err = blob_read(*path, &shaf);
    if (!err) err = blob_read(p_cod, &cod);
    free(p_cod);
    if (err) goto out;
    cursor s = {shaf.data, shaf.data + shaf.size}, c = {cod.data, cod.data + cod.size};
    if (!cur_lit(&s, '@') || !cur_u64(&s, &nb_shaf)) { /* d.c:673 */
        err = _FILE_STREAM_FAILED;
        goto out;
    }
    if (!cur_lit(&c, '@') || c.p >= c.end) {
        err = _FILE_STREAM_FAILED;
        goto out;
    }
    const char mode = (char)*c.p++;
    if (!cur_lit(&c, '@') || !cur_u64(&c, &nb)) { /* d.c:676: the .cod count wins */
        err = _FILE_STREAM_FAILED;
        goto out;
    }
    if (!((mode == 'N' && !decompress_rle) || mode == 'R') || !block_count_plausible(nb, cod.size)) { /* d.c:678 */
        err = _FILE_UNRECOGNIZABLE;
        goto out;
    }
    z = malloc((nb + 1) * sizeof *z);
    poff = malloc((nb + 1) * sizeof *poff);
    n = malloc((nb + 1) * sizeof *n);
    offsets = malloc((nb + 1) * sizeof *offsets);
    fin = malloc((nb + 1) * sizeof *fin);
    code = malloc((nb * 256 + 1) * sizeof *code);
    len = malloc(nb * 256 + 1);
    h_payload = cli_host_alloc(shaf.size + 16 * nb + 64);
    fld = malloc((nb + 1) * sizeof *fld);
    fll = malloc((nb + 1) * sizeof *fll);
    if (!z || !poff || !n || !offsets || !fin || !code || !len || !h_payload || !fld || !fll) {
        err = _LACK_OF_MEMORY;
        goto out;
    }
    uint64_t ptotal = 0;
    offsets[0] = 0;
    for (uint64_t b = 0; b < nb; b++) {
        const uint8_t *f;
        size_t fl;
        if (!cur_lit(&s, '@') || !cur_u64(&s, &z[b]) || !cur_lit(&s, '@') || (uint64_t)(s.end - s.p) < z[b]) { /* d.c:697,706 */
            err = _FILE_STREAM_FAILED;
            goto out;
        }
        poff[b] = ptotal; /* payload slots start 16-byte aligned on the device */
        memcpy(h_payload + ptotal, s.p, z[b]);
        s.p += z[b];
        ptotal += (z[b] + 15) & ~15ull;
        if (!cur_lit(&c, '@') || !cur_u64(&c, &n[b]) || !cur_lit(&c, '@') || !cur_field(&c, &f, &fl)) { /* d.c:709,716 */
            err = _FILE_STREAM_FAILED;
            goto out;
        }
        fld[b] = f;
        fll[b] = fl;
        if (cod_max_code_bits(f, fl) > 64) wide = true;
        if (!wide && (err = cod_parse_fields(f, fl, len + b * 256, code + b * 256))) goto out;
        /* a block decodes to at most 64 MiB + 1 KiB of symbols in any file module C can write (2n + 3 of a 64 MiB block
         * is the hard bound); anything larger is a forged header -- and would wrap the sums below */
        if (n[b] > 2ull * MIB_64 + 3) {
            err = _FILE_UNRECOGNIZABLE;
            goto out;
        }
        offsets[b + 1] = offsets[b] + n[b];
    }
    m = gpu_multi();
    if (!m) {
        err = _THREAD_CREATION_FAILED;
        goto out;
    }
    if (nb && wide) {
        /* a table with codes of 65..255 bits: SF decode of every block by shafa_cuda_sf_decode_wide_host, then the
         * ordinary RLE expansion of the symbol stream */
        code4 = malloc(nb * 256 * 4 * sizeof *code4);
        h_sym = cli_host_alloc(offsets[nb] + 64);
        if (!code4 || !h_sym) {
            err = _LACK_OF_MEMORY;
            goto out;
        }
        for (uint64_t b = 0; b < nb; b++)
            if ((err = cod_parse_fields_wide(fld[b], fll[b], len + b * 256, code4 + b * 1024))) goto out;
        if ((err = gpu_status(shafa_cuda_sf_decode_wide_host(shafa_cuda_multi_ctx(m, 0), h_payload, poff, z, n, nb, len, code4, h_sym, offsets[nb]))))
            goto out;
        if (!decompress_rle) {
            h_out = h_sym;
            h_sym = NULL;
            out_size = offsets[nb];
            memcpy(fin, n, nb * sizeof *fin);
        } else {
            uint64_t cap = 0;
            for (uint64_t b = 0; b < nb; b++) cap += rle_guess(n[b]);
            for (int attempt = 0; attempt < 2; attempt++) {
                h_out = cli_host_alloc(cap + 64);
                if (!h_out) {
                    err = _LACK_OF_MEMORY;
                    goto out;
                }
                err = gpu_status(shafa_cuda_rle_decode_host_multi(m, h_sym, offsets, nb, h_out, cap, &out_size, fin));
                if (err != _LACK_OF_MEMORY || attempt || out_size <= cap) break;
                cli_host_free(h_out);
                h_out = NULL;
                cap = out_size;
            }
            if (err) goto out;
        }
    } else if (nb) {
        /* SF decode (+ RLE expansion, d.c:574-586) of all blocks: chunks of blocks on worker streams of every device */
        uint64_t cap = 0;
        for (uint64_t b = 0; b < nb; b++) cap += decompress_rle ? rle_guess(n[b]) : n[b];
        for (int attempt = 0; attempt < 2; attempt++) {
            h_out = cli_host_alloc(cap + 64);
            if (!h_out) {
                err = _LACK_OF_MEMORY;
                goto out;
            }
            err = gpu_status(shafa_cuda_decode_host_multi(m, h_payload, poff, z, n, nb, len, code, decompress_rle, h_out, cap, &out_size, fin));
            if (err != _LACK_OF_MEMORY || attempt || out_size <= cap) break;
            cli_host_free(h_out);
            h_out = NULL;
            cap = out_size;
        }
        if (err) goto out;
    }
    err = write_all(p_out, h_out, out_size); /* d.c:613 */
out:
    if (h_payload) cli_host_free(h_payload);
    if (h_out) cli_host_free(h_out);
    if (h_sym) cli_host_free(h_sym);
    free(code4);
    free(fld);
    free(fll);
    if (!err) {
        summary(decompress_rle ? "SHAFA & RLE" : "SHAFA", now_ms() - t0, z, fin, nb, p_out);
        free(*path); /* d.c:814-815 */
        *path = p_out;
        p_out = NULL;
    }
    blob_free(&shaf);
    blob_free(&cod);
    free(p_out);
    free(z);
    free(poff);
    free(n);
    free(offsets);
    free(fin);
    free(code);
    free(len);
    return err;
}
