/*
 * mod_t.c -- module T: Shannon-Fano codes from a .freq file.
 * Interface and files: reference t.c:246-445 (get_shafa_codes).  Parsing of the
 * run-elided fields (t.c:27-65) and the .cod text (t.c:302,350-361,395-396) stay
 * here; sorting, split search and code assignment (t.c:74-210) are one CUDA
 * launch over all blocks (shafa_cuda_sf_build_codes).
 */
#include <stdlib.h>
#include <string.h>

#include "util.h"

/* t.c:27-65: 256 ';'-separated fields, an empty field repeats the previous value */
static _modules_error parse_fields(const uint8_t *p, size_t len, uint64_t *h)
{
    const uint8_t *end = p + len;
    for (int s = 0; s < 256; s++) {
        if (p < end && *p >= '0' && *p <= '9') {
            uint64_t v = 0;
            while (p < end && *p >= '0' && *p <= '9') v = v * 10 + (uint64_t)(*p++ - '0');
            h[s] = v;
        } else {
            if (s == 0) return _FILE_UNRECOGNIZABLE; /* t.c:31 */
            h[s] = h[s - 1];                           /* t.c:50-53 */
        }
        if (s != 255) {
            if (p >= end || *p != ';') return _FILE_UNRECOGNIZABLE; /* t.c:43,56 */
            p++;
        }
    }
    return p == end ? _SUCCESS : _FILE_UNRECOGNIZABLE; /* t.c:61 */
}

void t_summary(uint64_t nb, const uint64_t *sizes, double ms, const char *p_cod)
{
    printf("Module:T (Calculation of symbol codes) [libshafa_cuda]\nNumber of blocks: %llu\nSize of blocks analyzed in the symbol file: ",
           (unsigned long long)nb);
    for (uint64_t b = 0; b < nb; b++) printf("%llu%s", (unsigned long long)sizes[b], b + 1 == nb ? " bytes\n" : "/");
    printf("Module runtime (milliseconds): %f\nGenerated file %s\n", ms, p_cod);
}

_modules_error get_shafa_codes(const char *path)
{
    const double t0 = now_ms();
    char *p_freq = add_ext(path, FREQ_EXT), *p_cod = add_ext(path, CODES_EXT);
    if (!p_freq || !p_cod) {
        free(p_freq);
        free(p_cod);
        return _LACK_OF_MEMORY;
    }
    blob in;
    _modules_error err = blob_read(p_freq, &in);
    free(p_freq);
    if (err) {
        free(p_cod);
        return err;
    }
    cursor c = {in.data, in.data + in.size};
    uint64_t nb = 0, *sizes = NULL, *hist = NULL, *code = NULL;
    uint8_t *len = NULL;
    void *d_hist = NULL, *d_code = NULL, *d_len = NULL;
    shafa_cuda_ctx *g = gpu();
    char mode = 0;
    if (!cur_lit(&c, '@') || c.p >= c.end) {
        err = _FILE_UNRECOGNIZABLE;
        goto out;
    }
    mode = (char)*c.p++;
    if (!cur_lit(&c, '@') || !cur_u64(&c, &nb)) { /* t.c:278 */
        err = _FILE_UNRECOGNIZABLE;
        goto out;
    }
    if (mode != 'R' && mode != 'N') { /* t.c:281 */
        err = _FILE_UNRECOGNIZABLE;
        goto out;
    }
    if (!block_count_plausible(nb, in.size)) {
        err = _FILE_UNRECOGNIZABLE;
        goto out;
    }
    if (!g) {
        err = _THREAD_CREATION_FAILED;
        goto out;
    }
    sizes = malloc((nb + 1) * sizeof *sizes);
    hist = malloc((nb * 256 + 1) * sizeof *hist);
    code = malloc((nb * 256 + 1) * sizeof *code);
    len = malloc(nb * 256 + 1);
    if (!sizes || !hist || !code || !len) {
        err = _LACK_OF_MEMORY;
        goto out;
    }
    for (uint64_t b = 0; b < nb; b++) {
        const uint8_t *f;
        size_t fl;
        if (!cur_lit(&c, '@') || !cur_u64(&c, &sizes[b]) || !cur_lit(&c, '@') || !cur_field(&c, &f, &fl)) { /* t.c:320,332 */
            err = _FILE_STREAM_FAILED;
            goto out;
        }
        if (fl > 2559) { /* t.c:326,332 field buffer */
            err = _FILE_UNRECOGNIZABLE;
            goto out;
        }
        if ((err = parse_fields(f, fl, hist + b * 256))) goto out;
    }
    if (nb) {
        d_hist = shafa_cuda_dev_alloc(g, nb * 256 * 8);
        d_code = shafa_cuda_dev_alloc(g, nb * 256 * 8);
        d_len = shafa_cuda_dev_alloc(g, nb * 256);
        if (!d_hist || !d_code || !d_len) {
            err = _LACK_OF_MEMORY;
            goto out;
        }
        if ((err = gpu_status(shafa_cuda_memcpy_h2d(g, d_hist, hist, nb * 256 * 8)))) goto out;
        if ((err = gpu_status(shafa_cuda_sf_build_codes(g, d_hist, nb, d_len, d_code)))) goto out;
        if ((err = gpu_status(shafa_cuda_memcpy_d2h(g, code, d_code, nb * 256 * 8)))) goto out;
        if ((err = gpu_status(shafa_cuda_memcpy_d2h(g, len, d_len, nb * 256)))) goto out;
        if ((err = gpu_status(shafa_cuda_synchronize(g)))) goto out;
    }
    err = write_cod_file(p_cod, mode, nb, sizes, len, code);
out:
    if (d_hist) shafa_cuda_dev_free(g, d_hist);
    if (d_code) shafa_cuda_dev_free(g, d_code);
    if (d_len) shafa_cuda_dev_free(g, d_len);
    if (!err) t_summary(nb, sizes, now_ms() - t0, p_cod);
    blob_free(&in);
    free(sizes);
    free(hist);
    free(code);
    free(len);
    free(p_cod);
    return err;
}
