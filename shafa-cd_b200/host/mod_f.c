/*
 * mod_f.c -- module F: per-block RLE + byte frequencies.
 * Interface and files: reference f.c:180-412 (freq_rle_compress).  The block
 * loop, the RLE scan (f.c:29-55) and the histograms (f.c:63-79) run as batched
 * CUDA kernels over ALL blocks at once; this file only reads the input, asks
 * block 0 for the R/N decision (f.c:246-258) and writes .rle/.freq text.
 */
#include <stdlib.h>
#include <string.h>

#include "util.h"

void f_summary(uint64_t nb, const uint64_t *sizes, uint64_t n, const uint64_t *rle_sizes, double ms, const char *p_rle,
               const char *p_freq, const char *p_rle_freq)
{
    printf("Module: f (calculation of symbol frequencies) [libshafa_cuda]\nNumber of blocks: %llu\n", (unsigned long long)nb);
    printf("Size of blocks analyzed in the original file: ");
    for (uint64_t b = 0; b < nb; b++) printf("%llu%s", (unsigned long long)sizes[b], b + 1 == nb ? "\n" : "/");
    if (p_rle) {
        uint64_t tot = 0;
        for (uint64_t b = 0; b < nb; b++) tot += rle_sizes[b];
        printf("RLE Compression: %s (%f%% compression)\n", p_rle, 100.0f * ((float)((long)n - (long)tot) / (float)n));
        printf("Size of blocks analyzed in the RLE file: ");
        for (uint64_t b = 0; b < nb; b++) printf("%llu%s", (unsigned long long)rle_sizes[b], b + 1 == nb ? " bytes\n" : "/");
    }
    printf("Module runtime (milliseconds): %f\nGenerated files: ", ms);
    if (p_freq && p_rle_freq)
        printf("%s, %s\n", p_freq, p_rle_freq);
    else
        printf("%s\n", p_freq ? p_freq : p_rle_freq);
}

_modules_error freq_rle_compress(char **const path, const bool force_rle, const bool force_freq, const unsigned long block_size_in)
{
    const double t0 = now_ms();
    uint64_t bs = block_size_in;
    if (bs > MIB_64) return _FILE_STREAM_FAILED; /* utils/file.c:61 */
    if (bs == 0) bs = 524288;                    /* utils/file.c:62 */
    if (bs < 512) bs = 512;                      /* utils/file.c:63 */

    gpu_prefetch(); /* CUDA start-up overlaps the reads */
    blob in;
    _modules_error err = blob_read(*path, &in);
    if (err) return err;
    if (in.size < KIB_1) { /* f.c:220,366 */
        blob_free(&in);
        return _FILE_TOO_SMALL;
    }
    shafa_cuda_ctx *g = gpu();
    if (!g) {
        blob_free(&in);
        return _THREAD_CREATION_FAILED;
    }
    const uint64_t n = in.size, nb = (n + bs - 1) / bs; /* utils/file.c:80-85 */
    char *p_rle = add_ext(*path, RLE_EXT), *p_rle_freq = NULL, *p_freq = add_ext(*path, FREQ_EXT);
    if (p_rle) p_rle_freq = add_ext(p_rle, FREQ_EXT);
    uint64_t *sizes = malloc(nb * sizeof *sizes), *rle_sizes = malloc(nb * sizeof *rle_sizes);
    uint64_t *offsets = malloc((nb + 1) * sizeof *offsets), *hist = malloc(nb * 256 * sizeof *hist);
    uint8_t *d_in = NULL, *d_rle = NULL, *h_rle = NULL;
    uint64_t *d_hist = NULL;
    bool use_rle = true;
    if (!p_rle || !p_rle_freq || !p_freq || !sizes || !rle_sizes || !offsets || !hist) {
        err = _LACK_OF_MEMORY;
        goto out;
    }
    offsets[0] = 0;
    for (uint64_t b = 0; b < nb; b++) {
        sizes[b] = (b == nb - 1) ? n - b * bs : bs;
        offsets[b + 1] = offsets[b] + sizes[b];
    }
    d_in = shafa_cuda_dev_alloc(g, n + 64);
    d_hist = shafa_cuda_dev_alloc(g, nb * 256 * sizeof(uint64_t));
    if (!d_in || !d_hist) {
        err = _LACK_OF_MEMORY;
        goto out;
    }
    if ((err = gpu_status(shafa_cuda_memcpy_h2d(g, d_in, in.data, n)))) goto out;

    { /* block 0 decides for the whole file (f.c:246-258) */
        uint64_t rle0 = 0;
        if ((err = gpu_status(shafa_cuda_rle_encode_blocks(g, d_in, sizes[0], bs, NULL, 0, &rle0, NULL)))) goto out;
        use_rle = shafa_cuda_rle_gate(sizes[0], rle0, force_rle) != 0;
    }
    if (use_rle) {
        const uint64_t cap = 2 * n + 3 * nb + 64; /* f.c:244 */
        d_rle = shafa_cuda_dev_alloc(g, cap);
        if (!d_rle) {
            err = _LACK_OF_MEMORY;
            goto out;
        }
        if ((err = gpu_status(shafa_cuda_rle_encode_blocks(g, d_in, n, bs, d_rle, cap, rle_sizes, d_hist)))) goto out;
        uint64_t tot = 0;
        for (uint64_t b = 0; b < nb; b++) tot += rle_sizes[b];
        h_rle = cli_host_alloc(tot + 16);
        if (!h_rle) {
            err = _LACK_OF_MEMORY;
            goto out;
        }
        if ((err = gpu_status(shafa_cuda_memcpy_d2h(g, h_rle, d_rle, tot)))) goto out;
        if ((err = gpu_status(shafa_cuda_memcpy_d2h(g, hist, d_hist, nb * 256 * sizeof(uint64_t))))) goto out;
        if ((err = gpu_status(shafa_cuda_synchronize(g)))) goto out;
        FILE *f = fopen(p_rle, "wb"); /* f.c:263,307 */
        if (!f) {
            err = _FILE_INACCESSIBLE;
            goto out;
        }
        const bool ok = fwrite(h_rle, 1, tot, f) == tot;
        if (fclose(f) != 0 || !ok) {
            err = _FILE_STREAM_FAILED;
            goto out;
        }
        if ((err = write_freq_file(p_rle_freq, 'R', nb, rle_sizes, hist))) goto out; /* f.c:289,312-314 */
    }
    if (!use_rle || force_freq) { /* f.c:278-295,322-331 */
        if ((err = gpu_status(shafa_cuda_histogram_blocks(g, d_in, offsets, nb, d_hist)))) goto out;
        if ((err = gpu_status(shafa_cuda_memcpy_d2h(g, hist, d_hist, nb * 256 * sizeof(uint64_t))))) goto out;
        if ((err = gpu_status(shafa_cuda_synchronize(g)))) goto out;
        if ((err = write_freq_file(p_freq, 'N', nb, sizes, hist))) goto out;
    }
out:
    if (d_in) shafa_cuda_dev_free(g, d_in);
    if (d_rle) shafa_cuda_dev_free(g, d_rle);
    if (d_hist) shafa_cuda_dev_free(g, d_hist);
    if (h_rle) cli_host_free(h_rle);
    blob_free(&in);
    if (!err) {
        const bool wrote_freq = !use_rle || force_freq;
        f_summary(nb, sizes, n, rle_sizes, now_ms() - t0, use_rle ? p_rle : NULL, wrote_freq ? p_freq : NULL, use_rle ? p_rle_freq : NULL);
        if (use_rle) { /* f.c:396-399 */
            free(*path);
            *path = p_rle;
            p_rle = NULL;
        }
    }
    free(p_rle);
    free(p_rle_freq);
    free(p_freq);
    free(sizes);
    free(rle_sizes);
    free(offsets);
    free(hist);
    return err;
}
