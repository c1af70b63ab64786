/*
 * fused.c -- modules F, T and C of ONE invocation (`shafa <file>`, shafa.c:293-298) as one pass over the data.
 *
 * The reference runs the three modules back to back through the disk: F writes .rle/.freq, T re-reads the
 * .freq, C re-reads the .rle and the .cod it was just given (c.c:329-392).  When one process runs all three,
 * the symbol stream, the histograms and the codes stay on the GPU(s) (shafa_cuda_encode_host_multi: chunks of
 * whole blocks through H2D -> RLE+histogram -> codes -> bit packing -> D2H on worker streams, one file sharded
 * over every device of $SHAFA_DEVICES), and .rle / .rle.freq / .rle.cod / .rle.shaf (or .freq / .cod / .shaf in
 * N mode) are written once, as side outputs, byte-identical to what the three separate modules write.
 * CUDA start-up runs on a background thread while the input file is being read; the two large outputs
 * (.rle and .shaf) are written by two threads.
 */
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "util.h"

typedef struct {
    const char *path;
    const uint8_t *data;
    uint64_t n;
    _modules_error err;
} write_job;

static void *write_job_main(void *arg)
{
    write_job *j = arg;
    j->err = write_all(j->path, j->data, j->n);
    return NULL;
}

_modules_error shafa_encode_fused(char **const path, const bool force_rle, const bool force_freq, const unsigned long block_size_in)
{
    const double t0 = now_ms();
    uint64_t bs = block_size_in;
    if (bs > MIB_64) return _FILE_STREAM_FAILED; /* utils/file.c:61 */
    if (bs == 0) bs = 524288;                    /* utils/file.c:62 */
    if (bs < 512) bs = 512;                      /* utils/file.c:63 */

    gpu_prefetch(); /* context creation overlaps the read below */
    blob in;
    _modules_error err = blob_read(*path, &in);
    if (err) return err;
    if (in.size < KIB_1) { /* f.c:220,366 */
        blob_free(&in);
        return _FILE_TOO_SMALL;
    }
    shafa_cuda_multi *m = gpu_multi();
    if (!m) {
        blob_free(&in);
        return _THREAD_CREATION_FAILED;
    }
    const double t_read = now_ms();
    shafa_cuda_encoded enc;
    memset(&enc, 0, sizeof enc);
    err = gpu_status(shafa_cuda_encode_host_multi(m, in.data, in.size, bs, force_rle, force_freq, &enc));
    blob_free(&in);
    if (err) return err;
    const double t_gpu = now_ms();

    const bool rle = enc.rle_mode != 0;
    const uint64_t nb = enc.n_blocks;
    char *p_rle = add_ext(*path, RLE_EXT), *p_freq = add_ext(*path, FREQ_EXT);
    char *p_sym = rle ? p_rle : *path; /* what T and C work on */
    char *p_sym_freq = p_rle ? add_ext(p_sym, FREQ_EXT) : NULL, *p_cod = p_rle ? add_ext(p_sym, CODES_EXT) : NULL,
         *p_shaf = p_rle ? add_ext(p_sym, SHAFA_EXT) : NULL;
    if (!p_rle || !p_freq || !p_sym_freq || !p_cod || !p_shaf) {
        err = _LACK_OF_MEMORY;
        goto out;
    }
    {
        /* the .rle stream on its own thread, the rest here */
        write_job rle_job = {p_rle, enc.rle, enc.rle_total, _SUCCESS};
        pthread_t th;
        bool threaded = false;
        if (rle) threaded = pthread_create(&th, NULL, write_job_main, &rle_job) == 0;
        if (rle && !threaded) write_job_main(&rle_job);
        const char mode = rle ? 'R' : 'N';
        err = write_freq_file(p_sym_freq, mode, nb, enc.sym_sizes, enc.hist); /* f.c:289,312-314 / 327-329 */
        if (!err && rle && force_freq) err = write_freq_file(p_freq, 'N', nb, enc.block_sizes, enc.hist_orig); /* f.c:278-295 */
        if (!err) err = write_cod_file(p_cod, mode, nb, enc.sym_sizes, enc.code_len, enc.code_bits);
        if (!err) err = write_shaf_file(p_shaf, nb, enc.z, enc.payload, enc.payload_offsets);
        if (threaded) pthread_join(th, NULL);
        if (!err) err = rle_job.err;
    }
    if (!err) {
        const double t_end = now_ms();
        /* the three modules' summaries, in the order the reference prints them */
        f_summary(nb, enc.block_sizes, enc.n, enc.sym_sizes, t_gpu - t0, rle ? p_rle : NULL, (!rle || force_freq) ? p_freq : NULL,
                  rle ? p_sym_freq : NULL);
        t_summary(nb, enc.sym_sizes, 0.0, p_cod);
        c_summary(nb, enc.sym_sizes, enc.z, t_end - t_gpu, p_shaf);
        if (getenv("SHAFA_VERBOSE"))
            fprintf(stderr, "fused F+T+C: read %.0f ms (CUDA start-up overlapped), GPU pipeline %.0f ms, writes %.0f ms, %d device(s)\n",
                    t_read - t0, t_gpu - t_read, t_end - t_gpu, shafa_cuda_multi_device_count(m));
        free(*path); /* f.c:396-399 then c.c:460-461 */
        *path = p_shaf;
        p_shaf = NULL;
        if (rle) p_sym = NULL; /* (alias of p_rle, freed below) */
    }
out:
    shafa_cuda_encoded_free(&enc);
    free(p_rle);
    free(p_freq);
    free(p_sym_freq);
    free(p_cod);
    free(p_shaf);
    return err;
}
