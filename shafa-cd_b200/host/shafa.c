/*
 * shafa.c -- the `shafa` command line: same options, defaults, module order,
 * extension checks and messages as the reference driver
 * (/root/reference/src/shafa.c:47-132 parse, 150-259 execute_modules, 262-317
 * main), over the CUDA-backed modules of this directory.
 *
 *   shafa <file> [-m f|t|c|d]... [-b K|m|M] [-c r|f] [-d s|r] [--no-multithread]
 *
 * --no-multithread is accepted and ignored: there are no per-block host threads
 * any more, every module processes all blocks in one batched GPU call.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "util.h"

typedef struct {
    unsigned long block_size;
    bool f, t, c, d;          /* -m */
    bool force_rle, force_freq; /* -c r / -c f */
    bool d_shaf, d_rle;       /* -d s / -d r */
} options;

/* shafa.c:47-132: every option is "-x v" with a one-character value */
static bool parse_args(int argc, char *const argv[], options *o, const char **file)
{
    for (int i = 1; i < argc; i++) {
        const char *a = argv[i];
        if (strcmp(a, "--no-multithread") == 0) continue;
        if (a[0] != '-') {
            if (*file) return false; /* second path */
            *file = a;
            continue;
        }
        if (++i >= argc) return false;
        const char *v = argv[i];
        if (strlen(a) != 2 || strlen(v) != 1) return false;
        switch (a[1]) {
        case 'm':
            if (*v == 'f') o->f = true;
            else if (*v == 't') o->t = true;
            else if (*v == 'c') o->c = true;
            else if (*v == 'd') o->d = true;
            else return false;
            break;
        case 'b':
            if (*v == 'K') o->block_size = KIB_640;
            else if (*v == 'm') o->block_size = MIB_8;
            else if (*v == 'M') o->block_size = MIB_64;
            else return false;
            break;
        case 'c':
            if (*v == 'r') o->force_rle = true;
            else if (*v == 'f') o->force_freq = true;
            else return false;
            break;
        case 'd':
            if (*v == 's') o->d_shaf = true;
            else if (*v == 'r') o->d_rle = true;
            else return false;
            break;
        default:
            return false;
        }
    }
    return true;
}

/* shafa.c:150-259: fixed order F, T, C, D; each module renames *file */
static _modules_error run_modules(const options *o, char **file)
{
    _modules_error e;
    bool f = o->f, t = o->t, c = o->c;
    if (f && t && c && !getenv("SHAFA_NO_FUSE")) {
        /* F, T and C in one invocation: one pass over the data, the .rle/.freq/.cod/.shaf files as side outputs */
        if ((e = shafa_encode_fused(file, o->force_rle, o->force_freq, o->block_size))) {
            fputs("Module f/t/c: Something went wrong while compressing...\n", stderr);
            return e;
        }
        f = t = c = false;
    }
    if (f) {
        if ((e = freq_rle_compress(file, o->force_rle, o->force_freq, o->block_size))) {
            fputs("Module f: Something went wrong while compressing with RLE or creating frequencies' table...\n", stderr);
            return e;
        }
    }
    if (t) {
        if (!o->f) { /* stand-alone T is given "<x>.freq" (shafa.c:167-181) */
            if (!check_ext(*file, FREQ_EXT)) {
                fprintf(stderr, "Module t: Wrong extension... Should end in %s\n", FREQ_EXT);
                return _OUTSIDE_MODULE;
            }
            char *stem = rm_ext(*file);
            if (!stem) return _LACK_OF_MEMORY;
            free(*file);
            *file = stem;
        }
        if ((e = get_shafa_codes(*file))) {
            fputs("Module t: Something went wrong...\n", stderr);
            return e;
        }
    }
    if (c) {
        if (o->f && !o->t) {
            fputs("Module c: Can't execute module 'c' after 'f' without 't'...\n", stderr);
            return _OUTSIDE_MODULE;
        }
        if ((e = shafa_compress(file))) {
            fputs("Module c: Something went wrong...\n", stderr);
            return e;
        }
    }
    if (o->d) {
        if ((o->f && (!o->t || !o->c) && !check_ext(*file, RLE_EXT)) || (o->t && !o->c)) { /* shafa.c:208 */
            fputs("Module d: Can't execute module 'd' after 'f' without 't' or 'c', nor execute it after 't'  without 'c'...\n", stderr);
            return _OUTSIDE_MODULE;
        }
        bool done = false;
        if (o->d_shaf || !o->d_rle) { /* plain "-m d" or "-d s" (shafa.c:213-241) */
            if (!check_ext(*file, SHAFA_EXT)) {
                if (o->d_shaf) {
                    fprintf(stderr, "Module d: Wrong extension... Should end in %s\n", SHAFA_EXT);
                    return _OUTSIDE_MODULE;
                }
            } else {
                const bool is_rle_shaf = check_ext(*file, RLE_EXT SHAFA_EXT);
                if (o->d_rle && !is_rle_shaf) {
                    fprintf(stderr, "Module d: Wrong extension... Should end in %s\n", RLE_EXT SHAFA_EXT);
                    return _OUTSIDE_MODULE;
                }
                if ((e = shafa_decompress(file, (o->d_rle || !o->d_shaf) && is_rle_shaf))) {
                    fputs("Module d: Something went wrong while decompressing...\n", stderr);
                    return e;
                }
                done = true;
            }
        }
        if (!done && (o->d_rle || !o->d_shaf)) { /* RLE only (shafa.c:243-256) */
            if (!check_ext(*file, RLE_EXT)) {
                fprintf(stderr, "Module d: Wrong extension... Should end in %s\n", RLE_EXT);
                return _OUTSIDE_MODULE;
            }
            if ((e = rle_decompress(file))) {
                fputs("Module d: Something went wrong while decompressing...\n", stderr);
                return e;
            }
        }
    }
    return _SUCCESS;
}

int main(int argc, char *const argv[])
{
    options o;
    memset(&o, 0, sizeof o);
    const char *arg_file = NULL;
    if (argc <= 1) {
        fputs("No file input\n", stderr);
        return 1;
    }
    if (!parse_args(argc, argv, &o, &arg_file)) {
        fputs("Wrong Options' syntax\n", stderr);
        return 1;
    }
    if (!arg_file) {
        fputs("No file input\n", stderr);
        return 1;
    }
    char *file = add_ext(arg_file, ""); /* modules free/replace it */
    if (!file) {
        fputs("Not enough memory\n", stderr);
        return 1;
    }
    if (!o.f && !o.t && !o.c && !o.d) { /* shafa.c:293-298 */
        if (check_ext(file, SHAFA_EXT))
            o.d = true;
        else
            o.f = o.t = o.c = true;
    }
    if (!o.block_size) o.block_size = KIB_64; /* shafa.c:304-305 */
    const _modules_error e = run_modules(&o, &file);
    free(file);
    gpu_release();
    if (e) {
        if (e != _OUTSIDE_MODULE) fputs(error_msg(e), stderr);
        return 1;
    }
    return 0;
}
