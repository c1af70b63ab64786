/*
 * modules.h -- the five module entry points of the `shafa` host program.
 *
 * Names, argument meaning, path ownership and error numbers are the
 * reference's (/root/reference/src/modules/{f,t,c,d}.h, utils/errors.h), so
 * that a driver written against the reference (its shafa.c:150-259) links
 * against these unchanged.  The bodies are new: they do the file I/O and the
 * text containers on the host and hand every per-block computation to
 * libshafa_cuda (include/shafa_cuda.h).  There is no CPU implementation behind
 * them: without a CUDA device they return an error.
 */
#ifndef SHAFA_HOST_MODULES_H
#define SHAFA_HOST_MODULES_H

#include <stdbool.h>

typedef enum _error { /* utils/errors.h:5-16 */
    _SUCCESS = 0,
    _OUTSIDE_MODULE = 1,
    _LACK_OF_MEMORY = 2,
    _FILE_INACCESSIBLE = 3,
    _FILE_UNRECOGNIZABLE = 4,
    _FILE_STREAM_FAILED = 5,
    _FILE_TOO_SMALL = 6,
    _THREAD_CREATION_FAILED = 7,
    _THREAD_TERMINATION_FAILED = 8,
} _modules_error;

const char *error_msg(int num); /* utils/errors.c:26-36 */

/* f.h:16.  On success in R mode *path becomes "<path>.rle" (f.c:396-399).
 * block_size: 0 = the reference's default (524288, utils/file.c:62), values below 512 are raised to 512
 * (utils/file.c:63) and values above 64 MiB are refused, like the reference.  ONE DEVIATION: the block
 * size must be a multiple of 16 (every size the CLI can select -- 64 KiB, K, m, M -- is); any other value
 * returns _FILE_STREAM_FAILED with a message that says so, instead of being accepted as fsize() would. */
_modules_error freq_rle_compress(char **path, bool force_rle, bool force_freq, unsigned long block_size);
/* t.h:11.  Reads "<path>.freq", writes "<path>.cod". */
_modules_error get_shafa_codes(const char *path);
/* c.h:11.  Reads "<path>" and "<path>.cod", writes "<path>.shaf"; *path becomes it. */
_modules_error shafa_compress(char **path);
/* d.h:14.  "<x>.shaf" (+ "<x>.cod") -> "<x>" or, with decompress_rle, "<x minus .rle>". */
_modules_error shafa_decompress(char **path, bool decompress_rle);
/* d.h:22.  "<x>.rle" (+ "<x>.rle.freq") -> "<x>". */
_modules_error rle_decompress(char **path);

#endif
