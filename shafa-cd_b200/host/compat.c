/*
 * compat.c -- the small host utilities the reference keeps in utils/extensions.c, utils/errors.c and
 * utils/multithread.c's NO_MULTITHREAD flag, for OUR driver (host/shafa.c).  Same contracts as
 * utils/extensions.h:19-35 and utils/errors.c:26-36.
 *
 * A maintainer who links the REFERENCE's own shafa.c + extensions.c + errors.c against host/mod_*.c (see
 * INTEGRATION.md section 1) leaves this file out and adds host/ref_shim.c instead, so nothing is defined twice.
 */
#include "util.h"

#include <stdlib.h>
#include <string.h>

/* ---- extensions ---------------------------------------------------------------- */

bool check_ext(const char *path, const char *ext)
{
    if (!path) return false;
    const size_t lp = strlen(path), le = strlen(ext);
    return lp >= le && strcmp(path + (lp - le), ext) == 0;
}

char *add_ext(const char *path, const char *ext)
{
    const size_t lp = strlen(path), le = strlen(ext);
    char *out = malloc(lp + le + 1);
    if (out) {
        memcpy(out, path, lp);
        memcpy(out + lp, ext, le + 1);
    }
    return out;
}

char *rm_ext(const char *path)
{
    const char *dot = strrchr(path, '.');
    const size_t keep = dot ? (size_t)(dot - path) : strlen(path);
    char *out = malloc(keep + 1);
    if (out) {
        memcpy(out, path, keep);
        out[keep] = '\0';
    }
    return out;
}

/* ---- errors ------------------------------------------------------------------------ */

const char *error_msg(int num)
{
    switch (num) {
    case _SUCCESS: return "No error\n";
    case _LACK_OF_MEMORY: return "Not enough memory for allocation\n";
    case _FILE_INACCESSIBLE: return "File can't be accessed. Either lack of permissions or file doesn't exist\n";
    case _FILE_UNRECOGNIZABLE: return "File not recognized\n";
    case _FILE_STREAM_FAILED: return "Can't communicate properly with file's stream\n";
    case _FILE_TOO_SMALL: return "File too small for decompression\n";
    case _THREAD_CREATION_FAILED: return "Thread couldn't be created\n";
    case _THREAD_TERMINATION_FAILED: return "Thread didn't terminate properly\n";
    default: return "Unknown error";
    }
}

