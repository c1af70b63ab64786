#include "util.h"

#include <pthread.h>
#include <sys/mman.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ---- GPU context ------------------------------------------------------------------- */

static shafa_cuda_multi *g_multi;
static bool g_multi_tried;
static pthread_mutex_t g_multi_lock = PTHREAD_MUTEX_INITIALIZER;
static pthread_t g_prefetch_thread;
static bool g_prefetch_running;

/* "0-7", "0,2,5", "3" -> device list */
static int parse_devices(const char *spec, int *devs, int cap)
{
    int n = 0;
    const char *p = spec;
    while (*p && n < cap) {
        char *end;
        const long a = strtol(p, &end, 10);
        if (end == p || a < 0) return 0;
        long b = a;
        p = end;
        if (*p == '-') {
            b = strtol(p + 1, &end, 10);
            if (end == p + 1 || b < a) return 0;
            p = end;
        }
        for (long d = a; d <= b && n < cap; d++) devs[n++] = (int)d;
        if (*p == ',') p++;
        else if (*p) return 0;
    }
    return n;
}

shafa_cuda_multi *gpu_multi(void)
{
    pthread_mutex_lock(&g_multi_lock);
    if (!g_multi_tried) {
        g_multi_tried = true;
        int devs[64], n = 0;
        const char *spec = getenv("SHAFA_DEVICES");
        if (spec && *spec) n = parse_devices(spec, devs, 64);
        if (n == 0) {
            const char *dev = getenv("SHAFA_DEVICE");
            devs[0] = dev ? atoi(dev) : 0;
            n = 1;
        }
        if (shafa_cuda_multi_create(devs, n, &g_multi) != SHAFA_SUCCESS) g_multi = NULL;
    }
    shafa_cuda_multi *m = g_multi;
    pthread_mutex_unlock(&g_multi_lock);
    return m;
}

shafa_cuda_ctx *gpu(void)
{
    shafa_cuda_multi *m = gpu_multi();
    return m ? shafa_cuda_multi_ctx(m, 0) : NULL;
}

static void *prefetch_main(void *arg)
{
    (void)arg;
    gpu_multi();
    return NULL;
}

void gpu_prefetch(void)
{
    if (g_prefetch_running || g_multi_tried) return;
    g_prefetch_running = pthread_create(&g_prefetch_thread, NULL, prefetch_main, NULL) == 0;
}

void gpu_release(void)
{
    if (g_prefetch_running) {
        pthread_join(g_prefetch_thread, NULL);
        g_prefetch_running = false;
    }
    if (g_multi) shafa_cuda_multi_destroy(g_multi);
    g_multi = NULL;
}

_modules_error gpu_status(int rc)
{
    if (rc != SHAFA_SUCCESS && g_multi && getenv("SHAFA_VERBOSE"))
        fprintf(stderr, "libshafa_cuda: %s | %s\n", shafa_cuda_last_error(shafa_cuda_multi_ctx(g_multi, 0)), shafa_cuda_multi_last_error(g_multi));
    return (_modules_error)rc;
}

/* ---- files ----------------------------------------------------------------------------- */

/* Host buffers of the CLI are used once per process.  Page-locking a GiB costs several times what the
 * pageable copy costs over a pinned one, so they are plain page-aligned allocations unless SHAFA_CLI_PINNED
 * is set (a long-lived caller of the library should pin its buffers instead: that is what bench.py's e2e does). */
static bool cli_pinned(void)
{
    const char *e = getenv("SHAFA_CLI_PINNED");
    return e && *e && *e != '0';
}
void *cli_host_alloc(size_t bytes)
{
    if (cli_pinned()) return shafa_cuda_host_alloc(bytes);
    void *p = NULL;
    if (posix_memalign(&p, 2u << 20, bytes ? bytes : 1) != 0) return NULL;
#ifdef MADV_HUGEPAGE
    if (bytes >= (8u << 20)) madvise(p, bytes, MADV_HUGEPAGE); /* GiB-sized one-shot buffers: 512x fewer first-touch faults */
#endif
    return p;
}
void cli_host_free(void *p)
{
    if (!p) return;
    if (cli_pinned())
        shafa_cuda_host_free(p);
    else
        free(p);
}

_modules_error blob_read(const char *path, blob *b)
{
    b->data = NULL;
    b->size = 0;
    b->pinned = false;
    FILE *f = fopen(path, "rb");
    if (!f) return _FILE_INACCESSIBLE;
    if (fseeko(f, 0, SEEK_END) != 0) {
        fclose(f);
        return _FILE_STREAM_FAILED;
    }
    const off_t sz = ftello(f);
    if (sz < 0 || fseeko(f, 0, SEEK_SET) != 0) {
        fclose(f);
        return _FILE_STREAM_FAILED;
    }
    b->pinned = cli_pinned();
    b->data = cli_host_alloc((size_t)sz + 64);
    if (!b->data && b->pinned) { /* no CUDA device to pin against; the module will fail later */
        b->pinned = false;
        b->data = malloc((size_t)sz + 64);
    }
    if (!b->data) {
        fclose(f);
        return _LACK_OF_MEMORY;
    }
    if (fread(b->data, 1, (size_t)sz, f) != (size_t)sz) {
        fclose(f);
        blob_free(b);
        return _FILE_STREAM_FAILED;
    }
    memset(b->data + sz, 0, 64);
    b->size = (uint64_t)sz;
    fclose(f);
    return _SUCCESS;
}

void blob_free(blob *b)
{
    if (b->data) {
        if (b->pinned)
            shafa_cuda_host_free(b->data);
        else
            free(b->data);
    }
    b->data = NULL;
    b->size = 0;
}

/* ---- cursor -------------------------------------------------------------------------------- */

bool cur_lit(cursor *c, char ch)
{
    if (c->p < c->end && *c->p == (uint8_t)ch) {
        c->p++;
        return true;
    }
    return false;
}

bool cur_u64(cursor *c, uint64_t *v)
{
    if (c->p >= c->end || *c->p < '0' || *c->p > '9') return false;
    uint64_t x = 0;
    while (c->p < c->end && *c->p >= '0' && *c->p <= '9') x = x * 10 + (uint64_t)(*c->p++ - '0');
    *v = x;
    return true;
}

bool cur_field(cursor *c, const uint8_t **start, size_t *len)
{
    *start = c->p;
    while (c->p < c->end && *c->p != '@') c->p++;
    *len = (size_t)(c->p - *start);
    return *len > 0;
}

_modules_error cod_parse_fields(const uint8_t *p, size_t n, uint8_t *len, uint64_t *code)
{
    const uint8_t *end = p + n;
    for (int s = 0; s < 256; s++) {
        unsigned l = 0;
        uint64_t v = 0;
        while (p < end && (*p == '0' || *p == '1')) {
            if (++l > 64) return _FILE_UNRECOGNIZABLE;
            v = (v << 1) | (uint64_t)(*p++ == '1');
        }
        len[s] = (uint8_t)l;
        code[s] = v;
        if (s != 255) {
            if (p >= end || *p != ';') return _FILE_UNRECOGNIZABLE; /* c.c:127-131,165-169 */
            p++;
        }
    }
    return p == end ? _SUCCESS : _FILE_UNRECOGNIZABLE; /* c.c:174-177 */
}

size_t cod_max_code_bits(const uint8_t *p, size_t n)
{
    size_t best = 0, run = 0;
    for (size_t i = 0; i < n; i++) {
        if (p[i] == '0' || p[i] == '1') {
            if (++run > best) best = run;
        } else {
            run = 0;
        }
    }
    return best;
}

_modules_error cod_parse_fields_wide(const uint8_t *p, size_t n, uint8_t *len, uint64_t *code4)
{
    const uint8_t *end = p + n;
    for (int s = 0; s < 256; s++) {
        unsigned l = 0;
        uint64_t v[4] = {0, 0, 0, 0};
        while (p < end && (*p == '0' || *p == '1')) {
            if (++l > COD_WIDE_MAX_BITS) return _FILE_UNRECOGNIZABLE;
            v[3] = (v[3] << 1) | (v[2] >> 63); /* shift the 256-bit value left by one, new bit in at the bottom */
            v[2] = (v[2] << 1) | (v[1] >> 63);
            v[1] = (v[1] << 1) | (v[0] >> 63);
            v[0] = (v[0] << 1) | (uint64_t)(*p++ == '1');
        }
        len[s] = (uint8_t)l;
        for (int k = 0; k < 4; k++) code4[s * 4 + k] = v[k];
        if (s != 255) {
            if (p >= end || *p != ';') return _FILE_UNRECOGNIZABLE; /* c.c:127-131,165-169 */
            p++;
        }
    }
    return p == end ? _SUCCESS : _FILE_UNRECOGNIZABLE; /* c.c:174-177 */
}

bool block_count_plausible(uint64_t nb, uint64_t container_bytes) { return nb <= container_bytes / 258u; }

double now_ms(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec / 1e6;
}
