// startup_probe.cu -- where a `shafa` process spends its CUDA start-up: driver init, primary context, module load
// (first launch of a kernel from libshafa_cuda), first pinned allocation.  Prints one JSON line.
//   nvcc -O2 -o startup_probe tools/startup_probe.cu -lcuda -Ishafa-cd_b200/../include -Lshafa-cd_b200 -lshafa_cuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "shafa_cuda.h"

static double now_ms(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec / 1e6;
}

int main(void)
{
    const double t0 = now_ms();
    CUresult r = cuInit(0);
    const double t1 = now_ms();
    if (r != CUDA_SUCCESS) {
        printf("{\"error\": \"cuInit failed (%d)\"}\n", (int)r);
        return 1;
    }
    CUdevice dev;
    cuDeviceGet(&dev, 0);
    CUcontext pctx;
    cuDevicePrimaryCtxRetain(&pctx, dev);
    const double t2 = now_ms();
    cudaSetDevice(0);
    cudaFree(0);
    const double t3 = now_ms();
    shafa_cuda_ctx *g = NULL;
    shafa_cuda_create(0, &g);
    const double t4 = now_ms();
    // first kernel of the library: 4 KiB histogram (loads the module)
    uint8_t *d_in = (uint8_t *)shafa_cuda_dev_alloc(g, 4096 + 64);
    uint64_t *d_hist = (uint64_t *)shafa_cuda_dev_alloc(g, 256 * 8);
    uint64_t off[2] = {0, 4096};
    cudaMemset(d_in, 1, 4096);
    const double t5 = now_ms();
    shafa_cuda_histogram_blocks(g, d_in, off, 1, d_hist);
    shafa_cuda_synchronize(g);
    const double t6 = now_ms();
    void *pin = shafa_cuda_host_alloc(1ull << 30);
    const double t7 = now_ms();
    if (pin) memset(pin, 0, 1ull << 30);
    const double t8 = now_ms();
    shafa_cuda_host_free(pin);
    shafa_cuda_destroy(g);
    const double t9 = now_ms();
    const char *lazy = getenv("CUDA_MODULE_LOADING");
    printf("{\"cuInit_ms\": %.1f, \"primary_ctx_retain_ms\": %.1f, \"runtime_attach_ms\": %.1f, \"shafa_cuda_create_ms\": %.1f, "
           "\"dev_alloc_memset_ms\": %.1f, \"first_kernel_module_load_ms\": %.1f, \"pin_1GiB_ms\": %.1f, \"touch_1GiB_ms\": %.1f, "
           "\"teardown_ms\": %.1f, \"total_ms\": %.1f, \"CUDA_MODULE_LOADING\": \"%s\"}\n",
           t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5, t7 - t6, t8 - t7, t9 - t8, t9 - t0, lazy ? lazy : "(default)");
    return 0;
}
