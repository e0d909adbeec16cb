// api.cu -- library-level entry points and error plumbing of libctgan_sm100.
#include <stdarg.h>
#include <string.h>
#include "common.cuh"

namespace ctgan {

static thread_local char g_err[512] = "";
unsigned long long g_kernel_launches = 0;
int g_pdl = 1;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_status(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return 0;
    set_error("%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
    return (int)e;
}

int sm_count() {
    static int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        else
            cached = 148;
    }
    return cached;
}

int elementwise_grid(int64_t work_items, int threads) {
    int64_t blocks = (work_items + threads - 1) / threads;
    int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

}  // namespace ctgan

extern "C" int ctgan_version(void) { return 100; }

extern "C" unsigned long long ctgan_kernel_launches(void) { return ctgan::g_kernel_launches; }
extern "C" void ctgan_set_pdl(int on) { ctgan::g_pdl = on != 0; }

extern "C" const char* ctgan_last_error(void) { return ctgan::g_err; }

extern "C" int ctgan_tc_available(void) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
    return major == 10 ? 1 : 0;
}
