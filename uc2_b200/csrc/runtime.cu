// Library-wide runtime helpers: error string, device checks, launch counter.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace uc2 {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_last(const char* what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
        return UC2_ERR_CUDA;
    }
    return UC2_OK;
}

static int g_sms[64];
static int g_cc[64];

static int dev_query(int* sms, int* cc) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
    if (g_sms[dev] == 0) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return -1;
        g_cc[dev] = prop.major * 10 + prop.minor;
        g_sms[dev] = prop.multiProcessorCount;
    }
    *sms = g_sms[dev];
    *cc = g_cc[dev];
    return 0;
}

int num_sms() {
    int sms = 148, cc = 0;
    dev_query(&sms, &cc);
    return sms;
}

int require_sm100() {
    int sms = 0, cc = 0;
    if (dev_query(&sms, &cc) != 0) {
        set_error("no usable CUDA device (uc2_b200 has no CPU fallback)");
        return UC2_ERR_CUDA;
    }
    if (cc != 100) {
        set_error("uc2_b200 kernels are built for sm_100a only; device reports sm_%d", cc);
        return UC2_ERR_ARCH;
    }
    return UC2_OK;
}

}  // namespace uc2

extern "C" UC2_API const char* uc2_last_error(void) { return uc2::g_err; }
extern "C" UC2_API int uc2_version(void) { return 100; }
extern "C" UC2_API long long uc2_launch_count(void) { return uc2::g_launches.load(); }
