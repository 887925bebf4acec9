// Library-wide runtime helpers: error string, device checks, launch counter.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <vector>

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

// SMs the persistent kernels size their grids for: all of them, minus the ones set aside for communication kernels
// that run concurrently (uc2_reserve_sms; data-parallel training hands NCCL a few CTAs -- a persistent GEMM whose grid
// counts on every SM would otherwise wait for whichever SMs a collective is holding)
static std::atomic<int> g_reserved_sms{-1};
int num_sms() {
    int sms = 148, cc = 0;
    dev_query(&sms, &cc);
    int r = g_reserved_sms.load(std::memory_order_relaxed);
    if (r < 0) {
        const char* e = getenv("UC2_RESERVE_SMS");
        r = e ? atoi(e) : 0;
        if (r < 0 || r > sms / 2) r = 0;
        g_reserved_sms.store(r, std::memory_order_relaxed);
    }
    return sms - r;
}

// ---------------------------------------------------------------------------------------------
// TMA descriptors
// ---------------------------------------------------------------------------------------------
EncodeTiledFn tensormap_encoder() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(f);
    });
    return fn;
}

// 2-D bf16 tensor [rows][cols] with row pitch ld elements; box = [box_rows][64 cols], SWIZZLE_128B.
int make_tmap(CUtensorMap* m, const void* base, long long rows, long long cols, long long ld, int box_rows) {
    EncodeTiledFn enc = tensormap_encoder();
    UC2_REQUIRE(enc != nullptr, UC2_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
    cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    cuuint64_t gstr[1] = {static_cast<cuuint64_t>(ld) * 2};
    cuuint32_t box[2] = {64u, static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    UC2_REQUIRE(r == CUDA_SUCCESS, UC2_ERR_CUDA,
                "cuTensorMapEncodeTiled failed (%d): rows=%lld cols=%lld ld=%lld box_rows=%d base=%p", (int)r, rows,
                cols, ld, box_rows, base);
    return UC2_OK;
}

// dynamic tile scheduler of the persistent GEMM (gemm_tcgen05.cu: SCHED_R)
static std::atomic<int> g_gemm_dynamic{-1};
bool gemm_sched_dynamic() {
    int v = g_gemm_dynamic.load(std::memory_order_relaxed);
    if (v < 0) {
        const char* e = getenv("UC2_GEMM_SCHED");
        v = (e && e[0] == 'd') ? 1 : 0;
        g_gemm_dynamic.store(v, std::memory_order_relaxed);
    }
    return v == 1;
}

bool pdl_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("UC2_NO_PDL");
        on = (e && e[0] == '1') ? 0 : 1;
    }
    return on == 1;
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

// ---------------------------------------------------------------------------------------------
// optional per-launch timing with CUDA events on the launching stream (bench.py roofline numbers)
// ---------------------------------------------------------------------------------------------
static bool g_prof_on = false;
struct ProfRec { cudaEvent_t e0, e1; int kind; double work; };
static std::vector<ProfRec> g_prof;
static std::mutex g_prof_mu;

ProfScope::ProfScope(cudaStream_t s, int kind, double work) : stream_(s), idx_(-1) {
    if (!g_prof_on) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    ProfRec r;
    if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return;
    r.kind = kind;
    r.work = work;
    cudaEventRecord(r.e0, s);
    g_prof.push_back(r);
    idx_ = (int)g_prof.size() - 1;
}
ProfScope::~ProfScope() {
    if (idx_ < 0) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    cudaEventRecord(g_prof[idx_].e1, stream_);
}

}  // namespace uc2

extern "C" UC2_API int uc2_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(uc2::g_prof_mu);
    uc2::g_prof_on = on != 0;
    return UC2_OK;
}

// Sums the recorded launches per kind (0 = GEMM, 1 = attention, 2 = other) and clears the records.
extern "C" UC2_API int uc2_profile_collect(double* ms_by_kind, double* work_by_kind, int* launches_by_kind, int kinds) {
    using namespace uc2;
    if (cudaDeviceSynchronize() != cudaSuccess) { set_error("profile_collect: sync failed"); return UC2_ERR_CUDA; }
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (int k = 0; k < kinds; ++k) { ms_by_kind[k] = 0; work_by_kind[k] = 0; launches_by_kind[k] = 0; }
    for (auto& r : g_prof) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.e0, r.e1);
        if (r.kind >= 0 && r.kind < kinds) {
            ms_by_kind[r.kind] += ms; work_by_kind[r.kind] += r.work; launches_by_kind[r.kind] += 1;
        }
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    g_prof.clear();
    return UC2_OK;
}

extern "C" UC2_API const char* uc2_last_error(void) { return uc2::g_err; }
extern "C" UC2_API int uc2_version(void) { return 100; }
extern "C" UC2_API long long uc2_launch_count(void) { return uc2::g_launches.load(); }

// Test support: n_ctas CTAs that each take a whole SM (200 KB of shared memory) and spin for `cycles` clocks -- what a
// communication kernel holding SMs looks like to the persistent kernels (tests/test_gemm_gpu.py).
namespace uc2 {
__global__ void __launch_bounds__(128) occupy_sms_kernel(long long cycles, unsigned int* sink) {
    extern __shared__ unsigned int occ_smem[];
    occ_smem[threadIdx.x] = threadIdx.x;
    const long long t0 = clock64();
    while (clock64() - t0 < cycles) __nanosleep(200);
    if (sink && occ_smem[threadIdx.x] == 0xFFFFFFFFu) *sink = 1u;
}
}  // namespace uc2

extern "C" UC2_API int uc2_debug_occupy_sms(int n_ctas, long long cycles, void* stream) {
    using namespace uc2;
    UC2_REQUIRE(n_ctas > 0 && n_ctas <= 1024 && cycles >= 0 && cycles <= 4000000000LL, UC2_ERR_ARG,
                "debug_occupy_sms: bad args");
    static bool attr_set = false;
    const int smem = 200 * 1024;
    if (!attr_set) {
        UC2_CUDA(cudaFuncSetAttribute(occupy_sms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set = true;
    }
    occupy_sms_kernel<<<n_ctas, 128, smem, (cudaStream_t)stream>>>(cycles, nullptr);
    return check_last("occupy_sms_kernel");
}

extern "C" UC2_API int uc2_gemm_sched_dynamic(int on) {
    const int prev = uc2::gemm_sched_dynamic() ? 1 : 0;
    uc2::g_gemm_dynamic.store(on ? 1 : 0, std::memory_order_relaxed);
    return prev;
}

extern "C" UC2_API int uc2_reserve_sms(int n) {
    const int prev = uc2::g_reserved_sms.load(std::memory_order_relaxed);
    uc2::g_reserved_sms.store(n < 0 ? 0 : (n > 64 ? 64 : n & ~1), std::memory_order_relaxed);
    return prev < 0 ? 0 : prev;
}
