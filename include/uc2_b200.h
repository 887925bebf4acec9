/* uc2_b200 -- C ABI of the B200 (sm_100a) kernels behind the UC2 cross-modal encoder path.
 *
 * The reference (zmykevin/UC2) is pure Python/PyTorch and has no FFI of its own; each entry
 * point below names the reference function (file:line under the reference tree) whose device
 * work it replaces.  Conventions:
 *   - plain pointers + sizes, no torch types; every pointer is a DEVICE pointer unless noted;
 *   - the caller owns every buffer (inputs, outputs, saved-for-backward, workspace);
 *   - every function returns 0 on success or a negative UC2_ERR_* code and never throws;
 *     uc2_last_error() returns a message for the calling thread's last failure;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing syncs;
 *   - bf16 tensors are row-major uint16 storage (__nv_bfloat16), fp32 tensors are float.
 */
#ifndef UC2_B200_H
#define UC2_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(UC2_BUILD)
#define UC2_API __attribute__((visibility("default")))
#else
#define UC2_API
#endif

#define UC2_OK 0
#define UC2_ERR_ARG (-1)      /* bad shape / null / misaligned pointer */
#define UC2_ERR_ARCH (-2)     /* device is not sm_100 */
#define UC2_ERR_CUDA (-3)     /* a CUDA runtime/driver call failed */
#define UC2_ERR_UNSUPPORTED (-4)

UC2_API const char* uc2_last_error(void);
UC2_API int uc2_version(void);
/* Number of kernels this library has launched in the calling process (bench.py gpu_launches). */
UC2_API long long uc2_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * Dense layers: tcgen05/TMEM GEMM fed by TMA.
 * Replaces every nn.Linear on the path: model/layer.py:76-78 (Q,K,V), 112 (attention out),
 * 140 (FFN1), 153 (FFN2); model/model.py:359 (img_linear), heads 1153-1169; and their autograd
 * backward (dgrad / wgrad).
 *
 *   D[M,N] = A[M,K] * B[N,K]^T   (+ epilogue)
 * a_mn != 0 : A is stored transposed, i.e. a[K][M] row-major with leading dimension lda (wgrad).
 * b_mn != 0 : B is stored transposed, i.e. b[K][N] row-major with leading dimension ldb (dgrad).
 * Epilogue (applied in this order, every pointer optional unless noted):
 *   acc += bias[n] (fp32)                       -- nn.Linear bias
 *   if out_pre : out_pre[m,n] = bf16(acc)       -- pre-activation copy (saved for backward)
 *   if act == UC2_ACT_GELU : acc = gelu_erf(acc)          (model/layer.py:31-37)
 *   if act == UC2_ACT_DGELU: acc *= gelu_erf'(aux[m,n])   (backward of the above; aux = pre-act)
 *   if act == UC2_ACT_TANH : acc = tanh(acc)              (model/layer.py:184)
 *   acc += residual[m,n] (bf16)                 -- BertSelfOutput / BertOutput residual
 *   out_bf16[m,n] = bf16(acc)  and/or  out_f32[m,n] (= or +=, see accumulate) acc
 * split_k > 1 splits K over CTAs and atomically accumulates into out_f32 (requires accumulate=1,
 * out_f32 only, no bias/act/residual): the wgrad path, which also gives gradient accumulation.
 */
#define UC2_ACT_NONE 0
#define UC2_ACT_GELU 1
#define UC2_ACT_DGELU 2
#define UC2_ACT_TANH 3

typedef struct {
    const void* a; long long lda; int a_mn;
    const void* b; long long ldb; int b_mn;
    int M, N, K;
    const float* bias;
    const void* residual; long long ld_res;
    const void* aux; long long ld_aux;
    int act;
    void* out_bf16; long long ld_out;
    void* out_pre; long long ld_pre;
    float* out_f32; long long ld_f32;
    int accumulate;   /* out_f32 += acc (atomic) instead of = */
    int split_k;      /* >= 1 */
    int block_n;      /* 0 = auto; else 64, 128 or 256 */
} uc2_gemm_args;

UC2_API int uc2_gemm_bf16(const uc2_gemm_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UC2_B200_H */
