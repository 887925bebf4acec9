// Device-side batch assembly: the step BEFORE the encoder path.
//
// The reference builds every padded batch on the host, per sample, in Python (pad_tensors data/data.py:360-373,
// get_gather_index 376-384, _compute_ot_scatter / _compute_pad data/itm.py:264-278, _mask_img_feat /
// _get_feat_target / _get_img_tgt_mask data/mrm.py:22-39, _get_targets 213-218) and ships B x R x 8 KB of fp32
// region features through pinned memory every step.  On a 180 GB part the whole feature store of a dataset fits in
// HBM as a ragged arena (image i owns rows [row0_i, row0_i + nbb_i)), so a batch is described by a few integers per
// sample and assembled here: one HBM-bound row-copy kernel (features, boxes, soft labels) and one index kernel.
// Everything is copy / integer work: results are bit-identical to the host collate.
#include "common.cuh"

namespace uc2 {
namespace {

constexpr int PAD_WARPS = 8;

// one warp per output row (b, r): copy / zero-fill / divert to the compacted targets
template <bool VEC4>
__global__ void __launch_bounds__(PAD_WARPS * 32)
pad_rows_kernel(const float* __restrict__ arena, int D, const long long* __restrict__ row0,
                const int* __restrict__ nbb, const unsigned char* __restrict__ mask,
                const int* __restrict__ tgt_slot, int B, int R, int zero_masked, float* __restrict__ out,
                float* __restrict__ targets) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * PAD_WARPS + (threadIdx.x >> 5);
    if (row >= (long long)B * R) return;
    const int b = (int)(row / R), r = (int)(row % R);
    const bool valid = r < nbb[b];
    const bool masked = valid && mask && mask[row];
    const float* src = arena + (row0[b] + r) * D;
    float* dst = out ? out + row * D : nullptr;
    float* tgt = (masked && targets) ? targets + (long long)tgt_slot[row] * D : nullptr;
    const bool keep = valid && !(masked && zero_masked);
    if (VEC4) {
        for (int c = lane * 4; c < D; c += 128) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) v = *reinterpret_cast<const float4*>(src + c);
            if (tgt) *reinterpret_cast<float4*>(tgt + c) = v;
            if (dst) *reinterpret_cast<float4*>(dst + c) = keep ? v : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    } else {
        for (int c = lane; c < D; c += 32) {
            const float v = valid ? src[c] : 0.f;
            if (tgt) tgt[c] = v;
            if (dst) dst[c] = keep ? v : 0.f;
        }
    }
}

// one thread per (b, j), j < max(S, T, R)
__global__ void __launch_bounds__(256)
batch_index_kernel(const int* __restrict__ tl, const int* __restrict__ nbb, const unsigned char* __restrict__ img_mask,
                   int B, int T, int R, int S, long long* __restrict__ attn, long long* __restrict__ gather,
                   long long* __restrict__ ot_scatter, unsigned char* __restrict__ txt_pad,
                   unsigned char* __restrict__ img_pad, unsigned char* __restrict__ img_mask_tgt) {
    int W = S > T ? S : T;
    if (R > W) W = R;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * W) return;
    const int b = (int)(i / W), j = (int)(i % W);
    const int t = tl[b], n = nbb[b];
    if (j < S) {
        const bool img = j >= t && j < t + n;
        if (attn) attn[(long long)b * S + j] = j < t + n ? 1 : 0;
        // get_gather_index: identity, except the image block [tl, tl + nbb) which reads from T + (j - tl)
        if (gather) gather[(long long)b * S + j] = img ? j - t + T : j;
        // _compute_ot_scatter: text rows stay, everything from tl on lands at T + (j - tl)
        if (ot_scatter) ot_scatter[(long long)b * S + j] = j >= t ? j - t + T : j;
        // _get_img_tgt_mask: zeros(tl) ++ img_mask, zero padded to S
        if (img_mask_tgt) img_mask_tgt[(long long)b * S + j] = (img && img_mask[(long long)b * R + (j - t)]) ? 1 : 0;
    }
    if (txt_pad && j < T) txt_pad[(long long)b * T + j] = j >= t ? 1 : 0;
    if (img_pad && j < R) img_pad[(long long)b * R + j] = j >= n ? 1 : 0;
}

}  // namespace
}  // namespace uc2

using namespace uc2;

extern "C" UC2_API int uc2_pad_rows(const uc2_pad_args* a, float* out, float* targets, void* stream) {
    if (int rc = require_sm100()) return rc;
    UC2_REQUIRE(a && (out || targets) && a->arena && a->row0 && a->nbb, UC2_ERR_ARG, "pad_rows: null pointer");
    UC2_REQUIRE(a->B > 0 && a->R > 0 && a->D > 0, UC2_ERR_ARG, "pad_rows: bad shape B=%d R=%d D=%d", a->B, a->R, a->D);
    UC2_REQUIRE(!targets || (a->mask && a->tgt_slot), UC2_ERR_ARG, "pad_rows: targets need mask and tgt_slot");
    const long long rows = (long long)a->B * a->R;
    const unsigned blocks = (unsigned)((rows + PAD_WARPS - 1) / PAD_WARPS);
    const bool vec = a->D % 4 == 0 && aligned16(a->arena) && aligned16(out) && aligned16(targets);
    if (vec)
        pad_rows_kernel<true><<<blocks, PAD_WARPS * 32, 0, (cudaStream_t)stream>>>(
            a->arena, a->D, a->row0, a->nbb, a->mask, a->tgt_slot, a->B, a->R, a->zero_masked, out, targets);
    else
        pad_rows_kernel<false><<<blocks, PAD_WARPS * 32, 0, (cudaStream_t)stream>>>(
            a->arena, a->D, a->row0, a->nbb, a->mask, a->tgt_slot, a->B, a->R, a->zero_masked, out, targets);
    return check_last("pad_rows_kernel");
}

extern "C" UC2_API int uc2_batch_index(const int* txt_lens, const int* num_bbs, const unsigned char* img_mask, int B,
                                       int T, int R, int S, long long* attn_masks, long long* gather_index,
                                       long long* ot_scatter, unsigned char* txt_pad, unsigned char* img_pad,
                                       unsigned char* img_mask_tgt, void* stream) {
    if (int rc = require_sm100()) return rc;
    UC2_REQUIRE(txt_lens && num_bbs && B > 0 && T >= 0 && R >= 0 && S > 0, UC2_ERR_ARG, "batch_index: bad args");
    UC2_REQUIRE(!img_mask_tgt || img_mask, UC2_ERR_ARG, "batch_index: img_mask_tgt needs img_mask");
    int W = S > T ? S : T;
    if (R > W) W = R;
    const long long n = (long long)B * W;
    batch_index_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        txt_lens, num_bbs, img_mask, B, T, R, S, attn_masks, gather_index, ot_scatter, txt_pad, img_pad, img_mask_tgt);
    return check_last("batch_index_kernel");
}
