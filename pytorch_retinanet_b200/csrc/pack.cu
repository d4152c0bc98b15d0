// Ragged ground truth -> packed device buffers in ONE launch.
// The reference hands the loss a python list of per-image dicts (retinanet/losses.py:126-128, collate_fn in
// utils/detection_utils.py:7-9); packing them with torch costs two `cat`s plus a host->device copy of the
// offsets per step.  Here the per-image pointers and counts travel in the kernel's parameter space (no H2D
// copy at all) and one launch writes gt_boxes [sumG,4], gt_labels [sumG] and gt_off [N+1].  An optional
// per-image (ratio_h, ratio_w) applies torchvision's resize_boxes (tv:models/detection/transform.py:306-319:
// x * ratio_w, y * ratio_h, one fp32 multiply each) while packing, which is what
// GeneralizedRCNNTransform.forward does to the targets right before the path (models.py:279).
#include "rn_common.cuh"

namespace {

constexpr int PACK_MAX_IMAGES = 128;     // pointers + counts + ratios of one launch fit the 4 KB parameter space

struct PackParams {
    const float4 *boxes[PACK_MAX_IMAGES];
    const long long *labels[PACK_MAX_IMAGES];
    int off[PACK_MAX_IMAGES + 1];        // offsets relative to this launch's first image
    int n;
    int base_off;                        // packed offset of this launch's first image
    int image0;                          // index of this launch's first image
    int has_ratio;
    float ratio_h[PACK_MAX_IMAGES];
    float ratio_w[PACK_MAX_IMAGES];
};

__global__ void __launch_bounds__(128) pack_targets_kernel(const __grid_constant__ PackParams P, float4 *__restrict__ out_boxes,
                                                           long long *__restrict__ out_labels, int *__restrict__ out_off) {
    const int i = blockIdx.x;            // image inside this launch
    const int o = P.off[i], cnt = P.off[i + 1] - o;
    if (threadIdx.x == 0) {
        out_off[P.image0 + i] = P.base_off + o;
        if (i == P.n - 1) out_off[P.image0 + P.n] = P.base_off + P.off[P.n];
    }
    const float rw = P.has_ratio ? P.ratio_w[i] : 1.0f, rh = P.has_ratio ? P.ratio_h[i] : 1.0f;
    for (int k = threadIdx.x; k < cnt; k += blockDim.x) {
        float4 b = P.boxes[i][k];
        if (P.has_ratio) {
            b.x = __fmul_rn(b.x, rw); b.z = __fmul_rn(b.z, rw);
            b.y = __fmul_rn(b.y, rh); b.w = __fmul_rn(b.w, rh);
        }
        out_boxes[P.base_off + o + k] = b;
        if (out_labels) out_labels[P.base_off + o + k] = P.labels[i][k];
    }
}

}  // namespace

extern "C" int rn_pack_targets(const float *const *boxes_host, const int64_t *const *labels_host, const int32_t *counts_host,
                               int N, const float *ratios_hw_host, float *out_boxes, int64_t *out_labels, int32_t *out_off,
                               rn_stream_t stream) {
    RN_CHECK_ARG(out_off && N >= 0 && (N == 0 || (boxes_host && counts_host)), RN_E_BADARG, "rn_pack_targets: bad argument");
    RN_CHECK_ARG(N == 0 || !out_labels || labels_host, RN_E_BADARG, "rn_pack_targets: labels requested without label pointers");
    cudaStream_t s = (cudaStream_t)stream;
    if (N == 0) {
        cudaError_t e = cudaMemsetAsync(out_off, 0, sizeof(int32_t), s);
        return e == cudaSuccess ? 0 : (int)e;
    }
    long long total = 0;
    for (int i0 = 0; i0 < N; i0 += PACK_MAX_IMAGES) {
        PackParams P;
        P.n = N - i0 < PACK_MAX_IMAGES ? N - i0 : PACK_MAX_IMAGES;
        P.base_off = (int)total;
        P.image0 = i0;
        P.has_ratio = ratios_hw_host != nullptr;
        P.off[0] = 0;
        for (int i = 0; i < P.n; ++i) {
            const int c = counts_host[i0 + i];
            RN_CHECK_ARG(c >= 0, RN_E_BADARG, "rn_pack_targets: negative count for image %d", i0 + i);
            RN_CHECK_ARG(c == 0 || (boxes_host[i0 + i] && (!out_labels || labels_host[i0 + i])), RN_E_BADARG,
                         "rn_pack_targets: null pointer for image %d", i0 + i);
            RN_CHECK_ARG(c == 0 || (((uintptr_t)boxes_host[i0 + i]) & 15) == 0, RN_E_BADARG,
                         "rn_pack_targets: boxes of image %d are not 16-byte aligned", i0 + i);
            P.boxes[i] = (const float4 *)boxes_host[i0 + i];
            P.labels[i] = out_labels ? (const long long *)labels_host[i0 + i] : nullptr;
            P.off[i + 1] = P.off[i] + c;
            P.ratio_h[i] = ratios_hw_host ? ratios_hw_host[2 * (i0 + i)] : 1.0f;
            P.ratio_w[i] = ratios_hw_host ? ratios_hw_host[2 * (i0 + i) + 1] : 1.0f;
        }
        total += P.off[P.n];
        RN_CHECK_ARG(total < 0x7fffffffLL, RN_E_TOOLARGE, "rn_pack_targets: more than 2^31 boxes");
        RN_CHECK_ARG(total == 0 || out_boxes, RN_E_BADARG, "rn_pack_targets: null output");
        pack_targets_kernel<<<P.n, 128, 0, s>>>(P, (float4 *)out_boxes, (long long *)out_labels, out_off);
        RN_CHECK_LAUNCH("rn_pack_targets");
    }
    return 0;
}
