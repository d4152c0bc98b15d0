// Fused anchor x GT IoU + first-index argmax + strict fg/bg thresholds for a whole batch.
// Replaces matcher (retinanet/box_utils.py:51-80) + torchvision box_iou (tv:ops/boxes.py:319-371).
//
// Roofline: instruction throughput (ALU), not HBM — traffic is 16 B in + 8(+4) B out per anchor,
// work is A x G IoU pairs.  The [G,A] IoU matrix of the reference is never materialised.
//
// Design
//  * MATCH_K consecutive-by-32 anchors per thread (a warp owns 32*MATCH_K consecutive anchors), one CTA row per
//    image; the image's GT boxes are staged in shared memory (box, area, "malformed" flag) in tiles of GT_TILE;
//  * warp-cooperative culling: each warp reduces the bounding box of its 32 anchors with
//    shuffles, then the 32 lanes test 32 GT boxes at a time against it and a ballot yields the
//    list of GT boxes that can have non-zero intersection with ANY anchor of the warp; only those
//    are evaluated.  A culled pair has clamp(rb-lt,0)=0 in at least one axis, so its IoU is
//    exactly +0 — bit-identical to evaluating it;
//  * size culling: IoU <= min(area)/max(area), so a GT box whose area is below bg_thr*(1-2^-20) times the
//    smallest anchor area of the warp, or above the largest anchor area divided by that factor, cannot
//    reach bg_thr with any anchor of the warp and is skipped like a pruned pair (anchors of one pyramid
//    level only meet GT boxes of comparable size);
//  * the IEEE division is only issued for pairs whose IoU can reach bg_thr: pairs with
//    inter < bg_thr*(1-2^-20)*union are provably below bg_thr after rounding and can neither
//    change the fg/bg/ignore decision nor win the argmax of a foreground anchor;
//  * every surviving pair is evaluated with individually rounded fp32 ops (__fmul_rn/__fadd_rn/
//    __fsub_rn/__fdiv_rn, no FMA contraction) in the reference's operation order, so the result is
//    bit-identical to torch on CPU or CUDA; ties keep the lowest GT index, NaN propagates
//    (torch.max semantics);
//  * culling/pruning rely on 0 < bg_thr < fg_thr and on well-formed boxes (finite, x2>=x1,
//    y2>=y1, anchors with positive area); malformed GT boxes or anchors fall back — per GT box /
//    per warp — to the unculled NaN-propagating path, other thresholds disable both tricks.
#include "match_body.cuh"
#include "pp_internal.cuh"

namespace {

using namespace rnmatch;

#ifndef MATCH_K
#define MATCH_K 3          // anchors per thread (measured: K=1 43.9 us, 2 35.3, 3 31.2, 4 31.5, 6 33.8 at config 2; 78.7, 66.5, 66.4, 72.8, 85.3 us at config 5): lane l of warp w owns anchors base + 32*k + l (consecutive, spatially close)
#endif
constexpr int MATCH_SPAN = MATCH_BLOCK * MATCH_K;   // anchors per CTA

#ifndef MATCH_MINB
#define MATCH_MINB 5    // 48 registers: 30.5 us; unconstrained (64 regs) 31.2 us, 6 CTAs (40 regs, spills) 31.1 us at config 2
#endif
template <bool FAST>
__global__ void __launch_bounds__(MATCH_BLOCK, MATCH_MINB)
match_kernel(const float4 *__restrict__ anchors, long long A, long long anchor_stride,
             const float4 *__restrict__ gt,
             const long long *__restrict__ labels, const int *__restrict__ gt_off, float fg_thr, float bg_thr,
             float prune_c, long long *__restrict__ matches, int *__restrict__ codes, int *__restrict__ fg_count) {
    __shared__ float4 s_box[GT_TILE];
    __shared__ float s_area[GT_TILE];      // NaN marks a malformed GT box (evaluated by the generic path)

    const int n = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long a0 = (long long)blockIdx.x * MATCH_SPAN + (long long)warp * (32 * MATCH_K) + lane;
    const int g0 = gt_off[n];
    const int G = gt_off[n + 1] - g0;

    float4 a[MATCH_K];
    bool live[MATCH_K];
    int m[MATCH_K];
#pragma unroll
    for (int k = 0; k < MATCH_K; ++k) {
        const long long ai = a0 + 32 * k;
        live[k] = ai < A;
        a[k] = make_float4(0.f, 0.f, 1.f, 1.f);
        if (live[k]) a[k] = anchors[(long long)n * anchor_stride + ai];
    }
    match_block<FAST, MATCH_K>(s_box, s_area, a, live, gt + g0, G, fg_thr, bg_thr, prune_c, m);

    int nfg = 0;
#pragma unroll
    for (int k = 0; k < MATCH_K; ++k) {
        const long long ai = a0 + 32 * k;
        if (live[k]) {
            if (matches) matches[(long long)n * A + ai] = (long long)m[k];
            if (codes) codes[(long long)n * A + ai] = pack_code(m[k], labels + g0);
        }
        nfg += __popc(__ballot_sync(0xffffffffu, live[k] && m[k] >= 0));
    }
    if (fg_count && lane == 0 && nfg) atomicAdd(fg_count + n, nfg);
}

}  // namespace

extern "C" int rn_match(const float *anchors, int64_t A, int64_t anchor_image_stride, const float *gt_boxes, const int64_t *gt_labels,
                        const int32_t *gt_off, int N, int64_t gt_total, float fg_thr, float bg_thr, int64_t *matches,
                        int32_t *codes, int32_t *fg_count, rn_stream_t stream) {
    return rnpp::match_impl(anchors, A, anchor_image_stride, gt_boxes, gt_labels, gt_off, N, gt_total, fg_thr, bg_thr, matches,
                            codes, fg_count, stream, true);
}

// zero_fg = false: the caller (rn_train_detect's prep kernel) already zeroed fg_count on the stream
int rnpp::match_impl(const float *anchors, int64_t A, int64_t anchor_image_stride, const float *gt_boxes, const int64_t *gt_labels,
                     const int32_t *gt_off, int N, int64_t gt_total, float fg_thr, float bg_thr, int64_t *matches,
                     int32_t *codes, int32_t *fg_count, rn_stream_t stream, bool zero_fg) {
    RN_CHECK_ARG(anchors && gt_off, RN_E_BADARG, "rn_match: null anchors/gt_off");
    RN_CHECK_ARG(A >= 0 && N >= 0, RN_E_BADARG, "rn_match: negative size");
    RN_CHECK_ARG(anchor_image_stride == 0 || anchor_image_stride >= A, RN_E_BADARG, "rn_match: bad anchor_image_stride");
    RN_CHECK_ARG(fg_thr > bg_thr, RN_E_BADARG, "rn_match: match_thr (%g) must exceed back_thr (%g) (box_utils.py:66)",
                 (double)fg_thr, (double)bg_thr);
    RN_CHECK_ARG(!codes || gt_labels, RN_E_BADARG, "rn_match: codes requested without gt_labels");
    RN_CHECK_ARG(gt_total >= 0, RN_E_BADARG, "rn_match: negative gt_total");
    // the packed code keeps the GT index in 20 bits: every G_n <= sumG must fit (silent corruption otherwise)
    RN_CHECK_ARG(!codes || gt_total < (1LL << 20), RN_E_TOOLARGE,
                 "rn_match: %lld GT boxes in the batch; the packed codes need fewer than 2^20", (long long)gt_total);
    RN_CHECK_ARG(N <= 65535, RN_E_TOOLARGE, "rn_match: N=%d exceeds 65535 images per call", N);
    if (A == 0 || N == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    if (fg_count && zero_fg) {   // the per-image counters are accumulated with integer atomics: start from zero
        cudaError_t e = cudaMemsetAsync(fg_count, 0, (size_t)N * sizeof(int32_t), s);
        if (e != cudaSuccess) { rn_set_error("rn_match: memset failed: %s", cudaGetErrorString(e)); return (int)e; }
    }
    dim3 grid((unsigned)((A + MATCH_SPAN - 1) / MATCH_SPAN), (unsigned)N);
    const bool fast = bg_thr > 0.0f;  // then fg_thr > bg_thr > 0: culling and pruning are exact
    if (fast) {
        // inter < uni*bg*(1-2^-20)  =>  fl(inter/uni) < bg   (rounding slack is 2^-23 per op)
        float prune_c = bg_thr * (1.0f - 9.5367431640625e-07f);
        match_kernel<true><<<grid, MATCH_BLOCK, 0, s>>>((const float4 *)anchors, A, anchor_image_stride,
                                                        (const float4 *)gt_boxes,
                                                        (const long long *)gt_labels, gt_off, fg_thr, bg_thr, prune_c,
                                                        (long long *)matches, codes, fg_count);
    } else {
        match_kernel<false><<<grid, MATCH_BLOCK, 0, s>>>((const float4 *)anchors, A, anchor_image_stride,
                                                        (const float4 *)gt_boxes,
                                                         (const long long *)gt_labels, gt_off, fg_thr, bg_thr, 0.0f,
                                                         (long long *)matches, codes, fg_count);
    }
    RN_CHECK_LAUNCH("rn_match");
    return 0;
}
