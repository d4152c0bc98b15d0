// Fused anchor x GT IoU + first-index argmax + strict fg/bg thresholds for a whole batch.
// Replaces matcher (retinanet/box_utils.py:51-80) + torchvision box_iou (tv:ops/boxes.py:319-371).
//
// Roofline: instruction throughput (ALU), not HBM — traffic is 16 B in + 8(+4) B out per anchor,
// work is A x G IoU pairs.  The [G,A] IoU matrix of the reference is never materialised.
//
// Design
//  * one thread per anchor, one CTA row per image; the image's GT boxes are staged in shared
//    memory (box, area, "malformed" flag) in tiles of GT_TILE;
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

namespace {

using namespace rnmatch;

template <bool FAST>
__global__ void __launch_bounds__(MATCH_BLOCK)
match_kernel(const float4 *__restrict__ anchors, long long A, long long anchor_stride,
             const float4 *__restrict__ gt,
             const long long *__restrict__ labels, const int *__restrict__ gt_off, float fg_thr, float bg_thr,
             float prune_c, long long *__restrict__ matches, int *__restrict__ codes, int *__restrict__ fg_count) {
    __shared__ float4 s_box[GT_TILE];
    __shared__ float s_area[GT_TILE];      // NaN marks a malformed GT box (evaluated by the generic path)

    const int n = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const long long ai = (long long)blockIdx.x * MATCH_BLOCK + threadIdx.x;
    const bool live = ai < A;
    const int g0 = gt_off[n];
    const int G = gt_off[n + 1] - g0;

    float4 a = make_float4(0.f, 0.f, 1.f, 1.f);
    if (live) a = anchors[(long long)n * anchor_stride + ai];
    const int m = match_block<FAST>(s_box, s_area, a, live, gt + g0, G, fg_thr, bg_thr, prune_c);

    if (live) {
        if (matches) matches[(long long)n * A + ai] = (long long)m;
        if (codes) codes[(long long)n * A + ai] = pack_code(m, labels + g0);
    }
    if (fg_count) {
        unsigned fgm = __ballot_sync(0xffffffffu, live && m >= 0);
        if (lane == 0 && fgm) atomicAdd(fg_count + n, __popc(fgm));
    }
}

}  // namespace

extern "C" int rn_match(const float *anchors, int64_t A, int64_t anchor_image_stride, const float *gt_boxes, const int64_t *gt_labels,
                        const int32_t *gt_off, int N, float fg_thr, float bg_thr, int64_t *matches, int32_t *codes,
                        int32_t *fg_count, rn_stream_t stream) {
    RN_CHECK_ARG(anchors && gt_off, RN_E_BADARG, "rn_match: null anchors/gt_off");
    RN_CHECK_ARG(A >= 0 && N >= 0, RN_E_BADARG, "rn_match: negative size");
    RN_CHECK_ARG(anchor_image_stride == 0 || anchor_image_stride >= A, RN_E_BADARG, "rn_match: bad anchor_image_stride");
    RN_CHECK_ARG(fg_thr > bg_thr, RN_E_BADARG, "rn_match: match_thr (%g) must exceed back_thr (%g) (box_utils.py:66)",
                 (double)fg_thr, (double)bg_thr);
    RN_CHECK_ARG(!codes || gt_labels, RN_E_BADARG, "rn_match: codes requested without gt_labels");
    RN_CHECK_ARG(N <= 65535, RN_E_TOOLARGE, "rn_match: N=%d exceeds 65535 images per call", N);
    if (A == 0 || N == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    if (fg_count) {   // the per-image counters are accumulated with integer atomics: start from zero
        cudaError_t e = cudaMemsetAsync(fg_count, 0, (size_t)N * sizeof(int32_t), s);
        if (e != cudaSuccess) { rn_set_error("rn_match: memset failed: %s", cudaGetErrorString(e)); return (int)e; }
    }
    dim3 grid((unsigned)((A + MATCH_BLOCK - 1) / MATCH_BLOCK), (unsigned)N);
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
