// Internal interface between postprocess.cu and loss.cu for rn_train_detect (the loss kernel filters the scores while it
// streams the logits; the lazy NMS of postprocess.cu then runs on the candidate lists it filled).
#pragma once
#include "rn_common.cuh"

namespace rnpp {

// Where the in-loss score filter appends its candidates: per-image lists of 8-byte keys
// (~score_bits << 32) | (class * A + anchor), exactly what score_filter_kernel<.., LAZY> writes.
struct LazySink {
    unsigned *img_count;             // [N] candidates found per image (may exceed cap_n)
    unsigned long long *pool_key;    // [N][cap_n]
    unsigned cap_n;
    float x_lo, thr;                 // guard logit (sigmoid(x) <= thr for every x <= x_lo) and the score threshold
};

// zero = false: the caller zeroes img_count and out_status itself (rn_train_detect's prep kernel)
int lazy_begin(int N, int64_t A, int C, float score_thr, int max_det, int pre_nms_topk, const int64_t *level_off_host,
               int num_levels, int64_t cand_capacity, int32_t *out_status, void *workspace, size_t workspace_bytes,
               cudaStream_t s, LazySink *sink, bool zero);
int lazy_end(const float *bbox, const float *anchors, int64_t anchor_image_stride, const int32_t *im_hw, int N, int64_t A,
             int C, float score_thr, double nms_thr, int max_det, const float *weights_host, int pre_nms_topk,
             const int64_t *level_off_host, int num_levels, int64_t cand_capacity, float *out_boxes, float *out_scores,
             int64_t *out_labels, int32_t *out_count, int32_t *out_status, void *workspace, cudaStream_t s,
             const float *out_ratio_hw, int out_format);

// rn_match with the zeroing of fg_count optional (match.cu)
int match_impl(const float *anchors, int64_t A, int64_t anchor_image_stride, const float *gt_boxes, const int64_t *gt_labels,
               const int32_t *gt_off, int N, int64_t gt_total, float fg_thr, float bg_thr, int64_t *matches, int32_t *codes,
               int32_t *fg_count, rn_stream_t stream, bool zero_fg);

}  // namespace rnpp
