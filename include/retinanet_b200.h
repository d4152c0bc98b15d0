/* retinanet_b200 — C ABI of the B200-native dense per-anchor path of RetinaNet.
 *
 * The reference (benihime91/pytorch_retinanet) has NO FFI/plugin interface for this path: the
 * boundary is plain Python method dispatch inside `Retinanet` (retinanet/models.py:266,270,284,287).
 * This header is therefore the interface a native binding WOULD use; every entry point names the
 * reference function (file:line, relative to the reference root) it replaces.  The Python host
 * side in pytorch_retinanet_b200/ binds it with ctypes.
 *
 * Conventions
 *  - All pointers are DEVICE pointers on the current CUDA device unless the name ends in `_host`.
 *  - All tensors are contiguous, row-major; floats are fp32, indices int64 unless stated.
 *  - Memory is owned by the caller (torch allocates it); the library allocates nothing persistent
 *    and never frees caller memory.  Work is enqueued on `stream` (a cudaStream_t passed as void*);
 *    no entry point synchronises the host.
 *  - Return value: 0 on success; >0 = cudaError_t; <0 = argument error (RN_E_*).
 *    rn_last_error() returns a thread-local description.  Nothing throws across the ABI.
 *  - Anchors are normally shared by the batch ([A,4], anchor_image_stride = 0); per-image anchor
 *    sets ([N,A,4]) are supported with anchor_image_stride = A (in anchors, not bytes).
 *  - Ragged ground truth is packed: gt_boxes [sumG,4], gt_labels [sumG] (1-based class ids, as in
 *    the reference, README.md:132) and gt_off [N+1] int32 prefix offsets.
 */
#ifndef RETINANET_B200_H_
#define RETINANET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RN_ABI_VERSION 2

#define RN_E_BADARG   (-1)  /* null pointer / negative size / unsupported combination          */
#define RN_E_TOOLARGE (-2)  /* size exceeds a documented limit (levels, cell anchors, classes) */
#define RN_E_WORKSPACE (-3) /* workspace too small                                             */

#define RN_MAX_LEVELS 8

/* rn_postprocess algorithms (same results; see the function's comment) */
#define RN_PP_LAZY    0
#define RN_PP_GENERAL 1

typedef void *rn_stream_t; /* cudaStream_t */

/* ---- image-sharded multi-GPU exchange ----------------------------------------------
 * The reference computes the loss per image and averages over the batch (retinanet/losses.py:126-140); with the
 * batch sharded by image over the GPUs of one NVLink/NVSwitch box the ONLY exchange of the path is the sum of the
 * 16-byte vector out_total = {cls, reg, sum F, N_local} over the ranks.  It is folded into the loss's final
 * reduction kernel: the block that finishes the local sum stores its four values, each paired with a step
 * sequence number in one 8-byte word, straight into every peer's receive slots over NVLink (peer-mapped memory),
 * then reads the slots the peers filled for it and adds them in rank order — every rank ends with the same bits,
 * no separate collective launch, no host involvement, CUDA-graph capturable.
 * A receive buffer (rn_comm_bytes() bytes, zero-initialised, allocated with rn_comm_alloc so that it can be
 * exported) lives on every rank; peers[r] is rank r's buffer as mapped into THIS process (peers[rank] = the local
 * one).  Every rank must run the same sequence of exchanging calls (as with any collective).  A rank that does not
 * hear from a peer within ~5 s writes NaN into out_total and sets the error word (rn_comm_error) instead of hanging. */
#define RN_MAX_PEERS 16
typedef struct {
    void *peers[RN_MAX_PEERS]; /* device pointers: receive buffer of every rank, mapped into this process */
    int32_t rank, world;
} rn_exchange_t;
size_t rn_comm_bytes(void);
int rn_comm_alloc(void **out_ptr);                                 /* cudaMalloc + zero (current device)        */
int rn_comm_free(void *ptr);
int rn_comm_export(void *ptr, void *handle_out_host /*64 bytes*/); /* cudaIpcGetMemHandle                         */
int rn_comm_import(const void *handle_host /*64 bytes*/, void **out_ptr); /* cudaIpcOpenMemHandle (peer access)   */
int rn_comm_unmap(void *imported_ptr);                             /* cudaIpcCloseMemHandle                       */
int rn_comm_error(const void *local_buf, int32_t *out_host);       /* synchronous read of the error word (tests) */
/* The exchange by itself: total[4] <- sum over the ranks (one 32-thread launch).  For a rank whose shard is empty
 * (it launches no loss kernel but must take part) and for tests; rn_loss / rn_train_loss / rn_loss_levels do the
 * same inside their final reduction when given `exchange_host`.                                               */
int rn_exchange_total(float *total /*[4] in/out*/, const rn_exchange_t *exchange_host, rn_stream_t stream);

int rn_abi_version(void);
const char *rn_last_error(void);
/* Number of kernel launches this library has issued in the process so far (launches recorded into a CUDA graph
 * during stream capture count once, when recorded).  Instrumentation for bench.py's `gpu_launches`.            */
uint64_t rn_launch_count(void);

/* ---- anchors ---------------------------------------------------------------------------------
 * Replaces AnchorGenerator.grid_anchors / _compute_grid_offsets (retinanet/anchors.py:151-197) for
 * all pyramid levels in one launch.  `cells` is the concatenation of the per-level cell-anchor
 * tables ([sum_l na_l, 4], the BufferList of anchors.py:97-108, computed on the host in double as
 * anchors.py:110-135).  level_desc_host[l] = {H_l, W_l, stride_l, na_l}.  Output anchor index inside
 * a level is (y*W + x)*na + a, levels concatenated in order (anchors.py:228).                     */
int rn_anchor_grid(const float *cells, const int32_t *level_desc_host /*[L][4]*/, int num_levels,
                   double offset, float *out_anchors /*[A,4]*/, int64_t num_anchors, rn_stream_t stream);

/* ---- target packing -------------------------------------------------------
 * Packs the reference's per-image target lists (targets[i]["boxes"] [G_i,4] fp32, targets[i]["labels"] [G_i]
 * int64; retinanet/losses.py:126-128) into gt_boxes [sumG,4], gt_labels [sumG], gt_off [N+1] in ONE launch:
 * boxes_host / labels_host are HOST arrays of N DEVICE pointers, counts_host the N box counts.  No host->device
 * copy is issued (pointers and counts travel in the kernel parameters).  ratios_hw_host (optional, [N][2] =
 * ratio_h, ratio_w) additionally applies torchvision's resize_boxes to the boxes while packing (what
 * GeneralizedRCNNTransform.forward does to the targets, retinanet/models.py:279).  labels may be NULL.   */
int rn_pack_targets(const float *const *boxes_host, const int64_t *const *labels_host, const int32_t *counts_host, int N,
                    const float *ratios_hw_host /*[N][2] or NULL*/, float *out_boxes, int64_t *out_labels /*or NULL*/,
                    int32_t *out_off /*[N+1]*/, rn_stream_t stream);

/* ---- matcher ----------------------------------------------------------------------------------
 * Replaces matcher (retinanet/box_utils.py:51-80) + torchvision box_iou for a whole batch:
 * fused IoU + first-index argmax + strict 0.5/0.4 thresholds.  matches[n,a] = -2 ignore,
 * -1 background, g >= 0 index of the matched GT box inside image n.  Images with zero GT boxes
 * get -2 everywhere (box_utils.py:70-71).
 * Optional outputs (may be NULL):
 *   codes   [N,A] int32 — packed per-anchor target used by rn_loss: -2 / -1 / (g | cls << 20), cls = label-1 for
 *                         labels 1..2047 and 2047 ("no class column") otherwise — a label 0 keeps the regression term
 *                         and gets all-zero class targets like the reference (losses.py:96-103); a label > C, for
 *                         which the reference's one_hot raises, trains as background here.  Requires gt_labels,
 *                         sumG < 2^20 and C <= 2047.
 *   fg_count[N]   int32 — number of foreground anchors per image (zeroed by the call itself).       */
int rn_match(const float *anchors /*[A,4] or [N,A,4]*/, int64_t A, int64_t anchor_image_stride,
             const float *gt_boxes /*[sumG,4]*/,
             const int64_t *gt_labels /*[sumG] or NULL*/, const int32_t *gt_off /*[N+1]*/, int N,
             int64_t gt_total /* sumG; codes need every G_n <= sumG < 2^20, else RN_E_TOOLARGE */,
             float fg_thr, float bg_thr, int64_t *matches /*[N,A] or NULL*/, int32_t *codes /*[N,A] or NULL*/,
             int32_t *fg_count /*[N] or NULL*/, rn_stream_t stream);

/* ---- box coding -------------------------------------------------------------------------------
 * rn_encode replaces bbox_2_activ (box_utils.py:25-34); rn_decode replaces activ_2_bbox
 * (box_utils.py:37-48) INCLUDING its quirk (sizes = a_wh * exp(dx,dy)).  weights_host = 4 floats. */
int rn_encode(const float *boxes, const float *anchors, int64_t n, const float *weights_host,
              float *out, rn_stream_t stream);
int rn_decode(const float *activations, const float *anchors, int64_t n, const float *weights_host,
              float *out, rn_stream_t stream);

/* ---- losses -----------------------------------------------------------------------------------
 * Replaces RetinaNetLosses.calc_loss/forward (retinanet/losses.py:49-145) given the packed codes
 * written by rn_match: sigmoid focal loss on (logits + 1) with alpha applied inverted and weights
 * from detached probabilities (losses.py:29-47,84), smooth-L1 on encoded targets (losses.py:19-27),
 * both divided by max(1, F_n) per image, then averaged over `batch_div` images (losses.py:108-109,
 * 138-140).
 *   out_image [N,3]  = {cls_n / max(1,F_n), reg_n / max(1,F_n), F_n}   (fp32)
 *   out_total [4]    = {sum_n cls_n/max(1,F_n) / batch_div, sum_n reg_n/max(1,F_n) / batch_div, sum_n F_n, N}
 *                      (the 16-byte vector that image-sharded multi-GPU runs all-reduce)
 *   grad_logits / grad_bbox (optional, both or neither): d out_total[0] / d logits and
 *   d out_total[1] / d bbox_preds, written for EVERY element (zeros where no gradient flows).
 * Reductions are two-stage and order-fixed, so results are bit-reproducible run to run.
 * workspace: rn_loss_workspace_bytes(N, A, C).                                                   */
size_t rn_loss_workspace_bytes(int N, int64_t A, int C);
int rn_loss(const float *logits /*[N,A,C]*/, const float *bbox /*[N,A,4]*/, const float *anchors,
            int64_t anchor_image_stride, const float *gt_boxes /*[sumG,4]*/, const int32_t *gt_off /*[N+1]*/, const int32_t *codes /*[N,A]*/,
            const int32_t *fg_count /*[N]*/, int N, int64_t A, int C, float alpha, float gamma, float beta,
            const float *weights_host /*[4]*/, float batch_div, float *out_image /*[N,3]*/,
            float *out_total /*[4]*/, float *grad_logits /*[N,A,C] or NULL*/, float *grad_bbox /*[N,A,4] or NULL*/,
            void *workspace, size_t workspace_bytes, rn_stream_t stream,
            const rn_exchange_t *exchange_host /* NULL = single GPU; else out_total is summed over the ranks */);

/* The training half of the path behind ONE call: rn_match (codes + foreground counts) followed by rn_loss and its
 * final reduction on the same stream — same arguments, same outputs (retinanet/losses.py:113-145 with
 * box_utils.py:51-80 inside).  `codes` [N,A] and `fg_count` [N] are outputs here (scratch for the caller, same
 * contents as rn_match writes).  workspace: rn_train_loss_workspace_bytes(N, A, C).
 * (Two single-launch fusions of the two kernels were measured slower than the sequence on B200 and are not shipped:
 * see DESIGN.md 4.3b.)                                                                                     */
size_t rn_train_loss_workspace_bytes(int N, int64_t A, int C);
int rn_train_loss(const float *logits /*[N,A,C]*/, const float *bbox /*[N,A,4]*/, const float *anchors,
                  int64_t anchor_image_stride, const float *gt_boxes /*[sumG,4]*/, const int64_t *gt_labels /*[sumG]*/,
                  const int32_t *gt_off /*[N+1]*/, int N, int64_t gt_total, int64_t A, int C, float fg_thr, float bg_thr, float alpha,
                  float gamma, float beta, const float *weights_host /*[4]*/, float batch_div, int32_t *codes /*[N,A]*/,
                  int32_t *fg_count /*[N]*/, float *out_image /*[N,3]*/, float *out_total /*[4]*/,
                  float *grad_logits /*[N,A,C] or NULL*/, float *grad_bbox /*[N,A,4] or NULL*/, void *workspace,
                  size_t workspace_bytes, rn_stream_t stream, const rn_exchange_t *exchange_host /*or NULL*/);

/* Both halves of the path on the same head outputs with ONE pass over the logits: rn_train_loss whose loss kernel is
 * also the score filter of rn_postprocess(RN_PP_LAZY) (retinanet/losses.py:113-145 + retinanet/models.py:160-243 on one
 * batch, e.g. detections during training).  Arguments and outputs are those of the two calls; results are identical to
 * calling them one after the other.  Needs C % 4 == 0, 16-byte aligned logits, A*C < 2^32 and the default math mode.
 * out_status as in rn_postprocess: a candidate-pool overflow or the fallback flag is resolved by calling
 * rn_postprocess on the same inputs.  workspace: rn_train_detect_workspace_bytes(N, A, C, cand_capacity, max_det).
 * phases: mask of 1 (zeroing + matcher), 4 (loss + score filter + final reduction) and 2 (the NMS on the candidate lists
 * the loss phase filled); 7 = the whole step.  Separate calls with the same arguments let a caller put the phases on
 * different streams: the ALU-bound matcher of the next batch and the latency-bound NMS of the previous one then run
 * under the HBM-bound loss kernel of the current one (graphs.HotPathPipeline).                                      */
size_t rn_train_detect_workspace_bytes(int N, int64_t A, int C, int64_t cand_capacity, int max_det);
int rn_train_detect(const float *logits, const float *bbox, const float *anchors, int64_t anchor_image_stride,
                    const float *gt_boxes, const int64_t *gt_labels, const int32_t *gt_off, int N, int64_t gt_total,
                    int64_t A, int C, float fg_thr, float bg_thr, float alpha, float gamma, float beta,
                    const float *weights_host /*[4]*/, float batch_div, int32_t *codes, int32_t *fg_count,
                    float *out_image /*[N,3]*/, float *out_total /*[4]*/, float *grad_logits /*or NULL*/,
                    float *grad_bbox /*or NULL*/, const int32_t *im_hw /*[N,2]*/, float score_thr, double nms_thr, int max_det,
                    int pre_nms_topk, const int64_t *level_off_host /*[L+1] or NULL*/, int num_levels, int64_t cand_capacity,
                    float *out_boxes, float *out_scores, int64_t *out_labels, int32_t *out_count, int32_t *out_status,
                    const float *out_ratio_hw /*[N,2] or NULL*/, int out_format, void *workspace, size_t workspace_bytes,
                    rn_stream_t stream, const rn_exchange_t *exchange_host /*or NULL*/, int phases);

/* Dense element-wise losses, API parity with RetinaNetLosses.focal_loss (losses.py:29-47, arbitrary
 * float targets of the logits' shape, NO +1 shift) and RetinaNetLosses.smooth_l1_loss (losses.py:19-27).
 * out_sum [1] = sum over the n elements; grad (optional, [n]) = d out_sum / d input.
 * workspace: rn_dense_loss_workspace_bytes().                                                      */
size_t rn_dense_loss_workspace_bytes(void);
int rn_focal_loss_dense(const float *logits, const float *targets, int64_t n, float alpha, float gamma,
                        float *out_sum, float *grad /*[n] or NULL*/, void *workspace, size_t workspace_bytes,
                        rn_stream_t stream);
int rn_smooth_l1_dense(const float *input, const float *target, int64_t n, float beta, float *out_sum,
                       float *grad /*[n] or NULL*/, void *workspace, size_t workspace_bytes, rn_stream_t stream);

/* In-place scale of a gradient buffer by a DEVICE scalar (autograd's grad_output); every block
 * returns immediately when *scale == 1.0f, so the common case costs one launch and no traffic.    */
int rn_scale_by_device_scalar(float *buf, int64_t n, const float *scale, rn_stream_t stream);

/* ---- inference post-processing ----------------------------------------------------------------
 * Replaces Retinanet.process_detections (retinanet/models.py:160-243) for a whole batch:
 * sigmoid + strict score threshold, decode (with the reference quirk) + clip to im_hw, small-box
 * filter (w,h >= 0.01), per-(image,class) greedy NMS (strict IoU > nms_thr, stable score order,
 * ties -> lower anchor index), concatenation, labels+1, top `max_det` by (score desc, class asc,
 * anchor asc).  pre_nms_topk > 0 additionally keeps only the top-k scores per (image, pyramid
 * level) before NMS (an extension; 0 = exact reference behaviour); level_off_host [L+1] gives the
 * anchor offsets of the levels and is only read when pre_nms_topk > 0.
 *   out_boxes [N,max_det,4] fp32, out_scores [N,max_det] fp32, out_labels [N,max_det] int64,
 *   out_count [N] int32 (number of valid rows per image).
 *   out_status [4] int32: {candidates found, candidate capacity, fallback flag, 0}.  found > capacity
 *   means the workspace was too small and the call must be repeated with cand_capacity >= found.
 *   algo = RN_PP_LAZY (per-image greedy NMS in global score order with early exit; bounded work) sets
 *   the fallback flag when an image could not be completed within its round budget — the caller
 *   then repeats the call with algo = RN_PP_GENERAL (per-(image,class) segments, any size).  Both
 *   algorithms produce identical results.
 * Output epilogue options (both off = reference output of process_detections):
 *   out_ratio_hw [N,2] (ratio_h, ratio_w per image): the final boxes are multiplied like torchvision's
 *     resize_boxes does in transform.postprocess right after the path (retinanet/models.py:271);
 *   out_format 1: boxes are written as COCO xywh (utils/coco/coco_eval.py:159-161), after the resize.
 * workspace: rn_postprocess_workspace_bytes(N, A, C, cand_capacity, max_det).                      */
size_t rn_postprocess_workspace_bytes(int N, int64_t A, int C, int64_t cand_capacity, int max_det);
int rn_postprocess(const float *logits /*[N,A,C]*/, const float *bbox /*[N,A,4]*/, const float *anchors,
                   int64_t anchor_image_stride, const int32_t *im_hw /*[N,2] (h,w)*/, int N, int64_t A, int C, float score_thr, double nms_thr,
                   int max_det, const float *weights_host /*[4]*/, int pre_nms_topk,
                   const int64_t *level_off_host /*[L+1] or NULL*/, int num_levels, int algo, int64_t cand_capacity,
                   float *out_boxes, float *out_scores, int64_t *out_labels, int32_t *out_count,
                   int32_t *out_status, void *workspace, size_t workspace_bytes, rn_stream_t stream,
                   const float *out_ratio_hw /*[N,2] device or NULL*/, int out_format /*0 xyxy, 1 xywh*/);

/* ---- per-level NCHW entry points -------------------------------------------
 * Same semantics and outputs as rn_loss / rn_postprocess, but the class and box activations are the
 * RAW per-level conv outputs of the reference's head, before its view/permute/contiguous/cat
 * (retinanet/layers.py:189-195, 253-259): cls level l is [N, na_l*C, H_l, W_l], box level l is
 * [N, na_l*4, H_l, W_l] (channel = a*C + c resp. a*4 + k), contiguous.  level_desc_host[l] =
 * {H_l, W_l, na_l}; *_levels_host are HOST arrays of num_levels DEVICE pointers.  Gradients (optional)
 * are written in the same per-level layout.  The anchor of element (a, y, x) of level l is
 * off_l + (y*W_l + x)*na_l + a, i.e. exactly the reference's order, so rn_match's codes apply unchanged. */
size_t rn_loss_levels_workspace_bytes(int N, const int32_t *level_desc_host, int num_levels);
int rn_loss_levels(const float *const *cls_levels_host, const float *const *bbox_levels_host,
                   const int32_t *level_desc_host /*[L][3]*/, int num_levels, const float *anchors,
                   int64_t anchor_image_stride, const float *gt_boxes, const int32_t *gt_off, const int32_t *codes,
                   const int32_t *fg_count, int N, int64_t A, int C, float alpha, float gamma, float beta,
                   const float *weights_host, float batch_div, float *out_image /*[N,3]*/, float *out_total /*[4]*/,
                   float *const *grad_cls_levels_host /*or NULL*/, float *const *grad_bbox_levels_host /*or NULL*/,
                   void *workspace, size_t workspace_bytes, rn_stream_t stream,
                   const rn_exchange_t *exchange_host /*or NULL*/);
size_t rn_postprocess_levels_workspace_bytes(int N, int64_t A, int C, int64_t cand_capacity, int max_det);
int rn_postprocess_levels(const float *const *cls_levels_host, const float *const *bbox_levels_host,
                          const int32_t *level_desc_host /*[L][3]*/, int num_levels, const float *anchors,
                          int64_t anchor_image_stride, const int32_t *im_hw, int N, int64_t A, int C, float score_thr,
                          double nms_thr, int max_det, const float *weights_host, int pre_nms_topk, int algo,
                          int64_t cand_capacity, float *out_boxes, float *out_scores, int64_t *out_labels,
                          int32_t *out_count, int32_t *out_status, void *workspace, size_t workspace_bytes,
                          rn_stream_t stream, const float *out_ratio_hw /*[N,2] or NULL*/, int out_format);

/* Stand-alone batched greedy NMS (torchvision `nms` semantics, tv:ops/boxes.py:20-48) over
 * segments whose boxes are ALREADY sorted by score descending (ties: original order): boxes [K,4],
 * seg_off [S+1] int32, keep_flags [K] uint8 out (1 = kept).  workspace >= 16*K bytes.            */
int rn_nms_segments(const float *boxes, const int32_t *seg_off, int num_segments, int64_t total_boxes,
                    double nms_thr, uint8_t *keep_flags, void *workspace, size_t workspace_bytes,
                    rn_stream_t stream);

/* Test hook: 0 = fast SFU math in rn_loss (default), 1 = libdevice-precise math.  Returns the
 * previous mode.  Process-global.                                                                */
int rn_loss_set_math_mode(int mode);

#ifdef __cplusplus
}
#endif
#endif /* RETINANET_B200_H_ */
