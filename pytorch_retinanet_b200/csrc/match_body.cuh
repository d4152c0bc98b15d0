// Device-side body of the fused anchor x GT matcher (rn_match, match.cu), split into reusable pieces: warp-level
// culling set-up, GT tile staging, one warp x one staged tile, the final decision.  See match.cu for the design.
#pragma once
#include "rn_common.cuh"

namespace rnmatch {

constexpr int GT_TILE = 512;
constexpr int MATCH_BLOCK = 256;

// torch.max / torch.min / clamp(min=0) propagate NaN; fmaxf/fminf do not.
__device__ __forceinline__ float nan_max(float a, float b) { return (a != a) ? a : ((b != b) ? b : fmaxf(a, b)); }
__device__ __forceinline__ float nan_min(float a, float b) { return (a != a) ? a : ((b != b) ? b : fminf(a, b)); }

__device__ __forceinline__ float box_area(const float4 b) {
    return __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
}

// Full reference arithmetic, NaN-propagating (slow path for malformed boxes / exotic thresholds).
__device__ __forceinline__ float iou_generic(const float4 g, float ag, const float4 a, float aa) {
    float w = __fsub_rn(nan_min(g.z, a.z), nan_max(g.x, a.x));
    float h = __fsub_rn(nan_min(g.w, a.w), nan_max(g.y, a.y));
    w = (w != w) ? w : fmaxf(w, 0.0f);
    h = (h != h) ? h : fmaxf(h, 0.0f);
    float inter = __fmul_rn(w, h);
    float uni = __fsub_rn(__fadd_rn(ag, aa), inter);
    return __fdiv_rn(inter, uni);
}

__device__ __forceinline__ bool box_well_formed(const float4 b) {
    // finite coordinates and non-negative extent (NaN fails every comparison)
    return (b.z >= b.x) && (b.w >= b.y) && (fabsf(b.x) <= 3.0e38f) && (fabsf(b.y) <= 3.0e38f) &&
           (fabsf(b.z) <= 3.0e38f) && (fabsf(b.w) <= 3.0e38f);
}

// Per-warp matching state of one group of 32*K anchors (K per lane).
struct WarpCull {
    bool fast;                       // warp-uniform: culling / pruning / non-NaN fast path allowed
    float bx1, by1, bx2, by2;        // bounding box of the warp's anchors
    float ag_lo, ag_hi;              // GT areas outside this range cannot reach bg_thr with any anchor of the warp
};

template <bool FAST, int K>
__device__ __forceinline__ WarpCull warp_cull_setup(const float4 (&a)[K], const float (&aa)[K], const bool (&live)[K],
                                                    const float prune_c) {
    WarpCull w;
    w.fast = FAST;
    w.bx1 = w.by1 = w.bx2 = w.by2 = 0.f;
    w.ag_lo = 0.f;
    w.ag_hi = INFINITY;
    if (FAST) {
        // positive finite extents imply finite, ordered coordinates differences; NaN fails every test
        bool ok = true;
        float x1 = INFINITY, y1 = INFINITY, x2 = -INFINITY, y2 = -INFINITY, amin = INFINITY, amax = 0.0f;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            ok = ok && (!live[k] || ((a[k].z - a[k].x) > 0.0f && (a[k].w - a[k].y) > 0.0f && aa[k] <= 3.0e38f &&
                                     fabsf(a[k].x) <= 3.0e38f && fabsf(a[k].y) <= 3.0e38f));
            if (live[k]) {
                x1 = fminf(x1, a[k].x); y1 = fminf(y1, a[k].y);
                x2 = fmaxf(x2, a[k].z); y2 = fmaxf(y2, a[k].w);
                amin = fminf(amin, aa[k]); amax = fmaxf(amax, aa[k]);
            }
        }
        w.fast = __all_sync(0xffffffffu, ok);
        w.bx1 = rn::warp_min(x1);
        w.by1 = rn::warp_min(y1);
        w.bx2 = rn::warp_max(x2);
        w.by2 = rn::warp_max(y2);
        amin = rn::warp_min(amin);
        amax = rn::warp_max(amax);
        w.ag_lo = amin * prune_c * 0.999f;               // GT areas outside [ag_lo, ag_hi] give IoU < bg_thr with
        w.ag_hi = amax / (prune_c * 0.999f);             // every anchor of this warp (IoU <= area ratio)
    }
    return w;
}

// Stage `tn` GT boxes (box, area; NaN area marks a malformed box) into shared memory; `nthreads` threads, index `tid`.
__device__ __forceinline__ void stage_gt_tile(float4 *s_box, float *s_area, const float4 *__restrict__ gt, const int tn,
                                              const int tid, const int nthreads) {
    for (int j = tid; j < tn; j += nthreads) {
        float4 g = gt[j];
        float ag = box_area(g);
        s_box[j] = g;
        s_area[j] = (box_well_formed(g) && ag <= 3.0e38f) ? ag : __int_as_float(0x7fc00000);
    }
}

// One warp x one staged tile of `tn` GT boxes (global indices t0..t0+tn): cull against the warp's bounding box and
// area range, then evaluate the surviving GT boxes against the K anchors of every lane.
template <int K>
__device__ __forceinline__ void match_tile(const float4 *s_box, const float *s_area, const int tn, const int t0,
                                           const float4 (&a)[K], const float (&aa)[K], const WarpCull &w,
                                           const float prune_c, float (&best)[K], int (&bi)[K]) {
    const int lane = threadIdx.x & 31;
    for (int base = 0; base < tn; base += 32) {
        unsigned mask;
        {
            const int j = base + lane;
            bool hit = j < tn;
            if (hit && w.fast) {
                float4 g = s_box[j];
                const float ag = s_area[j];
                float ww = __fsub_rn(fminf(g.z, w.bx2), fmaxf(g.x, w.bx1));
                float hh = __fsub_rn(fminf(g.w, w.by2), fmaxf(g.y, w.by1));
                hit = (ag != ag) || (ww > 0.0f && hh > 0.0f && ag >= w.ag_lo && ag <= w.ag_hi);
            }
            mask = __ballot_sync(0xffffffffu, hit);
        }
        while (mask) {
            const int j = base + __ffs(mask) - 1;
            mask &= mask - 1;
            const float4 g = s_box[j];
            const float ag = s_area[j];
            const int gi = t0 + j;
            if (w.fast && ag == ag) {
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    // well-formed pair: no NaN possible, IoU is +0 unless both extents are positive
                    float ww = __fsub_rn(fminf(g.z, a[k].z), fmaxf(g.x, a[k].x));
                    float hh = __fsub_rn(fminf(g.w, a[k].w), fmaxf(g.y, a[k].y));
                    if (ww > 0.0f && hh > 0.0f) {
                        float inter = __fmul_rn(ww, hh);
                        float uni = __fsub_rn(__fadd_rn(ag, aa[k]), inter);
                        if (inter >= __fmul_rn(uni, prune_c)) {   // may reach bg_thr: exact quotient
                            float v = __fdiv_rn(inter, uni);
                            if (v > best[k]) { best[k] = v; bi[k] = gi; }
                        }
                    }
                }
            } else {
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    float v = iou_generic(g, box_area(g), a[k], aa[k]);   // s_area holds the NaN marker for malformed boxes
                    if (best[k] == best[k]) {                             // NaN, once taken, stays
                        if (v != v || v > best[k]) { best[k] = v; bi[k] = gi; }
                    }
                }
            }
        }
    }
}

__device__ __forceinline__ int match_decision(const float best, const int bi, const int G, const float fg_thr, const float bg_thr) {
    int m = -2;
    if (G > 0) {
        if (best < bg_thr) m = -1;
        if (best > fg_thr) m = bi;
    }
    return m;
}

// Matches the K anchors a[] of this thread (`live[k]` = it exists) against the G boxes gt[0..G) of one image.  Must
// be called by ALL MATCH_BLOCK threads of the CTA (barriers inside); s_box / s_area are GT_TILE-sized shared
// buffers.  m[k] = -2 ignore, -1 background, g >= 0 matched GT index (box_utils.py:51-80).  K > 1 amortises the
// per-warp work (bounding-box reductions, the cull loop over all GT boxes, staging) over 32*K anchors.
template <bool FAST, int K>
__device__ __forceinline__ void match_block(float4 *s_box, float *s_area, const float4 (&a)[K], const bool (&live)[K],
                                            const float4 *__restrict__ gt, const int G, const float fg_thr,
                                            const float bg_thr, const float prune_c, int (&m)[K]) {
    float aa[K];
#pragma unroll
    for (int k = 0; k < K; ++k) aa[k] = box_area(a[k]);
    const WarpCull w = warp_cull_setup<FAST, K>(a, aa, live, prune_c);
    float best[K];
    int bi[K];
#pragma unroll
    for (int k = 0; k < K; ++k) { best[k] = w.fast ? 0.0f : -INFINITY; bi[k] = 0; }
    for (int t0 = 0; t0 < G; t0 += GT_TILE) {
        const int tn = min(GT_TILE, G - t0);
        __syncthreads();
        stage_gt_tile(s_box, s_area, gt + t0, tn, threadIdx.x, MATCH_BLOCK);
        __syncthreads();
        match_tile<K>(s_box, s_area, tn, t0, a, aa, w, prune_c, best, bi);
    }
#pragma unroll
    for (int k = 0; k < K; ++k) m[k] = match_decision(best[k], bi[k], G, fg_thr, bg_thr);
}

// Packed per-anchor target of the loss kernels: -2 / -1 / (g | cls << 20); labels are 1-based (README.md:132), cls =
// label-1 for labels 1..2047.  Any other label gets cls = 2047, "no class column" (C <= 2047 is enforced by rn_loss):
// label 0 is the reference's class id 0, whose one-hot row is all zeros after the [:,1:] slice (losses.py:96-103) —
// the anchor stays foreground (regression term, F count) with all-negative class targets; labels the 11 bits cannot
// hold no longer corrupt the GT index or alias the -1 / -2 codes.
constexpr int kNoClass = 2047;
__device__ __forceinline__ int pack_code(int m, const long long *__restrict__ labels) {
    if (m < 0) return m;
    const long long lab = labels[m];
    const int cls = (lab >= 1 && lab <= (long long)kNoClass) ? (int)lab - 1 : kNoClass;
    return m | (cls << 20);
}

}  // namespace rnmatch
