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

// Per-warp matching state of one group of 32 anchors (one per lane).
struct WarpCull {
    bool fast;                       // warp-uniform: culling / pruning / non-NaN fast path allowed
    float bx1, by1, bx2, by2;        // bounding box of the warp's anchors
    float ag_lo, ag_hi;              // GT areas outside this range cannot reach bg_thr with any anchor of the warp
};

template <bool FAST>
__device__ __forceinline__ WarpCull warp_cull_setup(const float4 a, const float aa, const bool live, const float prune_c) {
    WarpCull w;
    w.fast = FAST;
    w.bx1 = w.by1 = w.bx2 = w.by2 = 0.f;
    w.ag_lo = 0.f;
    w.ag_hi = INFINITY;
    if (FAST) {
        // positive finite extents imply finite, ordered coordinates differences; NaN fails every test
        bool ok = !live || ((a.z - a.x) > 0.0f && (a.w - a.y) > 0.0f && aa <= 3.0e38f && fabsf(a.x) <= 3.0e38f && fabsf(a.y) <= 3.0e38f);
        w.fast = __all_sync(0xffffffffu, ok);
        w.bx1 = rn::warp_min(live ? a.x : INFINITY);
        w.by1 = rn::warp_min(live ? a.y : INFINITY);
        w.bx2 = rn::warp_max(live ? a.z : -INFINITY);
        w.by2 = rn::warp_max(live ? a.w : -INFINITY);
        const float amin = rn::warp_min(live ? aa : INFINITY), amax = rn::warp_max(live ? aa : 0.0f);
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

// One warp x one staged tile of `tn` GT boxes (global indices t0..t0+tn): cull, then evaluate the surviving pairs.
__device__ __forceinline__ void match_tile(const float4 *s_box, const float *s_area, const int tn, const int t0,
                                           const float4 a, const float aa, const WarpCull &w, const float prune_c,
                                           float &best, int &bi) {
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
                // well-formed pair: no NaN possible, IoU is +0 unless both extents are positive
                float ww = __fsub_rn(fminf(g.z, a.z), fmaxf(g.x, a.x));
                float hh = __fsub_rn(fminf(g.w, a.w), fmaxf(g.y, a.y));
                if (ww > 0.0f && hh > 0.0f) {
                    float inter = __fmul_rn(ww, hh);
                    float uni = __fsub_rn(__fadd_rn(ag, aa), inter);
                    if (inter >= __fmul_rn(uni, prune_c)) {   // may reach bg_thr: exact quotient
                        float v = __fdiv_rn(inter, uni);
                        if (v > best) { best = v; bi = gi; }
                    }
                }
            } else {
                float v = iou_generic(g, box_area(g), a, aa);   // s_area holds the NaN marker for malformed boxes
                if (best == best) {                             // NaN, once taken, stays
                    if (v != v || v > best) { best = v; bi = gi; }
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

// Matches anchor `a` (this thread's; `live` = it exists) against the G boxes gt[0..G) of one image.  Must be
// called by ALL MATCH_BLOCK threads of the CTA (barriers inside); s_box / s_area are GT_TILE-sized shared
// buffers.  Returns -2 ignore, -1 background, g >= 0 matched GT index (box_utils.py:51-80).
template <bool FAST>
__device__ __forceinline__ int match_block(float4 *s_box, float *s_area, const float4 a, const bool live,
                                           const float4 *__restrict__ gt, const int G, const float fg_thr,
                                           const float bg_thr, const float prune_c) {
    const float aa = box_area(a);
    const WarpCull w = warp_cull_setup<FAST>(a, aa, live, prune_c);
    float best = w.fast ? 0.0f : -INFINITY;
    int bi = 0;
    for (int t0 = 0; t0 < G; t0 += GT_TILE) {
        const int tn = min(GT_TILE, G - t0);
        __syncthreads();
        stage_gt_tile(s_box, s_area, gt + t0, tn, threadIdx.x, MATCH_BLOCK);
        __syncthreads();
        match_tile(s_box, s_area, tn, t0, a, aa, w, prune_c, best, bi);
    }
    return match_decision(best, bi, G, fg_thr, bg_thr);
}

// Packed per-anchor target of the loss kernels: -2 / -1 / (g | (label-1) << 20); labels are 1-based (README.md:132).
__device__ __forceinline__ int pack_code(int m, const long long *__restrict__ labels) {
    return m >= 0 ? (m | (((int)labels[m] - 1) << 20)) : m;
}

}  // namespace rnmatch
