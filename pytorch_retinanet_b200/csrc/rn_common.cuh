// Shared helpers for the retinanet_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "retinanet_b200.h"

#define RN_SM_COUNT_B200 148

void rn_set_error(const char *fmt, ...);
void rn_note_launch(void);   // counts the kernel launches this library issues (rn_launch_count)

#define RN_CHECK_ARG(cond, code, ...)   \
    do {                                \
        if (!(cond)) {                  \
            rn_set_error(__VA_ARGS__);  \
            return (code);              \
        }                               \
    } while (0)

#define RN_CHECK_LAUNCH(name)                                                      \
    do {                                                                           \
        cudaError_t e__ = cudaGetLastError();                                      \
        if (e__ != cudaSuccess) {                                                  \
            rn_set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));  \
            return (int)e__;                                                       \
        }                                                                          \
        rn_note_launch();                                                          \
    } while (0)

namespace rn {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// 128-bit streaming load: read-only path, do not allocate in L1 (data is touched exactly once).
__device__ __forceinline__ float4 ld_stream_f4(const float4 *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
// 128-bit streaming store (write-once gradients).
__device__ __forceinline__ void st_stream_f4(float4 *p, const float4 &v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}

// Reference box helpers — every operation rounds separately (no FMA contraction), mirroring the
// reference's eager ATen ops so that index decisions are bit-exact.
struct BoxCS {
    float cx, cy, w, h;
};
// convert_xywh, retinanet/box_utils.py:11-15: centre = (tl + br) / 2, size = br - tl
__device__ __forceinline__ BoxCS to_center_size(const float4 b) {
    BoxCS r;
    r.cx = __fdiv_rn(__fadd_rn(b.x, b.z), 2.0f);
    r.cy = __fdiv_rn(__fadd_rn(b.y, b.w), 2.0f);
    r.w = __fsub_rn(b.z, b.x);
    r.h = __fsub_rn(b.w, b.y);
    return r;
}

// activ_2_bbox, retinanet/box_utils.py:37-48 (+ convert_x1y1x2y2 :18-22) for one anchor.
// NB the reference quirk, reproduced on purpose: sizes = a_wh * exp(d.xy); d.zw are ignored.
__device__ __forceinline__ float4 decode_box(const float4 d, const float4 anchor, const float4 wts) {
    BoxCS a = to_center_size(anchor);
    float dx = __fdiv_rn(d.x, wts.x), dy = __fdiv_rn(d.y, wts.y);
    float cx = __fadd_rn(__fmul_rn(a.w, dx), a.cx);
    float cy = __fadd_rn(__fmul_rn(a.h, dy), a.cy);
    float w = __fmul_rn(a.w, expf(dx));
    float h = __fmul_rn(a.h, expf(dy));
    float hw = __fdiv_rn(w, 2.0f), hh = __fdiv_rn(h, 2.0f);
    return make_float4(__fsub_rn(cx, hw), __fsub_rn(cy, hh), __fadd_rn(cx, hw), __fadd_rn(cy, hh));
}

// bbox_2_activ, retinanet/box_utils.py:25-34 for one (gt, anchor) pair.
__device__ __forceinline__ float4 encode_box(const float4 gt, const float4 anchor, const float4 wts) {
    BoxCS g = to_center_size(gt), a = to_center_size(anchor);
    float4 t;
    t.x = __fmul_rn(__fdiv_rn(__fsub_rn(g.cx, a.cx), a.w), wts.x);
    t.y = __fmul_rn(__fdiv_rn(__fsub_rn(g.cy, a.cy), a.h), wts.y);
    t.z = __fmul_rn(logf(__fadd_rn(__fdiv_rn(g.w, a.w), 1e-8f)), wts.z);
    t.w = __fmul_rn(logf(__fadd_rn(__fdiv_rn(g.h, a.h), 1e-8f)), wts.w);
    return t;
}

// torch.clamp(min=lo, max=hi) on a float (NaN propagates like ATen's clamp).
__device__ __forceinline__ float clampf(float v, float lo, float hi) {
    return (v != v) ? v : fminf(fmaxf(v, lo), hi);
}

// sigmoid exactly as ATen's CUDA kernel evaluates it in fp32: 1 / (1 + exp(-x)).
__device__ __forceinline__ float sigmoid_ref(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }

}  // namespace rn
