// Shared fp32 math of the focal-loss kernels (loss.cu: [N,A,C] layout, levels.cu: per-level NCHW layout).
#pragma once
#include "rn_common.cuh"

namespace rnloss {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kSmallE = 0.0625f;               // series path valid for e <= 1/16
constexpr float kSmallX = -2.7725887f;           // x <= ln(1/16)

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// log1p(e) for 0 <= e <= 1/16: alternating series, truncation error e^6/7 < 1e-8
__device__ __forceinline__ float log1p_small(float e) {
    float s = fmaf(e, -1.0f / 6.0f, 0.2f);
    s = fmaf(e, s, -0.25f);
    s = fmaf(e, s, 1.0f / 3.0f);
    s = fmaf(e, s, -0.5f);
    s = fmaf(e, s, 1.0f);
    return e * s;
}

// p = sigmoid(x), sp = softplus(x) = log(1 + exp(x))
template <bool PRECISE>
__device__ __forceinline__ void sigmoid_softplus(float x, float &p, float &sp) {
    if (PRECISE) {
        float e = expf(-fabsf(x));
        float r = __fdiv_rn(1.0f, 1.0f + e);
        p = x >= 0.0f ? r : e * r;
        sp = fmaxf(x, 0.0f) + log1pf(e);
    } else {
        float e = ex2_approx(-fabsf(x) * kLog2e);
        float d = 1.0f + e;
        float r = rcp_approx(d);
        p = x >= 0.0f ? r : e * r;
        float l = e <= kSmallE ? log1p_small(e) : lg2_approx(d) * kLn2;
        sp = fmaxf(x, 0.0f) + l;
    }
}
// warp-uniform small-x path (x <= -2.77): e = exp(x) <= 1/16
__device__ __forceinline__ void sigmoid_softplus_small(float v, float &p, float &sp) {
    float e = ex2_approx(fmaf(v, kLog2e, kLog2e));   // exp(v + 1)
    p = e * rcp_approx(1.0f + e);
    sp = log1p_small(e);
}

template <bool GAMMA2>
__device__ __forceinline__ float pow_gamma(float b, float gamma) {
    if (GAMMA2) return b * b;
    // generic gamma (not the default): lg2.approx's absolute error, multiplied by gamma, would cost ~1e-5 of relative
    // accuracy on the loss, so the logarithm is libdevice's log2f (<= 1 ulp); ex2.approx (2 ulp relative) is enough
    return b > 0.0f ? ex2_approx(gamma * log2f(b)) : (gamma == 0.0f ? 1.0f : 0.0f);
}

}  // namespace rnloss
