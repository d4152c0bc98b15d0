// Element-wise box coding entry points.
// rn_encode replaces bbox_2_activ (retinanet/box_utils.py:25-34); rn_decode replaces activ_2_bbox
// (retinanet/box_utils.py:37-48) including its exp(dx,dy) quirk.  HBM-bound: 32 B in, 16 B out per
// box, one 128-bit load per operand and one 128-bit store per thread.
#include "rn_common.cuh"

namespace {

template <bool ENCODE>
__global__ void __launch_bounds__(256) boxcode_kernel(const float4 *__restrict__ in, const float4 *__restrict__ anchors,
                                                      long long n, float4 wts, float4 *__restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = ENCODE ? rn::encode_box(in[i], anchors[i], wts) : rn::decode_box(in[i], anchors[i], wts);
}

template <bool ENCODE>
int launch(const float *in, const float *anchors, int64_t n, const float *w, float *out, rn_stream_t stream,
           const char *name) {
    RN_CHECK_ARG(n >= 0, RN_E_BADARG, "%s: negative size", name);
    if (n == 0) return 0;
    RN_CHECK_ARG(in && anchors && w && out, RN_E_BADARG, "%s: null pointer", name);
    RN_CHECK_ARG(((((uintptr_t)in) | ((uintptr_t)anchors) | ((uintptr_t)out)) & 15) == 0, RN_E_BADARG,
                 "%s: pointers must be 16-byte aligned", name);
    const long long grid = (n + 255) / 256;
    boxcode_kernel<ENCODE><<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(
        (const float4 *)in, (const float4 *)anchors, (long long)n, make_float4(w[0], w[1], w[2], w[3]), (float4 *)out);
    RN_CHECK_LAUNCH(name);
    return 0;
}

}  // namespace

extern "C" int rn_encode(const float *boxes, const float *anchors, int64_t n, const float *weights_host, float *out,
                         rn_stream_t stream) {
    return launch<true>(boxes, anchors, n, weights_host, out, stream, "rn_encode");
}
extern "C" int rn_decode(const float *activations, const float *anchors, int64_t n, const float *weights_host,
                         float *out, rn_stream_t stream) {
    return launch<false>(activations, anchors, n, weights_host, out, stream, "rn_decode");
}
