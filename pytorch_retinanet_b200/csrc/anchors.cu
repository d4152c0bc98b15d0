// Anchor grid for all pyramid levels in one launch.
// Replaces AnchorGenerator._compute_grid_offsets / grid_anchors (retinanet/anchors.py:151-197).
//
// HBM-bound, write-only: 16 B per anchor, one float4 store per thread, fully coalesced.  The
// shift is evaluated as offset*stride + i*stride in double and rounded to fp32 once, which is what
// ATen's CPU arange does (exact for the default offset 0 and any dyadic offset), followed by ONE
// fp32 add per coordinate exactly as anchors.py:190-194.
#include "rn_common.cuh"

namespace {

struct LevelTable {
    int n;
    int H[RN_MAX_LEVELS], W[RN_MAX_LEVELS], stride[RN_MAX_LEVELS], na[RN_MAX_LEVELS];
    int cell_off[RN_MAX_LEVELS];        // row offset into `cells`
    long long out_off[RN_MAX_LEVELS + 1];  // anchor offset of each level
};

__global__ void __launch_bounds__(256) anchor_grid_kernel(const float4 *__restrict__ cells, LevelTable t,
                                                          double offset, float4 *__restrict__ out,
                                                          long long total) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int l = 0;
#pragma unroll
    for (int k = 1; k < RN_MAX_LEVELS; ++k)
        if (k < t.n && i >= t.out_off[k]) l = k;
    long long r = i - t.out_off[l];
    int na = t.na[l];
    int a = (int)(r % na);
    long long loc = r / na;
    int x = (int)(loc % t.W[l]);
    int y = (int)(loc / t.W[l]);
    double st = (double)t.stride[l];
    float sx = (float)(offset * st + (double)x * st);
    float sy = (float)(offset * st + (double)y * st);
    float4 c = cells[t.cell_off[l] + a];
    out[i] = make_float4(__fadd_rn(sx, c.x), __fadd_rn(sy, c.y), __fadd_rn(sx, c.z), __fadd_rn(sy, c.w));
}

}  // namespace

extern "C" int rn_anchor_grid(const float *cells, const int32_t *level_desc_host, int num_levels, double offset,
                              float *out_anchors, int64_t num_anchors, rn_stream_t stream) {
    RN_CHECK_ARG(cells && level_desc_host && out_anchors, RN_E_BADARG, "rn_anchor_grid: null pointer");
    RN_CHECK_ARG(num_levels >= 1 && num_levels <= RN_MAX_LEVELS, RN_E_TOOLARGE,
                 "rn_anchor_grid: num_levels=%d outside [1,%d]", num_levels, RN_MAX_LEVELS);
    LevelTable t;
    t.n = num_levels;
    long long total = 0;
    int coff = 0;
    for (int l = 0; l < RN_MAX_LEVELS; ++l) {
        if (l < num_levels) {
            const int32_t *d = level_desc_host + 4 * l;
            RN_CHECK_ARG(d[0] >= 0 && d[1] >= 0 && d[2] > 0 && d[3] > 0, RN_E_BADARG,
                         "rn_anchor_grid: bad level %d descriptor {%d,%d,%d,%d}", l, d[0], d[1], d[2], d[3]);
            t.H[l] = d[0]; t.W[l] = d[1]; t.stride[l] = d[2]; t.na[l] = d[3];
            t.cell_off[l] = coff;
            t.out_off[l] = total;
            coff += d[3];
            total += (long long)d[0] * d[1] * d[3];
        } else {
            t.H[l] = t.W[l] = 0; t.stride[l] = t.na[l] = 1; t.cell_off[l] = 0; t.out_off[l] = total;
        }
    }
    t.out_off[RN_MAX_LEVELS] = total;
    RN_CHECK_ARG(total == num_anchors, RN_E_BADARG, "rn_anchor_grid: level table yields %lld anchors, caller says %lld",
                 total, (long long)num_anchors);
    if (total == 0) return 0;
    int block = 256;
    long long grid = (total + block - 1) / block;
    anchor_grid_kernel<<<(unsigned)grid, block, 0, (cudaStream_t)stream>>>((const float4 *)cells, t, offset,
                                                                           (float4 *)out_anchors, total);
    RN_CHECK_LAUNCH("rn_anchor_grid");
    return 0;
}
