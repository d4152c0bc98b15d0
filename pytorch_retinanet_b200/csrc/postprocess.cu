// Inference post-processing for a whole batch.
// Replaces Retinanet.process_detections (retinanet/models.py:160-243) and the torchvision ops it
// calls (clip_boxes_to_image, remove_small_boxes, nms).
//
// Two algorithms share the streaming filter K1:
//  * LAZY (default): one CTA per image runs greedy class-aware NMS in GLOBAL score order and stops as
//    soon as max_det boxes are kept.  Kept status of a candidate only depends on higher-ranked
//    candidates of its own class, so the first max_det kept boxes in global order ARE the
//    reference's output; the top candidates are obtained with a radix select (8-bit digits, MSB
//    first, early exit) + a 1024-key bitonic sort, LZ_ROUNDS rounds at most.  If an image still has
//    unprocessed candidates and fewer than max_det boxes after the last round, a flag is raised and
//    the host re-runs the batch with the general algorithm.
//  * GENERAL: per-(image,class) segments, exact for any number of candidates (pipeline below).
//
// General pipeline (all on one stream, no host synchronisation):
//  K1 score_filter   HBM-bound streaming pass over the [N,A,C] logits with 128-bit loads.  The
//                    strict test sigmoid(x) > thr is decided by ONE float compare against a guard
//                    logit x_lo (slightly below logit(thr)); only the rare survivors evaluate the
//                    exact sigmoid 1/(1+exp(-x)) (bit-identical to torch's CUDA kernel), decode +
//                    clip their box (reference quirk included) and apply the 0.01 small-box filter.
//                    Survivors are staged in shared memory and flushed to a global pool with one
//                    atomic per CTA; a histogram over (image,class) segments is built on the fly.
//                    key = (~score_bits << 32) | anchor  — ascending key = score desc, anchor asc.
//  K2 segment_scan   exclusive scan of the segment histogram (one CTA).
//  K3 scatter        counting-sort scatter of the pool into segment-contiguous order.
//  K4 nms            one CTA per (image,class) segment: bitonic sort of the keys (shared memory,
//                    or in place in global memory for huge segments), then greedy NMS in chunks of
//                    256: (a) suppression by previously kept boxes, (b) 256x256 IoU bitmask,
//                    (c) warp-cooperative serial resolve of the bitmask, (d) compaction of the kept
//                    boxes.  Exact torchvision CPU semantics: stable score order (ties -> lower
//                    anchor index), suppress iff fl(inter/(a_i+a_j-inter)) > thr, per class only.
//  K5 image_topk     one CTA per image: 4-pass 8-bit radix select of the max_det-th score over
//                    all kept boxes, deterministic tie handling (class asc, anchor asc), bitonic
//                    sort of the <= max_det winners, output write (labels + 1).
#include <cstring>

#include "rn_common.cuh"

namespace {

typedef unsigned long long u64;
typedef unsigned int u32;

constexpr int PP_BLOCK = 256;
constexpr int PP_WSPAN = 128;      // anchors per warp task in K1
constexpr int PP_U = 4;            // 128-bit loads in flight per lane
constexpr int PP_WSTAGE = 96;      // staged candidates per warp
constexpr int NMS_CHUNK = 256;
constexpr int SORT_SMEM = 2048;    // keys sorted in shared memory
constexpr int MAX_DET_CAP = 1024;
constexpr int LZ_BLOCK = 1024;
constexpr int LZ_M = 1024;         // candidates taken per round (radix select + sort)
constexpr int LZ_ROUNDS = 5;
constexpr int LZ_CACHE = 16384;    // candidate keys cached in shared memory (128 KB)
constexpr int LZ_CHUNK = 256;

struct PPWorkspace {
    u32 *pool_count;     // [1] candidates found (may exceed capacity)
    u32 *img_count;      // [N] candidates found per image (LAZY; may exceed cap_n)
    int *need_v1;        // [N] LAZY: image left to lazy_nms_kernel by lazy2_nms_kernel
    int *seg_count;      // [S]
    int *seg_off;        // [S+1]
    int *cursor;         // [S]
    int *kept_count;     // [S]
    u64 *pool_key;       // [cap]
    u32 *pool_seg;       // [cap]
    u64 *sorted_key;     // [cap]
    u64 *kept_key;       // [cap]
    float4 *kept_box;    // [cap]
    size_t zero_bytes;   // leading bytes that must be zeroed per call
    size_t total_bytes;
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

PPWorkspace carve(void *base, int N, int C, int64_t cap) {
    PPWorkspace w;
    const size_t S = (size_t)N * (size_t)C;
    char *p = (char *)base;
    size_t o = 0;
    w.pool_count = (u32 *)(p + o); o += 256;
    w.img_count = (u32 *)(p + o); o = align_up(o + (size_t)N * 4, 256);
    w.need_v1 = (int *)(p + o); o = align_up(o + (size_t)N * 4, 256);
    w.seg_count = (int *)(p + o); o = align_up(o + S * 4, 256);
    w.zero_bytes = o;
    w.seg_off = (int *)(p + o); o = align_up(o + (S + 1) * 4, 256);
    w.cursor = (int *)(p + o); o = align_up(o + S * 4, 256);
    w.kept_count = (int *)(p + o); o = align_up(o + S * 4, 256);
    w.pool_key = (u64 *)(p + o); o = align_up(o + (size_t)cap * 8, 256);
    w.pool_seg = (u32 *)(p + o); o = align_up(o + (size_t)cap * 4, 256);
    w.sorted_key = (u64 *)(p + o); o = align_up(o + (size_t)cap * 8, 256);
    w.kept_key = (u64 *)(p + o); o = align_up(o + (size_t)cap * 8, 256);
    w.kept_box = (float4 *)(p + o); o = align_up(o + (size_t)cap * 16, 256);
    w.total_bytes = o;
    return w;
}

// Where the box activations live: the reference head's [N,A,4] tensor, or (row N1) the raw per-level conv outputs
// [N, na*4, H, W] — only the few thousand candidates that reach NMS are ever read, so the level layout is indexed
// in place (4 strided loads per candidate) instead of being re-laid out.
struct BoxSource {
    const float4 *nac;                    // [N,A,4] or null
    const float *lvl[RN_MAX_LEVELS];      // level l: [N, na_l*4, HW_l]
    long long off[RN_MAX_LEVELS + 1];     // anchor offsets of the levels
    int HW[RN_MAX_LEVELS], na[RN_MAX_LEVELS];
    int nlev;
};

__device__ __forceinline__ float4 load_activation(const BoxSource &B, const int n, const long long A, const long long anchor) {
    if (B.nac) return __ldg(B.nac + (long long)n * A + anchor);
    int l = 0;
#pragma unroll
    for (int k = 1; k < RN_MAX_LEVELS; ++k)
        if (k < B.nlev && anchor >= B.off[k]) l = k;
    const int na = B.na[l], HW = B.HW[l];
    const int r = (int)(anchor - B.off[l]);
    const int pos = r / na, a = r - pos * na;
    const float *b0 = B.lvl[l] + ((long long)(n * na + a) * 4) * HW + pos;
    return make_float4(__ldg(b0), __ldg(b0 + HW), __ldg(b0 + 2LL * HW), __ldg(b0 + 3LL * HW));
}

// decode + clip of one anchor: activ_2_bbox (box_utils.py:37-48) then clip_boxes_to_image
// (tv:ops/boxes.py:149-182: x in [0,w], y in [0,h]).
__device__ __noinline__ float4 decode_clip(const BoxSource &box, const int n, const long long A, const long long anchor,
                                           const float4 *__restrict__ anchors, long long anchor_row, float4 wts, float imw,
                                           float imh) {
    float4 b = rn::decode_box(load_activation(box, n, A, anchor), __ldg(anchors + anchor_row), wts);
    b.x = rn::clampf(b.x, 0.0f, imw);
    b.z = rn::clampf(b.z, 0.0f, imw);
    b.y = rn::clampf(b.y, 0.0f, imh);
    b.w = rn::clampf(b.w, 0.0f, imh);
    return b;
}

// ------------------------------------------------------------------------------------------- K1
struct FilterParams {
    const float *logits;
    const float4 *bbox;
    const float4 *anchors;
    const int *im_hw;
    long long A;
    long long anchor_stride;
    int C;
    unsigned magic;
    float x_lo, thr;
    float4 wts;
    u32 cap;
    u32 cap_n;           // LAZY: per-image capacity of the candidate lists (cap / N)
    PPWorkspace w;
};

// Rare path, deliberately NOT inlined: the streaming loop stays a few dozen instructions (the
// v1 kernel inlined 16 copies of sigmoid+decode and stalled on instruction fetch).
// Evaluates the exact score of the <= 4 logits of one vector and stages the survivors in the
// warp's shared-memory buffer.  Box decoding / small-box filtering is deferred to the NMS kernels,
// where it runs on converged warps instead of one lane at a time.
// LAZY keys: (~score_bits << 32) | (class*A + anchor)  -> ascending = score desc, class asc, anchor asc
// GENERAL  : (~score_bits << 32) | anchor, plus the (image,class) segment id
template <bool LAZY>
__device__ __noinline__ void emit_candidates(const FilterParams &P, float4 v, int nvals, int n, int c0,
                                             long long anchor, u64 *st_key, u32 *st_seg, int *st_n) {
    const float vals[4] = {v.x, v.y, v.z, v.w};
    for (int k = 0; k < nvals; ++k) {
        if (!(vals[k] > P.x_lo)) continue;
        const float s = rn::sigmoid_ref(vals[k]);
        if (!(s > P.thr)) continue;                                // models.py:196 strict >
        const u32 lo = LAZY ? (u32)((long long)(c0 + k) * P.A + anchor) : (u32)anchor;
        const u64 key = ((u64)(~__float_as_uint(s)) << 32) | (u64)lo;
        const u32 seg = (u32)(n * P.C + c0 + k);
        const int pos = atomicAdd(st_n, 1);
        if (pos < PP_WSTAGE) {
            st_key[pos] = key;
            if (!LAZY) st_seg[pos] = seg;
        } else if (LAZY) {                                         // staging full: go straight to the list
            const u32 gp = atomicAdd(P.w.img_count + n, 1u);
            if (gp < P.cap_n) P.w.pool_key[(size_t)n * P.cap_n + gp] = key;
        } else {
            const u32 gp = atomicAdd(P.w.pool_count, 1u);
            if (gp < P.cap) {
                P.w.pool_key[gp] = key;
                P.w.pool_seg[gp] = seg;
                atomicAdd(P.w.seg_count + seg, 1);
            }
        }
    }
}

// Warp-autonomous streaming filter: every warp owns PP_WSPAN consecutive anchors of one image, reads
// them with 128-bit loads (PP_U in flight per lane), stages its rare survivors in its own slice of
// shared memory and flushes them with ONE global atomic per warp.  No block-level barrier anywhere.
template <int VEC, bool LAZY>
__global__ void __launch_bounds__(PP_BLOCK, 4) score_filter_kernel(const __grid_constant__ FilterParams P) {
    __shared__ u64 s_key[PP_BLOCK / 32][PP_WSTAGE];
    __shared__ u32 s_seg[LAZY ? 1 : PP_BLOCK / 32][LAZY ? 1 : PP_WSTAGE];
    __shared__ int s_n[PP_BLOCK / 32];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = blockIdx.y;
    const long long a0 = ((long long)blockIdx.x * (PP_BLOCK / 32) + warp) * PP_WSPAN;
    if (a0 >= P.A) return;
    const int span = (int)min((long long)PP_WSPAN, P.A - a0);
    const long long row0 = (long long)n * P.A + a0;
    const int CV = P.C / VEC;
    const int nvec = span * CV;
    const float *src = P.logits + row0 * P.C;
    u64 *st_key = s_key[warp];
    u32 *st_seg = LAZY ? nullptr : s_seg[warp];
    int *st_n = &s_n[warp];
    if (lane == 0) *st_n = 0;
    __syncwarp();

    for (int base = 0; base < nvec; base += 32 * PP_U) {
        float4 v[PP_U];
#pragma unroll
        for (int u = 0; u < PP_U; ++u) {
            const int f = base + u * 32 + lane;
            v[u] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
            if (f < nvec) {
                if (VEC == 4) v[u] = rn::ld_stream_f4((const float4 *)src + f);
                else v[u].x = __ldg(src + f);
            }
        }
#pragma unroll
        for (int u = 0; u < PP_U; ++u) {
            const float vmax = VEC == 4 ? fmaxf(fmaxf(v[u].x, v[u].y), fmaxf(v[u].z, v[u].w)) : v[u].x;
            if (vmax > P.x_lo) {                                   // rare
                const int f = base + u * 32 + lane;
                const int al = P.magic ? (int)__umulhi((unsigned)f, P.magic) : f;
                emit_candidates<LAZY>(P, v[u], VEC, n, (f - al * CV) * VEC, a0 + al, st_key, st_seg, st_n);
            }
        }
    }
    __syncwarp();
    const int staged = min(*st_n, PP_WSTAGE);
    if (staged == 0) return;
    u32 gbase = 0;
    if (lane == 0) gbase = atomicAdd(LAZY ? P.w.img_count + n : P.w.pool_count, (u32)staged);
    gbase = __shfl_sync(0xffffffffu, gbase, 0);
    for (int i = lane; i < staged; i += 32) {
        const u32 gp = gbase + (u32)i;
        if (LAZY) {
            if (gp < P.cap_n) P.w.pool_key[(size_t)n * P.cap_n + gp] = st_key[i];
        } else if (gp < P.cap) {
            P.w.pool_key[gp] = st_key[i];
            P.w.pool_seg[gp] = st_seg[i];
            atomicAdd(P.w.seg_count + st_seg[i], 1);
        }
    }
}

// ------------------------------------------------------------------------------------------- K2
__global__ void __launch_bounds__(1024) segment_scan_kernel(const int *__restrict__ seg_count, int S,
                                                            int *__restrict__ seg_off, int *__restrict__ cursor,
                                                            const u32 *__restrict__ pool_count, u32 cap,
                                                            int *__restrict__ status) {
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        s_carry = 0;
        if (status) { status[0] = (int)min(*pool_count, 0x7fffffffu); status[1] = (int)cap; }
    }
    __syncthreads();
    for (int base = 0; base < S; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < S ? seg_count[i] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int y = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += y;
            }
            s_warp[lane] = w;
        }
        __syncthreads();
        const int excl = s_carry + (warp ? s_warp[warp - 1] : 0) + x - v;
        if (i < S) { seg_off[i] = excl; cursor[i] = excl; }
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) seg_off[S] = s_carry;
}

// ------------------------------------------------------------------------------------------- K3
__global__ void __launch_bounds__(256) scatter_kernel(const u64 *__restrict__ pool_key, const u32 *__restrict__ pool_seg,
                                                      const u32 *__restrict__ pool_count, u32 cap,
                                                      int *__restrict__ cursor, u64 *__restrict__ sorted_key) {
    const u32 total = min(*pool_count, cap);
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int pos = atomicAdd(cursor + pool_seg[i], 1);
        sorted_key[pos] = pool_key[i];
    }
}

// ------------------------------------------------------------------------------------------- K4
// Bitonic sort (ascending) of n keys for arbitrary n: every merge runs in the same direction
// (first step pairs i with i ^ (k-1)), so absent elements behave as +inf padding at the end.
__device__ __noinline__ void block_bitonic_sort(u64 *keys, int n) {
    for (int k = 2; (k >> 1) < n; k <<= 1) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int j = i ^ (k - 1);
            if (j > i && j < n) {
                const u64 a = keys[i], b = keys[j];
                if (a > b) { keys[i] = b; keys[j] = a; }
            }
        }
        __syncthreads();
        for (int d = k >> 2; d > 0; d >>= 1) {
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const int j = i ^ d;
                if (j > i && j < n) {
                    const u64 a = keys[i], b = keys[j];
                    if (a > b) { keys[i] = b; keys[j] = a; }
                }
            }
            __syncthreads();
        }
    }
}

// torchvision nms (CPU kernel) pair test: suppressed iff fl(inter / (area_i + area_j - inter)) > thr.
__device__ __forceinline__ bool nms_suppresses(const float4 a, float area_a, const float4 b, float area_b, float thr) {
    const float w = __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x));
    // disjoint in x (the common case): inter = 0 -> ovr is 0 (or NaN if both areas are 0), never > thr >= 0
    if (!(w > 0.0f) && thr >= 0.0f) return false;
    const float h = __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y));
    if (!(h > 0.0f) && thr >= 0.0f) return false;
    const float inter = __fmul_rn(fmaxf(w, 0.0f), fmaxf(h, 0.0f));
    const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter));
    return ovr > thr;
}
__device__ __forceinline__ float nms_area(const float4 b) {
    return __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
}

struct NmsParams {
    BoxSource box;           // DECODE mode: box activations
    const float4 *anchors;
    const int *im_hw;
    const float4 *boxes;     // RAW mode: [K,4] boxes already sorted per segment
    unsigned char *keep_flags;  // RAW mode output
    long long A;
    long long anchor_stride;
    int C;
    float thr;
    float4 wts;
    const int *seg_off;
    u64 *sorted_key;
    u64 *kept_key;
    float4 *kept_box;
    int *kept_count;
};

template <bool RAW>
__global__ void __launch_bounds__(NMS_CHUNK) nms_kernel(const __grid_constant__ NmsParams P) {
    __shared__ u64 s_keys[RAW ? 1 : SORT_SMEM];
    __shared__ float4 s_box[NMS_CHUNK];
    __shared__ float s_area[NMS_CHUNK];
    __shared__ u32 s_mask[NMS_CHUNK][NMS_CHUNK / 32];
    __shared__ u32 s_removed[NMS_CHUNK / 32];
    __shared__ int s_wbase[NMS_CHUNK / 32 + 1];

    const int seg = blockIdx.x;
    const int off = P.seg_off[seg];
    const int cnt = P.seg_off[seg + 1] - off;
    if (cnt <= 0) {
        if (!RAW && threadIdx.x == 0) P.kept_count[seg] = 0;
        return;
    }
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    long long anc_row = 0;
    int n = 0;
    float imw = 0.f, imh = 0.f;
    u64 *keys = nullptr;
    if (!RAW) {
        n = seg / P.C;
        anc_row = (long long)n * P.anchor_stride;
        imh = (float)P.im_hw[2 * n];
        imw = (float)P.im_hw[2 * n + 1];
        keys = P.sorted_key + off;
        if (cnt <= SORT_SMEM) {
            for (int i = t; i < cnt; i += NMS_CHUNK) s_keys[i] = keys[i];
            __syncthreads();
            block_bitonic_sort(s_keys, cnt);
            for (int i = t; i < cnt; i += NMS_CHUNK) keys[i] = s_keys[i];
        } else {
            block_bitonic_sort(keys, cnt);
        }
        __syncthreads();
    }

    int K = 0;   // kept so far (block-uniform)
    for (int c0 = 0; c0 < cnt; c0 += NMS_CHUNK) {
        const int m = min(NMS_CHUNK, cnt - c0);
        // (0) load this chunk's boxes
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        u64 key = 0;
        if (t < m) {
            if (RAW) {
                b = P.boxes[off + c0 + t];
            } else {
                key = keys[c0 + t];
                const long long anchor = (long long)(u32)key;
                b = decode_clip(P.box, n, P.A, anchor, P.anchors, anc_row + anchor, P.wts, imw, imh);
            }
        }
        const float area = nms_area(b);
        // remove_small_boxes(min_size=1e-2), models.py:203: such candidates are neither kept nor suppress
        const bool box_ok = RAW || ((__fsub_rn(b.z, b.x) >= 0.01f) && (__fsub_rn(b.w, b.y) >= 0.01f));
        // (a) suppression by boxes kept in earlier chunks (tiles of NMS_CHUNK through shared memory)
        bool alive = t < m && box_ok;
        for (int k0 = 0; k0 < K; k0 += NMS_CHUNK) {
            const int kn = min(NMS_CHUNK, K - k0);
            __syncthreads();
            if (t < kn) {
                const float4 kb = P.kept_box[off + k0 + t];
                s_box[t] = kb;
                s_area[t] = nms_area(kb);
            }
            __syncthreads();
            if (alive) {
                for (int k = 0; k < kn; ++k)
                    if (nms_suppresses(s_box[k], s_area[k], b, area, P.thr)) { alive = false; break; }
            }
        }
        __syncthreads();
        // (b) in-chunk bitmask: row t = boxes j > t suppressed by t
        s_box[t] = b;
        s_area[t] = area;
        {
            const u32 dead = __ballot_sync(0xffffffffu, !alive);
            if (lane == 0) s_removed[warp] = dead;
        }
        __syncthreads();
#pragma unroll 1
        for (int w = 0; w < NMS_CHUNK / 32; ++w) {
            u32 bits = 0;
            if (alive && w * 32 + 31 > t && w * 32 < m) {
                const int jlo = max(w * 32, t + 1), jhi = min(w * 32 + 32, m);
                for (int j = jlo; j < jhi; ++j)
                    if (nms_suppresses(b, area, s_box[j], s_area[j], P.thr)) bits |= 1u << (j & 31);
            }
            s_mask[t][w] = bits;
        }
        __syncthreads();
        // (c) warp-cooperative suppression scan (parallel form of the greedy scan, see lazy_nms_kernel):
        // per 32-row group, K <- alive & no kept j<i with M[j][i], iterated to the fixed point with one
        // warp-wide OR reduction (REDUX) per sweep; kept rows are OR-reduced into the later words.
        if (warp == 0) {
            u32 rw = lane < NMS_CHUNK / 32 ? s_removed[lane] : 0u;
            const int groups = (m + 31) >> 5;
#pragma unroll 1
            for (int g = 0; g < groups; ++g) {
                const int row = g * 32 + lane;
                const u32 rem0 = __shfl_sync(0xffffffffu, rw, g);
                const bool valid = row < m && !((rem0 >> lane) & 1u);
                const u32 diag = valid ? s_mask[row][g] : 0u;
                bool kp = valid;
#pragma unroll 1
                for (int sweep = 0; sweep < 32; ++sweep) {
                    const u32 rem = __reduce_or_sync(0xffffffffu, kp ? diag : 0u);
                    const bool nk = valid && !((rem >> lane) & 1u);
                    const bool changed = nk != kp;
                    kp = nk;
                    if (!__any_sync(0xffffffffu, changed)) break;
                }
                const u32 keptm = __ballot_sync(0xffffffffu, kp);
                if (lane == g) rw = ~keptm;
#pragma unroll
                for (int w = 0; w < NMS_CHUNK / 32; ++w) {
                    if (w > g) {
                        const u32 acc = __reduce_or_sync(0xffffffffu, kp ? s_mask[row][w] : 0u);
                        if (lane == w) rw |= acc;
                    }
                }
            }
            if (lane < NMS_CHUNK / 32) s_removed[lane] = rw;
        }
        __syncthreads();
        // (d) compaction of the kept boxes of this chunk
        const bool kept = t < m && !((s_removed[warp] >> lane) & 1u);
        const u32 km = __ballot_sync(0xffffffffu, kept);
        if (lane == 0) s_wbase[warp + 1] = __popc(km);
        __syncthreads();
        if (t == 0) {
            s_wbase[0] = 0;
            for (int w = 0; w < NMS_CHUNK / 32; ++w) s_wbase[w + 1] += s_wbase[w];
        }
        __syncthreads();
        if (RAW) {
            if (t < m) P.keep_flags[off + c0 + t] = kept ? 1 : 0;
        }
        if (kept) {
            const int pos = off + K + s_wbase[warp] + __popc(km & ((1u << lane) - 1u));
            P.kept_box[pos] = b;
            if (!RAW) P.kept_key[pos] = key;
        }
        K += s_wbase[NMS_CHUNK / 32];
        __syncthreads();
    }
    if (!RAW && t == 0) P.kept_count[seg] = K;
}

// Output epilogue shared by both algorithms:
//  * ratio (optional, [N,2] = ratio_h, ratio_w): torchvision resize_boxes of the final boxes back to the original
//    image size, as transform.postprocess does right after the path (retinanet/models.py:271,
//    tv:models/detection/transform.py:306-319): x * ratio_w, y * ratio_h, one fp32 multiply each;
//  * format 1: COCO xywh, as CocoEvaluator.prepare_for_coco_detection converts them
//    (utils/coco/coco_eval.py:159-161): (x1, y1, x2 - x1, y2 - y1), applied after the resize.
__device__ __forceinline__ float4 finish_box(float4 b, const float *__restrict__ ratio, int n, int format) {
    if (ratio) {
        const float rh = __ldg(ratio + 2 * n), rw = __ldg(ratio + 2 * n + 1);
        b.x = __fmul_rn(b.x, rw); b.z = __fmul_rn(b.z, rw);
        b.y = __fmul_rn(b.y, rh); b.w = __fmul_rn(b.w, rh);
    }
    if (format == 1) { b.z = __fsub_rn(b.z, b.x); b.w = __fsub_rn(b.w, b.y); }
    return b;
}

// ------------------------------------------------------------------------------------------- LAZY
struct LazyParams {
    BoxSource box;
    const float4 *anchors;
    const int *im_hw;
    long long A;
    long long anchor_stride;
    int C;
    int N;
    float thr;
    float4 wts;
    int max_det;
    const u64 *cand_key;     // [N][cap_n]
    const u32 *img_count;    // [N]
    u32 cap_n;
    float *out_boxes;
    float *out_scores;
    long long *out_labels;
    int *out_count;
    int *status;             // [0] max candidates/image * N (atomicMax), [2] fallback flag
    const float *out_ratio;  // [N,2] or null (row N2)
    int out_format;          // 0 xyxy, 1 xywh (row N4)
    int topk;                // pre_nms_topk per (image, pyramid level); 0 = off (reference behaviour)
    int nlev;
    long long lvl_off[RN_MAX_LEVELS + 1];   // anchor offsets of the pyramid levels
    int *need_v1;            // [N] or null: written by lazy2_nms_kernel (1 = image left to lazy_nms_kernel), read by lazy_nms_kernel
    int bin_shift;           // lazy2: score-bin width (see lazy2_nms_kernel)
    int capacity;            // status[1]: candidate capacity of the call (N * cap_n), reported by lazy_nms_kernel
};

struct LazySmem {
    u64 sel[LZ_M];
    float4 box[LZ_CHUNK];
    float area[LZ_CHUNK];
    int cls[LZ_CHUNK];
    u32 mask[LZ_CHUNK][LZ_CHUNK / 32];
    float4 kbox[MAX_DET_CAP];
    float karea[MAX_DET_CAP];
    int kcls[MAX_DET_CAP];
    u32 kscore[MAX_DET_CAP];
    u32 hist[256];
    u32 removed[LZ_CHUNK / 32];
    int lvl_cnt[RN_MAX_LEVELS];                       // candidates seen so far per level (pre_nms_topk)
    int lvl_wcnt[LZ_CHUNK / 32][RN_MAX_LEVELS];
    int wbase[LZ_CHUNK / 32 + 1];
    u64 prefix, mask_bits, thr_key;
    u32 need;
    int nsel, done;
};

__global__ void __launch_bounds__(LZ_BLOCK, 1) lazy_nms_kernel(const __grid_constant__ LazyParams P) {
    extern __shared__ __align__(16) unsigned char lz_raw[];
    LazySmem &S = *reinterpret_cast<LazySmem *>(lz_raw);
    u64 *s_cand = reinterpret_cast<u64 *>(lz_raw + ((sizeof(LazySmem) + 15) & ~(size_t)15));

    const int n = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (n == 0 && t == 0) P.status[1] = P.capacity;  // (this kernel always closes the lazy pipeline)
    if (P.need_v1 && P.need_v1[n] == 0) return;     // lazy2_nms_kernel finished this image
#ifdef RN_LAZY_TIMING
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tprev = clock64();
    int n_rounds = 0, n_chunks = 0;
#define LZ_TICK(id) do { if (t == 0) { long long now_ = clock64(); tacc[id] += now_ - tprev; tprev = now_; } } while (0)
#else
#define LZ_TICK(id) do { } while (0)
#endif
    const u32 found = P.img_count[n];
    if (t == 0) {
        const unsigned long long scaled = (unsigned long long)found * (unsigned long long)P.N;
        atomicMax(P.status + 0, (int)min(scaled, 0x7fffffffULL));
    }
    const int K = (int)min(found, P.cap_n);
    if (K == 0) {
        if (t == 0) P.out_count[n] = 0;
        return;
    }
    if (t < RN_MAX_LEVELS) S.lvl_cnt[t] = 0;
    const u64 *g_cand = P.cand_key + (size_t)n * P.cap_n;
    const bool cached = K <= LZ_CACHE;
    if (cached) {
#pragma unroll 4
        for (int i = t; i < K; i += LZ_BLOCK) s_cand[i] = g_cand[i];
    }
    __syncthreads();
    const u64 *cand = cached ? s_cand : g_cand;
    LZ_TICK(0);
    const float imh = (float)P.im_hw[2 * n], imw = (float)P.im_hw[2 * n + 1];
    const long long anc_row = (long long)n * P.anchor_stride;
    const u32 A32 = (u32)P.A;

    int kept = 0, processed = 0;
    u64 last = 0;                                  // keys are > 0 (score <= 1 -> ~bits >= 0xC07FFFFF)
#pragma unroll 1
    // with pre_nms_topk the walk continues until every level has seen its k candidates (no round budget)
    for (int round = 0; (round < LZ_ROUNDS || P.topk > 0) && processed < K && kept < P.max_det; ++round) {
        if (P.topk > 0) {
            bool full = true;
            for (int l = 0; l < P.nlev; ++l) full = full && S.lvl_cnt[l] >= P.topk;
            if (full) { processed = K; break; }               // nothing below this rank is eligible any more
        }
        const int remaining = K - processed;
        const int cap_r = LZ_M;
        const int take = min(remaining, cap_r);
        // ---- threshold key T: the take-th smallest key among keys > last (radix select) ----
        u64 T = ~0ULL;
        if (remaining > cap_r) {
            if (t == 0) { S.prefix = 0; S.mask_bits = 0; S.need = (u32)take; S.done = 0; S.thr_key = ~0ULL; }
            __syncthreads();
#pragma unroll 1
            for (int shift = 56; shift >= 0; shift -= 8) {
                if (t < 256) S.hist[t] = 0;
                __syncthreads();
                const u64 prefix = S.prefix, mbits = S.mask_bits;
                // warp-aggregated histogram: lanes with the same digit elect one leader (match.any), so a
                // skewed digit (e.g. the 3 possible top bytes of a score in (0.05,1]) costs 3 shared-memory
                // atomics per warp instead of 32 serialised ones
#pragma unroll 1
                for (int i0 = warp * 32; i0 < K; i0 += LZ_BLOCK) {
                    const int i = i0 + lane;
                    u32 digit = 256u;
                    if (i < K) {
                        const u64 k = cand[i];
                        if (k > last && (k & mbits) == prefix) digit = (u32)(k >> shift) & 255u;
                    }
                    const u32 peers = __match_any_sync(0xffffffffu, digit);
                    if (digit < 256u && lane == __ffs(peers) - 1) atomicAdd(&S.hist[digit], (u32)__popc(peers));
                }
                __syncthreads();
                if (warp == 0) {
                    u32 cnt[8], sum = 0;
#pragma unroll
                    for (int i = 0; i < 8; ++i) { cnt[i] = S.hist[lane * 8 + i]; sum += cnt[i]; }
                    u32 incl = sum;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const u32 y = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o) incl += y;
                    }
                    const u32 need = S.need, excl = incl - sum;
                    __syncwarp();                                  // every lane has read S.need before one lane rewrites it
                    if (excl < need && need <= incl) {            // exactly one lane
                        u32 acc = excl;
                        int b = 0;
                        for (; b < 8; ++b) {
                            if (acc + cnt[b] >= need) break;
                            acc += cnt[b];
                        }
                        const u64 np = prefix | ((u64)(lane * 8 + b) << shift);
                        S.need = need - acc;
                        S.prefix = np;
                        S.mask_bits = mbits | (255ULL << shift);
                        if (need - acc == cnt[b] || shift == 0) {  // the whole bucket is taken: stop early
                            S.done = 1;
                            S.thr_key = np | (shift ? ((1ULL << shift) - 1ULL) : 0ULL);
                        }
                    }
                }
                __syncthreads();
                if (S.done) break;
            }
            T = S.thr_key;
        }
        LZ_TICK(1);
        // ---- gather + sort the selected keys ----
        if (t == 0) S.nsel = 0;
        __syncthreads();
#pragma unroll 1
        for (int i0 = warp * 32; i0 < K; i0 += LZ_BLOCK) {            // warp-aggregated append
            const int i = i0 + lane;
            const u64 k = i < K ? cand[i] : 0ULL;
            const bool pick = i < K && k > last && k <= T;
            const u32 pm = __ballot_sync(0xffffffffu, pick);
            int base = 0;
            if (lane == 0 && pm) base = atomicAdd(&S.nsel, __popc(pm));
            base = __shfl_sync(0xffffffffu, base, 0);
            const int slot = base + __popc(pm & ((1u << lane) - 1u));
            if (pick && slot < LZ_M) S.sel[slot] = k;
        }
        __syncthreads();
        const int nsel = min(S.nsel, LZ_M);
        LZ_TICK(2);
        block_bitonic_sort(S.sel, nsel);
        LZ_TICK(3);

        // ---- greedy class-aware NMS over the sorted keys, LZ_CHUNK at a time ----
#pragma unroll 1
        for (int c0 = 0; c0 < nsel && kept < P.max_det; c0 += LZ_CHUNK) {
            const int m = min(LZ_CHUNK, nsel - c0);
            float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
            int cls = -1, lvl = -1, lrank = 0;
            bool alive = false;
            if (t < LZ_CHUNK) {
                if (t < m) {
                    const u32 lo = (u32)S.sel[c0 + t];
                    cls = (int)(lo / A32);
                    const long long anchor = (long long)(lo - (u32)cls * A32);
                    b = decode_clip(P.box, n, P.A, anchor, P.anchors, anc_row + anchor, P.wts, imw, imh);
                    // remove_small_boxes(min_size=1e-2), models.py:203
                    alive = (__fsub_rn(b.z, b.x) >= 0.01f) && (__fsub_rn(b.w, b.y) >= 0.01f);
                    if (P.topk > 0) {
                        lvl = 0;
                        for (int l = 1; l < P.nlev; ++l) lvl += anchor >= P.lvl_off[l];
                    }
                }
                if (P.topk > 0) {                              // rank of this candidate inside its level (global order)
                    for (int l = 0; l < P.nlev; ++l) {
                        const u32 bl = __ballot_sync(0xffffffffu, lvl == l);
                        if (lane == 0) S.lvl_wcnt[warp][l] = __popc(bl);
                        if (lvl == l) lrank = __popc(bl & ((1u << lane) - 1u));
                    }
                }
            }
            if (P.topk > 0) {
                __syncthreads();
                if (t < LZ_CHUNK && lvl >= 0) {
                    int before = S.lvl_cnt[lvl] + lrank;
                    for (int w = 0; w < warp; ++w) before += S.lvl_wcnt[w][lvl];
                    alive = alive && before < P.topk;          // only the top-k scores of a level survive
                }
                __syncthreads();
                if (t < P.nlev) {
                    int add = 0;
                    for (int w = 0; w < LZ_CHUNK / 32; ++w) add += S.lvl_wcnt[w][t];
                    S.lvl_cnt[t] += add;
                }
            }
            if (t < LZ_CHUNK) {
                const float area = nms_area(b);
                if (alive) {                                   // (a) boxes kept so far, same class only
#pragma unroll 1
                    for (int k = 0; k < kept; ++k)
                        if (S.kcls[k] == cls && nms_suppresses(S.kbox[k], S.karea[k], b, area, P.thr)) { alive = false; break; }
                }
                S.box[t] = b;
                S.area[t] = area;
                S.cls[t] = cls;
                const u32 dead = __ballot_sync(0xffffffffu, !alive);
                if (lane == 0) S.removed[warp] = dead;
            }
            __syncthreads();
            LZ_TICK(4);
            // (b) bitmask: 256 rows x 8 words = 2048 (row, word) tasks, two per thread, word uniform per warp
#pragma unroll 1
            for (int q = 0; q < 2; ++q) {
                const int task = q * LZ_BLOCK + t;
                const int row = task & (LZ_CHUNK - 1), w = task >> 8;
                u32 bits = 0;
                const bool row_alive = row < m && !((S.removed[row >> 5] >> (row & 31)) & 1u);
                if (row_alive && w * 32 + 31 > row && w * 32 < m) {
                    const float4 b = S.box[row];
                    const float area = S.area[row];
                    const int cls = S.cls[row];
                    const int jlo = max(w * 32, row + 1), jhi = min(w * 32 + 32, m);
#pragma unroll 1
                    for (int j = jlo; j < jhi; ++j)
                        if (S.cls[j] == cls && nms_suppresses(b, area, S.box[j], S.area[j], P.thr)) bits |= 1u << (j & 31);
                }
                S.mask[row][w] = bits;
            }
            __syncthreads();
            LZ_TICK(5);
            // (c) warp-cooperative suppression scan, 32 rows at a time: the 32x32 diagonal block is
            // resolved in registers (shuffle-broadcast rows, warp-uniform ALU chain); the kept rows' mask
            // words are then OR-ed into the later removed-words by all 32 lanes (4 row subsets x 8 words).
            // Stops as soon as the image's max_det boxes are found: later rows are simply dropped.
            if (warp == 0) {
                // Parallel form of the greedy scan.  Within a 32-row group, K[i] = alive[i] & no kept j<i with
                // M[j][i]; iterating K <- f(K) from K = alive converges to the greedy answer (row i is final
                // after at most i sweeps; chains are short in practice).  Each sweep is one warp-wide OR
                // reduction (REDUX) of the kept rows' diagonal words; the kept rows are then OR-reduced into
                // the later removed-words the same way.  No serial 32-step chain, no divergent loops.
                u32 rw = lane < LZ_CHUNK / 32 ? S.removed[lane] : 0u;
                const int groups = (m + 31) >> 5;
                int room = P.max_det - kept;
#pragma unroll 1
                for (int g = 0; g < groups; ++g) {
                    if (room <= 0) {                               // the cut is behind us: drop the rest
                        if (lane >= g && lane < LZ_CHUNK / 32) rw = 0xffffffffu;
                        break;
                    }
                    const int row = g * 32 + lane;
                    const u32 rem0 = __shfl_sync(0xffffffffu, rw, g);           // removed by earlier groups / dead
                    const bool valid = row < m && !((rem0 >> lane) & 1u);
                    const u32 diag = valid ? S.mask[row][g] : 0u;
                    bool kp = valid;
#pragma unroll 1
                    for (int sweep = 0; sweep < 32; ++sweep) {
                        const u32 rem = __reduce_or_sync(0xffffffffu, kp ? diag : 0u);
                        const bool nk = valid && !((rem >> lane) & 1u);
                        const bool changed = nk != kp;
                        kp = nk;
                        if (!__any_sync(0xffffffffu, changed)) break;
                    }
                    const u32 keptm = __ballot_sync(0xffffffffu, kp);
                    room -= __popc(keptm);
                    if (lane == g) rw = ~keptm;                    // everything not kept in this group counts as removed
#pragma unroll
                    for (int w = 0; w < LZ_CHUNK / 32; ++w) {
                        if (w > g) {                               // warp-uniform
                            const u32 acc = __reduce_or_sync(0xffffffffu, kp ? S.mask[row][w] : 0u);
                            if (lane == w) rw |= acc;
                        }
                    }
                }
                if (lane < LZ_CHUNK / 32) S.removed[lane] = rw;
            }
            __syncthreads();
            LZ_TICK(6);
            // (d) append the kept boxes, in order, to the image's output list (first max_det only)
            if (t < LZ_CHUNK) {
                const bool kp = t < m && !((S.removed[warp] >> lane) & 1u);
                const u32 km = __ballot_sync(0xffffffffu, kp);
                if (lane == 0) S.wbase[warp + 1] = __popc(km);
            }
            __syncthreads();
            if (t == 0) {
                S.wbase[0] = 0;
                for (int w = 0; w < LZ_CHUNK / 32; ++w) S.wbase[w + 1] += S.wbase[w];
            }
            __syncthreads();
            if (t < LZ_CHUNK) {
                const bool kp = t < m && !((S.removed[warp] >> lane) & 1u);
                const u32 km = __ballot_sync(0xffffffffu, kp);
                const int pos = kept + S.wbase[warp] + __popc(km & ((1u << lane) - 1u));
                if (kp && pos < P.max_det) {
                    S.kbox[pos] = S.box[t];
                    S.karea[pos] = S.area[t];
                    S.kcls[pos] = S.cls[t];
                    S.kscore[pos] = ~(u32)(S.sel[c0 + t] >> 32);
                }
            }
            kept = min(P.max_det, kept + S.wbase[LZ_CHUNK / 32]);
            __syncthreads();
            LZ_TICK(7);
#ifdef RN_LAZY_TIMING
            ++n_chunks;
#endif
        }
        processed += nsel;
        last = T;
#ifdef RN_LAZY_TIMING
        ++n_rounds;
#endif
    }
    if (kept < P.max_det && processed < K && t == 0) atomicOr(P.status + 2, 1);   // general algorithm needed
    for (int i = t; i < kept; i += LZ_BLOCK) {
        const long long o = (long long)n * P.max_det + i;
        ((float4 *)P.out_boxes)[o] = finish_box(S.kbox[i], P.out_ratio, n, P.out_format);
        P.out_scores[o] = __uint_as_float(S.kscore[i]);
        P.out_labels[o] = (long long)S.kcls[i] + 1;            // models.py:230 labels + 1
    }
    if (t == 0) P.out_count[n] = kept;
#ifdef RN_LAZY_TIMING
    if (t == 0)
        printf("lazy img %d K %d rounds %d chunks %d kept %d | load %lld radix %lld gather %lld sort %lld decode+a %lld mask %lld resolve %lld append %lld\n",
               n, K, n_rounds, n_chunks, kept, tacc[0], tacc[1], tacc[2], tacc[3], tacc[4], tacc[5], tacc[6], tacc[7]);
#endif
#undef LZ_TICK
}

// ------------------------------------------------------------------------------------------- LAZY v2
// The common case of the lazy algorithm without a selection sort and without a serial suppression scan.
// One CTA per image, three ideas:
//  * the top of the global order comes from a COUNTING SORT on score bins: one histogram pass over the image's
//    candidate keys (2048 bins between the score threshold and 1.0), a block scan picks the largest prefix of bins
//    that fits LZ2_CAP candidates, one scatter pass groups that prefix by bin, and every key ranks itself inside its
//    (small) bin — the prefix is then in exact global order (score desc, class asc, anchor asc);
//  * greedy NMS is independent per class (a candidate's fate depends only on higher-ranked candidates of ITS class),
//    so the prefix is regrouped by class with a stable counting sort (per-32-chunk class counts + a column scan) and
//    the 32 warps of the CTA run the classes in parallel: 32 candidates at a time against the class's kept list, the
//    32x32 in-chunk conflicts resolved in registers by the REDUX fixed-point iteration of lazy_nms_kernel;
//  * the kept flags, indexed by global rank, are compacted by a block scan: the first max_det kept candidates ARE
//    the reference's output (models.py:193-240 with the documented tie rule).
// pre_nms_topk (extension): a candidate is eligible iff fewer than k candidates of its pyramid level rank above it — a
// second per-chunk count table, over the levels, gives that rank.
// Exactness is never traded: whatever this kernel cannot finish — more than LZ2_MAXC classes, a score
// bin with more than LZ2_BIN_MAX entries (mass ties), or fewer than max_det survivors while candidates remain beyond
// the prefix — is flagged in need_v1[n] and redone from scratch by lazy_nms_kernel (launched right behind, its CTAs
// exit at once for finished images), which in turn can hand over to the general algorithm.
constexpr int LZ2_BLOCK = 1024;
constexpr int LZ2_CAP = 4096;                 // candidates of the prefix
constexpr int LZ2_BINS = 2048;
constexpr int LZ2_MAXC = 128;                 // classes
constexpr int LZ2_CHUNKS = LZ2_CAP / 32;
constexpr int LZ2_BIN_MAX = 256;              // heaviest bin ranked in place (quadratic in the bin size)
#ifndef LZ2_WAVE_DEF
#define LZ2_WAVE_DEF 512
#endif
constexpr int LZ2_WAVE = LZ2_WAVE_DEF;        // ranks per lazy wave of the per-class NMS
#ifndef LZ2_FIRST_DEF
#define LZ2_FIRST_DEF 1024
#endif
constexpr int LZ2_FIRST = LZ2_FIRST_DEF;      // size aimed at by the first (short) prefix
constexpr int LZ2_LD = 4;                     // candidate keys in flight per thread in the two passes over the list

struct Lazy2Smem {
    u64 key[LZ2_CAP];                         // the prefix in global order (rank -> key)
    float4 box[LZ2_CAP];                      // class-grouped order (also the scatter target of the counting sort, as u64)
    float area[LZ2_CAP];
    unsigned short rank_of[LZ2_CAP];          // grouped position -> rank
    unsigned short pos_of[LZ2_CAP];           // rank -> grouped position
    unsigned short kidx[LZ2_CAP];             // per class segment: grouped positions of the kept boxes, in order
    unsigned char ok[LZ2_CAP];                // grouped: passes remove_small_boxes
    unsigned char kept[LZ2_CAP];              // by rank
    u32 hist[LZ2_BINS];                       // bin counts, then scatter cursors (= inclusive prefix when done)
    unsigned short ctab[LZ2_CHUNKS][LZ2_MAXC];   // class counts per 32-rank chunk, then exclusive prefix over the chunks
    int cls_base[LZ2_MAXC + 1];
    int cls_cnt[LZ2_MAXC], cls_cur[LZ2_MAXC], cls_kept[LZ2_MAXC];    // per class: members / members processed / boxes kept
    unsigned char cls_order[LZ2_MAXC];                  // classes by descending member count
    unsigned short ltab[LZ2_CHUNKS][RN_MAX_LEVELS];   // pre_nms_topk: level counts per chunk, then exclusive prefix
    int lvl_total[RN_MAX_LEVELS];
    int warp_tot[32];
    int cut_bin, first_bin, prefix, next_cls, heavy, total_kept;
};

// exclusive block scan of one int per thread (LZ2_BLOCK threads); returns the exclusive prefix, `total` = block sum
__device__ __forceinline__ int lz2_block_scan(int v, int *warp_tot, int &total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    __syncthreads();                          // warp_tot may still be read from a previous call
    if (lane == 31) warp_tot[warp] = x;
    __syncthreads();
    int w = warp_tot[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += y;
    }
    total = __shfl_sync(0xffffffffu, w, 31);
    const int before = warp ? __shfl_sync(0xffffffffu, w, warp - 1) : 0;
    return before + x - v;
}

__global__ void __launch_bounds__(LZ2_BLOCK, 1) lazy2_nms_kernel(const __grid_constant__ LazyParams P) {
    extern __shared__ __align__(16) unsigned char lz2_raw[];
    Lazy2Smem &S = *reinterpret_cast<Lazy2Smem *>(lz2_raw);
    const int n = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
#ifdef RN_LAZY_TIMING
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tprev = clock64();
    int n_waves = 0;
#define LZ2_TICK(id) do { __syncthreads(); if (t == 0) { long long now_ = clock64(); tacc[id] += now_ - tprev; tprev = now_; } } while (0)
#else
#define LZ2_TICK(id) do { } while (0)
#endif
    const u32 found = P.img_count[n];
    if (t == 0) {
        const unsigned long long scaled = (unsigned long long)found * (unsigned long long)P.N;
        atomicMax(P.status + 0, (int)min(scaled, 0x7fffffffULL));
    }
    const int K = (int)min(found, P.cap_n);
    if (K == 0) {
        if (t == 0) { P.out_count[n] = 0; P.need_v1[n] = 0; }
        return;
    }
    const u64 *cand = P.cand_key + (size_t)n * P.cap_n;
    const u32 hi0 = 0xC07FFFFFu;              // ~bits(1.0f): the smallest possible upper key half
    const int shift = P.bin_shift;
    auto bin_of = [&](u64 k) -> int {
        const u32 hi = (u32)(k >> 32);
        return hi <= hi0 ? 0 : (int)min((hi - hi0) >> shift, (u32)(LZ2_BINS - 1));
    };
    const int C = P.C;
    const u32 A32 = (u32)P.A;
    const bool topk = P.topk > 0;
    const float imh = (float)P.im_hw[2 * n], imw = (float)P.im_hw[2 * n + 1];
    const long long anc_row = (long long)n * P.anchor_stride;
    u64 *tmp = reinterpret_cast<u64 *>(S.box);

    // Two attempts at most: a SHORT prefix (the first bins holding >= LZ2_FIRST candidates — the common case needs a
    // few hundred ranks for max_det survivors), then the largest prefix that fits LZ2_CAP.
    int Pn = 0, total_kept = 0;
#pragma unroll 1
    for (int attempt = 0; attempt < 2; ++attempt) {
        // ---- (1) histogram of the score bins (LZ2_LD keys in flight per thread) ----
        for (int i = t; i < LZ2_BINS; i += LZ2_BLOCK) S.hist[i] = 0;
        for (int i = t; i < LZ2_CAP; i += LZ2_BLOCK) S.kept[i] = 0;
        if (t == 0) { S.cut_bin = -1; S.first_bin = LZ2_BINS; S.prefix = 0; S.next_cls = 0; S.heavy = 0; S.total_kept = 0; }
        __syncthreads();
        for (int i0 = 0; i0 < K; i0 += LZ2_BLOCK * LZ2_LD) {
            u64 k[LZ2_LD];
#pragma unroll
            for (int u = 0; u < LZ2_LD; ++u) {
                const int i = i0 + u * LZ2_BLOCK + t;
                k[u] = i < K ? __ldg(cand + i) : ~0ULL;
            }
#pragma unroll
            for (int u = 0; u < LZ2_LD; ++u)
                if (i0 + u * LZ2_BLOCK + t < K) atomicAdd(&S.hist[bin_of(k[u])], 1u);
        }
        __syncthreads();
        LZ2_TICK(0);
        // ---- (2) the cut: hist becomes the scatter cursor (exclusive prefix) ----
        {
            const int c0 = (int)S.hist[2 * t], c1 = (int)S.hist[2 * t + 1];
            int total;
            const int ex = lz2_block_scan(c0 + c1, S.warp_tot, total);
            const int in0 = ex + c0, in1 = in0 + c1;           // inclusive prefix after bin 2t / 2t+1 (monotone)
            if (in1 <= LZ2_CAP) atomicMax(&S.cut_bin, 2 * t + 1);          // largest bin whose prefix fits
            else if (in0 <= LZ2_CAP) atomicMax(&S.cut_bin, 2 * t);
            if (in0 >= LZ2_FIRST) atomicMin(&S.first_bin, 2 * t);          // smallest bin whose prefix reaches LZ2_FIRST
            else if (in1 >= LZ2_FIRST) atomicMin(&S.first_bin, 2 * t + 1);
            __syncthreads();
            const int cutb = (attempt == 0 && S.first_bin <= S.cut_bin) ? S.first_bin : S.cut_bin;
            if ((c0 > LZ2_BIN_MAX && 2 * t <= cutb) || (c1 > LZ2_BIN_MAX && 2 * t + 1 <= cutb)) S.heavy = 1;
            if (cutb == 2 * t) S.prefix = in0;
            if (cutb == 2 * t + 1) S.prefix = in1;
            __syncthreads();
            if (t == 0) S.cut_bin = cutb;
            S.hist[2 * t] = (u32)ex;
            S.hist[2 * t + 1] = (u32)in0;
        }
        __syncthreads();
        const int cut = S.cut_bin;
        Pn = S.prefix;
        if (cut < 0 || S.heavy || Pn == 0) {      // first bin alone overflows / mass ties: not for this kernel
            if (t == 0) P.need_v1[n] = 1;
            return;
        }
        LZ2_TICK(1);
        // ---- (3) scatter the prefix by bin, then every key ranks itself inside its bin ----
        for (int i0 = 0; i0 < K; i0 += LZ2_BLOCK * LZ2_LD) {
            u64 k[LZ2_LD];
#pragma unroll
            for (int u = 0; u < LZ2_LD; ++u) {
                const int i = i0 + u * LZ2_BLOCK + t;
                k[u] = i < K ? __ldg(cand + i) : ~0ULL;
            }
#pragma unroll
            for (int u = 0; u < LZ2_LD; ++u) {
                if (i0 + u * LZ2_BLOCK + t < K) {
                    const int b = bin_of(k[u]);
                    if (b <= cut) tmp[atomicAdd(&S.hist[b], 1u)] = k[u];
                }
            }
        }
        __syncthreads();
        LZ2_TICK(2);
        for (int i = t; i < Pn; i += LZ2_BLOCK) {
            const u64 k = tmp[i];
            const int b = bin_of(k);
            const int lo = b ? (int)S.hist[b - 1] : 0, hi = (int)S.hist[b];     // cursors now sit at the bins' ends
            int r = lo;
            for (int q = lo; q < hi; ++q) r += tmp[q] < k;
            S.key[r] = k;
        }
        __syncthreads();
        LZ2_TICK(3);
        // ---- (4) class of every ranked candidate; stable regrouping by class; decode into the grouped arrays ----
        const int nchunks = (Pn + 31) >> 5;
        for (int i = t; i < nchunks * LZ2_MAXC; i += LZ2_BLOCK) S.ctab[i / LZ2_MAXC][i % LZ2_MAXC] = 0;
        for (int i = t; i < nchunks * RN_MAX_LEVELS; i += LZ2_BLOCK) S.ltab[i / RN_MAX_LEVELS][i % RN_MAX_LEVELS] = 0;
        __syncthreads();
        int my_lvl[LZ2_CHUNKS / 32], my_lric[LZ2_CHUNKS / 32];
        int my_cls[LZ2_CHUNKS / 32], my_ric[LZ2_CHUNKS / 32];
#pragma unroll
        for (int q = 0; q < LZ2_CHUNKS / 32; ++q) {
            const int chunk = warp + 32 * q, r = chunk * 32 + lane;
            my_cls[q] = -1;
            my_ric[q] = 0;
            my_lvl[q] = -1;
            my_lric[q] = 0;
            if (chunk < nchunks) {                                  // warp-uniform
                const int cls = r < Pn ? (int)((u32)S.key[r] / A32) : -1;
                const u32 peers = __match_any_sync(0xffffffffu, cls);
                my_cls[q] = cls;
                my_ric[q] = __popc(peers & ((1u << lane) - 1u));
                if (cls >= 0 && lane == __ffs(peers) - 1) S.ctab[chunk][cls] = (unsigned short)__popc(peers);
                if (topk) {                                         // rank of the candidate inside its pyramid level
                    int lvl = -1;
                    if (cls >= 0) {
                        const long long anchor = (long long)((u32)S.key[r] - (u32)cls * A32);
                        lvl = 0;
                        for (int l = 1; l < P.nlev; ++l) lvl += anchor >= P.lvl_off[l];
                    }
                    const u32 lp = __match_any_sync(0xffffffffu, lvl);
                    my_lvl[q] = lvl;
                    my_lric[q] = __popc(lp & ((1u << lane) - 1u));
                    if (lvl >= 0 && lane == __ffs(lp) - 1) S.ltab[chunk][lvl] = (unsigned short)__popc(lp);
                }
            }
        }
        __syncthreads();
        int ccount = 0;
        if (t < C) {                                                // column scan over the chunks, 8 loads in flight
            for (int ch0 = 0; ch0 < nchunks; ch0 += 8) {
                int v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = ch0 + u < nchunks ? S.ctab[ch0 + u][t] : 0;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (ch0 + u < nchunks) S.ctab[ch0 + u][t] = (unsigned short)ccount;
                    ccount += v[u];
                }
            }
        }
        if (topk && t >= LZ2_BLOCK - RN_MAX_LEVELS) {               // the last warp's lanes scan the level columns
            const int l = t - (LZ2_BLOCK - RN_MAX_LEVELS);
            int run = 0;
            for (int ch = 0; ch < nchunks; ++ch) {
                const int v = S.ltab[ch][l];
                S.ltab[ch][l] = (unsigned short)run;
                run += v;
            }
            S.lvl_total[l] = run;
        }
        {
            int total;
            const int ex = lz2_block_scan(t < C ? ccount : 0, S.warp_tot, total);
            if (t < C) { S.cls_base[t] = ex; S.cls_cnt[t] = ccount; S.cls_cur[t] = 0; S.cls_kept[t] = 0; }
            if (t == 0) S.cls_base[C] = total;
        }
        __syncthreads();
        if (t < C) {                                                // heaviest classes first (longest-processing-time order)
            const int mine = S.cls_cnt[t];
            int rk = 0;
            for (int c = 0; c < C; ++c) {
                const int o = S.cls_cnt[c];
                rk += (o > mine) || (o == mine && c < t);
            }
            S.cls_order[rk] = (unsigned char)t;
        }
        LZ2_TICK(4);
#pragma unroll
        for (int q = 0; q < LZ2_CHUNKS / 32; ++q) {
            const int chunk = warp + 32 * q, r = chunk * 32 + lane, cls = my_cls[q];
            if (cls >= 0) {
                const int g = S.cls_base[cls] + S.ctab[chunk][cls] + my_ric[q];
                const long long anchor = (long long)((u32)S.key[r] - (u32)cls * A32);
                const float4 b = decode_clip(P.box, n, P.A, anchor, P.anchors, anc_row + anchor, P.wts, imw, imh);
                S.box[g] = b;
                S.area[g] = nms_area(b);
                bool ok = (__fsub_rn(b.z, b.x) >= 0.01f) && (__fsub_rn(b.w, b.y) >= 0.01f);    // remove_small_boxes, models.py:203
                if (topk) ok = ok && (S.ltab[chunk][my_lvl[q]] + my_lric[q] < P.topk);         // only the level's top-k scores
                S.ok[g] = ok;
                S.rank_of[g] = (unsigned short)r;
                S.pos_of[r] = (unsigned short)g;
            }
        }
        __syncthreads();
        LZ2_TICK(5);
        // ---- (5) greedy NMS, one warp per class at a time, LAZILY in waves of LZ2_WAVE ranks: every class advances
        // through its (rank-sorted) members below the wave's end; as soon as max_det candidates are kept among the
        // ranks seen so far, the answer is complete (a kept flag only depends on higher-ranked members of the class) ----
        for (int wave_end = LZ2_WAVE; ; wave_end += LZ2_WAVE) {
            for (;;) {
                int ci = 0;
                if (lane == 0) ci = atomicAdd(&S.next_cls, 1);
                ci = __shfl_sync(0xffffffffu, ci, 0);
                if (ci >= C) break;
                const int c = S.cls_order[ci];
                const int base = S.cls_base[c], nc = S.cls_cnt[c];
                int s0 = S.cls_cur[c], kc = S.cls_kept[c];
                int added = 0;
                while (s0 < nc && kc < P.max_det) {
                    const int j = s0 + lane;
                    const bool valid = j < nc && (int)S.rank_of[base + j] < wave_end;      // ranks ascend inside a class
                    const u32 vm = __ballot_sync(0xffffffffu, valid);
                    if (vm == 0) break;
                    const float4 b = valid ? S.box[base + j] : make_float4(0.f, 0.f, 0.f, 0.f);
                    const float ar = valid ? S.area[base + j] : 0.f;
                    bool alive = valid && S.ok[base + j];
                    for (int k = 0; k < kc; ++k) {                    // boxes kept earlier in this class
                        const int kp = S.kidx[base + k];
                        if (alive && nms_suppresses(S.box[kp], S.area[kp], b, ar, P.thr)) alive = false;
                    }
                    const u32 am = __ballot_sync(0xffffffffu, alive);
                    u32 diag = 0;                                     // later boxes of this chunk that box `lane` suppresses
                    if (am & (am - 1)) {                              // at least two alive
                        u32 rest = am & (am - 1);                     // the first alive box is suppressed by nobody
                        while (rest) {
                            const int jj = __ffs(rest) - 1;
                            rest &= rest - 1;
                            if (alive && jj > lane && nms_suppresses(b, ar, S.box[base + s0 + jj], S.area[base + s0 + jj], P.thr))
                                diag |= 1u << jj;
                        }
                    }
                    bool kp = alive;
#pragma unroll 1
                    for (int sweep = 0; sweep < 32; ++sweep) {        // fixed point of kept = alive & no kept earlier conflict
                        const u32 rem = __reduce_or_sync(0xffffffffu, kp ? diag : 0u);
                        const bool nk = alive && !((rem >> lane) & 1u);
                        const bool changed = nk != kp;
                        kp = nk;
                        if (!__any_sync(0xffffffffu, changed)) break;
                    }
                    const u32 km = __ballot_sync(0xffffffffu, kp);
                    if (kp) {
                        S.kidx[base + kc + __popc(km & ((1u << lane) - 1u))] = (unsigned short)(base + j);
                        S.kept[S.rank_of[base + j]] = 1;
                    }
                    kc += __popc(km);
                    added += __popc(km);
                    s0 += __popc(vm);                                 // valid lanes form a prefix of the chunk
                    __syncwarp();
                    if (__popc(vm) < 32) break;                       // the rest of the class lies beyond this wave
                }
                __syncwarp();                                         // every lane has read the class state lane 0 rewrites
                if (lane == 0) {
                    S.cls_cur[c] = s0;
                    S.cls_kept[c] = kc;
                    if (added) atomicAdd(&S.total_kept, added);
                }
            }
            __syncthreads();
#ifdef RN_LAZY_TIMING
            ++n_waves;
#endif
            if (S.total_kept >= P.max_det || wave_end >= Pn) break;
            __syncthreads();
            if (t == 0) S.next_cls = 0;
            __syncthreads();
        }
        LZ2_TICK(6);
        total_kept = S.total_kept;
        bool more = total_kept < P.max_det && Pn < K;         // survivors may hide beyond the prefix ...
        if (more && topk) {                                   // ... unless every level already saw its k candidates
            bool full = true;
            for (int l = 0; l < P.nlev; ++l) full = full && S.lvl_total[l] >= P.topk;
            more = !full;
        }
        if (!more) break;
        if (attempt == 1 || Pn >= LZ2_CAP) {                  // nothing larger fits: lazy_nms_kernel takes the image
            if (t == 0) P.need_v1[n] = 1;
            return;
        }
        __syncthreads();                                      // second attempt: everything is rebuilt with the long prefix
    }
    // ---- (6) the first max_det kept candidates in global order ----
    {
        const int r0 = 4 * t;                                     // LZ2_CAP = 4 * LZ2_BLOCK
        int f[4], cnt = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) { f[q] = (r0 + q < Pn) ? S.kept[r0 + q] : 0; cnt += f[q]; }
        int total;
        int pos = lz2_block_scan(cnt, S.warp_tot, total);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (f[q]) {
                if (pos < P.max_det) {
                    const u64 k = S.key[r0 + q];
                    const long long o = (long long)n * P.max_det + pos;
                    ((float4 *)P.out_boxes)[o] = finish_box(S.box[S.pos_of[r0 + q]], P.out_ratio, n, P.out_format);
                    P.out_scores[o] = __uint_as_float(~(u32)(k >> 32));
                    P.out_labels[o] = (long long)((u32)k / A32) + 1;          // models.py:230 labels + 1
                }
                ++pos;
            }
        }
        if (t == 0) {
            P.out_count[n] = min(total, P.max_det);
            P.need_v1[n] = 0;
        }
    }
#ifdef RN_LAZY_TIMING
    LZ2_TICK(7);
    if (t == 0)
        printf("lazy2 img %d K %d prefix %d waves %d | hist %lld cut %lld scatter %lld rank %lld classes %lld decode %lld nms %lld out %lld\n",
               n, K, Pn, n_waves, tacc[0], tacc[1], tacc[2], tacc[3], tacc[4], tacc[5], tacc[6], tacc[7]);
#endif
#undef LZ2_TICK
}

// ------------------------------------------------------------------------------------------- K5
struct TopkParams {
    const int *seg_off;
    const int *kept_count;
    const u64 *kept_key;
    const float4 *kept_box;
    int C;
    int max_det;
    float *out_boxes;
    float *out_scores;
    long long *out_labels;
    int *out_count;
    const float *out_ratio;  // [N,2] or null (row N2)
    int out_format;          // 0 xyxy, 1 xywh (row N4)
};

constexpr int TOPK_BLOCK = 1024;
constexpr int TOPK_CACHE = 16384;     // inverted-score words cached in shared memory

// One CTA per image.  The kept list of every class is sorted by (score desc, anchor asc), so only
// the first max_det entries of each class can reach the image's top max_det: the kernel works on the
// "flattened" sequence e = 0..total-1 of those heads (class-major), whose order among equal scores
// is exactly the documented tie rule (class asc, anchor asc).
__global__ void __launch_bounds__(TOPK_BLOCK) image_topk_kernel(const TopkParams P) {
    extern __shared__ u32 s_dyn[];                  // [C+1] class prefix | [TOPK_CACHE] cached keys
    __shared__ u32 s_hist[256];
    __shared__ u32 s_prefix, s_mask_bits, s_need;
    __shared__ int s_carry, s_nsel;
    __shared__ int s_warp[TOPK_BLOCK / 32];
    __shared__ u64 s_sel_hi[MAX_DET_CAP];           // (inverted score << 32) | class
    __shared__ u32 s_sel_lo[MAX_DET_CAP];           // anchor
    __shared__ int s_sel_pos[MAX_DET_CAP];          // index into the kept arrays

    const int n = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int C = P.C, seg0 = n * C;
    int *s_pref = (int *)s_dyn;
    u32 *s_hi = s_dyn + (C + 1);

    // ---- exclusive scan of m_c = min(kept_count[c], max_det) over the classes ----
    if (t == 0) { s_carry = 0; s_nsel = 0; }
    __syncthreads();
    for (int base = 0; base < C; base += TOPK_BLOCK) {
        const int c = base + t;
        const int v = c < C ? min(P.kept_count[seg0 + c], P.max_det) : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += y;
            }
            s_warp[lane] = w;
        }
        __syncthreads();
        const int excl = s_carry + (warp ? s_warp[warp - 1] : 0) + x - v;
        if (c < C) s_pref[c] = excl;
        __syncthreads();
        if (t == TOPK_BLOCK - 1) s_carry = excl + v;
        __syncthreads();
    }
    const int total = s_carry;
    if (t == 0) s_pref[C] = total;
    const int want = min(total, P.max_det);
    if (t == 0) P.out_count[n] = want;
    if (want == 0) return;
    __syncthreads();

    // flattened index -> position in the kept arrays (binary search over the class prefix)
    auto locate = [&](int e, int &c) -> int {
        int lo = 0, hi = C;                          // largest c with s_pref[c] <= e
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (s_pref[mid] <= e) lo = mid; else hi = mid;
        }
        c = lo;
        return P.seg_off[seg0 + lo] + (e - s_pref[lo]);
    };
    const bool cached = total <= TOPK_CACHE;
    if (cached) {
        for (int e = t; e < total; e += TOPK_BLOCK) {
            int c;
            s_hi[e] = (u32)(P.kept_key[locate(e, c)] >> 32);
        }
        __syncthreads();
    }
    auto get_hi = [&](int e) -> u32 {
        if (cached) return s_hi[e];
        int c;
        return (u32)(P.kept_key[locate(e, c)] >> 32);
    };

    // ---- radix select (4 x 8 bits, MSB first) of the want-th smallest inverted score ----
    const bool select_all = total <= P.max_det;
    u32 thr_hi = 0xffffffffu;
    int need_ties = 0;
    if (!select_all) {
        if (t == 0) { s_prefix = 0; s_mask_bits = 0; s_need = (u32)want; }
        __syncthreads();
        for (int shift = 24; shift >= 0; shift -= 8) {
            if (t < 256) s_hist[t] = 0;
            __syncthreads();
            const u32 prefix = s_prefix, mbits = s_mask_bits;
            for (int e0 = warp * 32; e0 < total; e0 += TOPK_BLOCK) {     // warp-aggregated (see lazy_nms_kernel)
                const int e = e0 + lane;
                u32 digit = 256u;
                if (e < total) {
                    const u32 hi = get_hi(e);
                    if ((hi & mbits) == prefix) digit = (hi >> shift) & 255u;
                }
                const u32 peers = __match_any_sync(0xffffffffu, digit);
                if (digit < 256u && lane == __ffs(peers) - 1) atomicAdd(&s_hist[digit], (u32)__popc(peers));
            }
            __syncthreads();
            if (warp == 0) {                          // find the bucket holding the need-th entry
                u32 cnt[8], sum = 0;
#pragma unroll
                for (int i = 0; i < 8; ++i) { cnt[i] = s_hist[lane * 8 + i]; sum += cnt[i]; }
                u32 incl = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const u32 y = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += y;
                }
                const u32 need = s_need;
                const u32 excl = incl - sum;
                __syncwarp();                          // every lane has read s_need before one lane rewrites it
                if (excl < need && need <= incl) {    // exactly one lane
                    u32 acc = excl;
                    int b = 0;
                    for (; b < 8; ++b) {
                        if (acc + cnt[b] >= need) break;
                        acc += cnt[b];
                    }
                    s_need = need - acc;
                    s_prefix = prefix | ((u32)(lane * 8 + b) << shift);
                    s_mask_bits = mbits | (255u << shift);
                }
            }
            __syncthreads();
        }
        thr_hi = s_prefix;
        need_ties = (int)s_need;                      // entries with hi == thr_hi to take, in flattened order
    }

    // ---- collect the winners; ties at the threshold are ranked by flattened order ----
    if (t == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < total; base += TOPK_BLOCK) {
        const int e = base + t;
        u32 hi = 0;
        bool is_lt = false, is_tie = false;
        if (e < total) {
            hi = get_hi(e);
            is_lt = select_all || hi < thr_hi;
            is_tie = !select_all && hi == thr_hi;
        }
        bool take = is_lt;
        if (!select_all) {
            const u32 tm = __ballot_sync(0xffffffffu, is_tie);
            if (lane == 0) s_warp[warp] = __popc(tm);
            __syncthreads();
            if (warp == 0) {
                int w = s_warp[lane];
                const int own = w;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int y = __shfl_up_sync(0xffffffffu, w, o);
                    if (lane >= o) w += y;
                }
                s_warp[lane] = w - own;               // exclusive
                if (lane == 31) s_hist[0] = (u32)w;   // chunk total
            }
            __syncthreads();
            if (is_tie) take = s_carry + s_warp[warp] + __popc(tm & ((1u << lane) - 1u)) < need_ties;
            __syncthreads();
            if (t == 0) s_carry += (int)s_hist[0];
            __syncthreads();
        }
        if (take) {
            int c;
            const int pos = locate(e, c);
            const int slot = atomicAdd(&s_nsel, 1);
            if (slot < MAX_DET_CAP) {
                s_sel_hi[slot] = ((u64)hi << 32) | (u64)(u32)c;
                s_sel_lo[slot] = (u32)P.kept_key[pos];
                s_sel_pos[slot] = pos;
            }
        }
    }
    __syncthreads();
    const int nsel = min(s_nsel, want);

    // ---- order the winners (score desc, class asc, anchor asc) by rank counting; nsel <= 1024 ----
    for (int i = t; i < nsel; i += TOPK_BLOCK) {
        const u64 hi = s_sel_hi[i];
        const u32 lo = s_sel_lo[i];
        int rank = 0;
        for (int j = 0; j < nsel; ++j) {
            const u64 hj = s_sel_hi[j];
            rank += (hj < hi) || (hj == hi && s_sel_lo[j] < lo);
        }
        const float4 b = P.kept_box[s_sel_pos[i]];
        const long long o = (long long)n * P.max_det + rank;
        ((float4 *)P.out_boxes)[o] = finish_box(b, P.out_ratio, n, P.out_format);
        P.out_scores[o] = __uint_as_float(~(u32)(hi >> 32));
        P.out_labels[o] = (long long)(u32)hi + 1;          // models.py:230 labels + 1
    }
}

}  // namespace

extern "C" size_t rn_postprocess_workspace_bytes(int N, int64_t A, int C, int64_t cand_capacity, int max_det) {
    (void)A; (void)max_det;
    if (N <= 0 || C <= 0 || cand_capacity < 0) return 0;
    return carve(nullptr, N, C, cand_capacity).total_bytes;
}

// Everything after the streaming filter (shared by the [N,A,C] and the per-level entry points).
static int pp_tail(const PPWorkspace &w, const FilterParams &F, const BoxSource &box, const float *anchors,
                   int64_t anchor_image_stride, const int32_t *im_hw, int N, int64_t A, int C, float score_thr, double nms_thr,
                   int max_det, int pre_nms_topk, const int64_t *level_off_host, int num_levels, bool lazy,
                   int64_t cand_capacity, float *out_boxes, float *out_scores, int64_t *out_labels,
                   int32_t *out_count, int32_t *out_status, const float *out_ratio_hw, int out_format, cudaStream_t s) {
    // round the double threshold DOWN to fp32: (double)ovr > thr  <=>  ovr > thr_f  for every fp32 ovr
    float thr_f = (float)nms_thr;
    if ((double)thr_f > nms_thr) thr_f = nextafterf(thr_f, -INFINITY);

    if (lazy) {
        LazyParams Z;
        Z.box = box; Z.anchors = (const float4 *)anchors; Z.im_hw = im_hw; Z.A = A;
        Z.anchor_stride = anchor_image_stride; Z.C = C; Z.N = N; Z.thr = thr_f; Z.wts = F.wts; Z.max_det = max_det;
        Z.cand_key = w.pool_key; Z.img_count = w.img_count; Z.cap_n = F.cap_n; Z.out_boxes = out_boxes;
        Z.out_scores = out_scores; Z.out_labels = (long long *)out_labels; Z.out_count = out_count; Z.status = out_status;
        Z.topk = pre_nms_topk; Z.nlev = pre_nms_topk ? num_levels : 0;
        Z.out_ratio = out_ratio_hw; Z.out_format = out_format;
        for (int l = 0; l <= RN_MAX_LEVELS; ++l)
            Z.lvl_off[l] = (pre_nms_topk && l <= num_levels) ? (long long)level_off_host[l] : (long long)A;
        Z.need_v1 = nullptr;
        Z.bin_shift = 0;
        Z.capacity = (int)min((long long)F.cap_n * N, 0x7fffffffLL);
        if (C <= LZ2_MAXC && w.need_v1) {
            // score bins of lazy2_nms_kernel: LZ2_BINS equal-width bins (in float bit patterns) between thr and 1.0
            const float thr_pos = score_thr > 0.0f ? score_thr : 0.0f;
            uint32_t lo_bits, one_bits;
            const float one = 1.0f;
            memcpy(&lo_bits, &thr_pos, 4);
            memcpy(&one_bits, &one, 4);
            const uint32_t range = one_bits > lo_bits ? one_bits - lo_bits : 1u;
            int shift = 0;
            while (((range) >> shift) >= (uint32_t)LZ2_BINS) ++shift;
            Z.bin_shift = shift;
            Z.need_v1 = w.need_v1;
            static_assert(sizeof(Lazy2Smem) <= 227 * 1024, "lazy2 shared memory");
            cudaFuncSetAttribute(lazy2_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Lazy2Smem));
            lazy2_nms_kernel<<<N, LZ2_BLOCK, sizeof(Lazy2Smem), s>>>(Z);
            RN_CHECK_LAUNCH("rn_postprocess/lazy2_nms");
        }
        const size_t smem = ((sizeof(LazySmem) + 15) & ~(size_t)15) + (size_t)LZ_CACHE * sizeof(u64);
        cudaFuncSetAttribute(lazy_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        lazy_nms_kernel<<<N, LZ_BLOCK, smem, s>>>(Z);
        RN_CHECK_LAUNCH("rn_postprocess/lazy_nms");
        return 0;
    }

    const int S = N * C;
    segment_scan_kernel<<<1, 1024, 0, s>>>(w.seg_count, S, w.seg_off, w.cursor, w.pool_count, (u32)cand_capacity, out_status);
    RN_CHECK_LAUNCH("rn_postprocess/segment_scan");
    scatter_kernel<<<RN_SM_COUNT_B200 * 4, 256, 0, s>>>(w.pool_key, w.pool_seg, w.pool_count, (u32)cand_capacity, w.cursor,
                                                        w.sorted_key);
    RN_CHECK_LAUNCH("rn_postprocess/scatter");

    NmsParams M;
    M.box = box; M.anchors = (const float4 *)anchors; M.im_hw = im_hw; M.boxes = nullptr;
    M.keep_flags = nullptr; M.A = A; M.anchor_stride = anchor_image_stride; M.C = C; M.thr = thr_f; M.wts = F.wts; M.seg_off = w.seg_off;
    M.sorted_key = w.sorted_key; M.kept_key = w.kept_key; M.kept_box = w.kept_box; M.kept_count = w.kept_count;
    nms_kernel<false><<<S, NMS_CHUNK, 0, s>>>(M);
    RN_CHECK_LAUNCH("rn_postprocess/nms");

    TopkParams T;
    T.seg_off = w.seg_off; T.kept_count = w.kept_count; T.kept_key = w.kept_key; T.kept_box = w.kept_box;
    T.C = C; T.max_det = max_det; T.out_boxes = out_boxes; T.out_scores = out_scores;
    T.out_ratio = out_ratio_hw; T.out_format = out_format;
    T.out_labels = (long long *)out_labels; T.out_count = out_count;
    const size_t topk_smem = ((size_t)C + 1 + TOPK_CACHE) * sizeof(u32);
    if (topk_smem > 48 * 1024)
        cudaFuncSetAttribute(image_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)topk_smem);
    image_topk_kernel<<<N, TOPK_BLOCK, topk_smem, s>>>(T);
    RN_CHECK_LAUNCH("rn_postprocess/image_topk");
    return 0;
}

// ---- pieces of the LAZY post-processing reused by rn_train_detect (loss.cu), whose loss kernel does the score
// filtering itself while it streams the logits: (a) validation, zeroing and carving of the workspace -> the candidate
// lists the filter appends to; (b) everything after the filter. ----
#include "pp_internal.cuh"

static float guard_logit(float thr) {
    // every x <= x_lo has sigmoid(x) <= thr with a 1e-3 margin (SFU/libm error is ~1e-7)
    if (!(thr > 0.0f)) return -INFINITY;
    if (thr >= 1.0f) return INFINITY;
    const double xt = log((double)thr / (1.0 - (double)thr));
    return (float)(xt - 1e-3 * fmax(1.0, fabs(xt)));
}

int rnpp::lazy_begin(int N, int64_t A, int C, float score_thr, int max_det, int pre_nms_topk, const int64_t *level_off_host,
                     int num_levels, int64_t cand_capacity, int32_t *out_status, void *workspace, size_t workspace_bytes,
                     cudaStream_t s, rnpp::LazySink *sink, bool zero) {
    RN_CHECK_ARG(out_status && workspace && sink, RN_E_BADARG, "rn_train_detect: null pointer");
    RN_CHECK_ARG(N > 0 && A > 0 && C > 0 && N <= 65535 && C <= 8192, RN_E_BADARG, "rn_train_detect: bad N/A/C");
    RN_CHECK_ARG(max_det >= 1 && max_det <= MAX_DET_CAP, RN_E_TOOLARGE, "rn_train_detect: max_det=%d outside [1,%d]", max_det, MAX_DET_CAP);
    RN_CHECK_ARG(cand_capacity >= 1 && cand_capacity < 0x7fffffffLL, RN_E_BADARG, "rn_train_detect: bad cand_capacity");
    RN_CHECK_ARG(pre_nms_topk >= 0 && (pre_nms_topk == 0 || (level_off_host && num_levels >= 1 && num_levels <= RN_MAX_LEVELS)),
                 RN_E_BADARG, "rn_train_detect: pre_nms_topk needs 1..%d level offsets", RN_MAX_LEVELS);
    RN_CHECK_ARG((unsigned long long)A * (unsigned long long)C < (1ULL << 32), RN_E_TOOLARGE,
                 "rn_train_detect: the lazy algorithm needs A*C < 2^32");
    PPWorkspace w = carve(workspace, N, C, cand_capacity);
    RN_CHECK_ARG(workspace_bytes >= w.total_bytes, RN_E_WORKSPACE, "rn_train_detect: post-processing workspace too small (%zu < %zu)",
                 workspace_bytes, w.total_bytes);
    if (zero) {
        cudaError_t e = cudaMemsetAsync(workspace, 0, w.zero_bytes, s);
        if (e == cudaSuccess) e = cudaMemsetAsync(out_status, 0, 4 * sizeof(int32_t), s);
        if (e != cudaSuccess) { rn_set_error("rn_train_detect: memset failed: %s", cudaGetErrorString(e)); return (int)e; }
    }
    sink->img_count = w.img_count;
    sink->pool_key = w.pool_key;
    sink->cap_n = (u32)max((int64_t)1, cand_capacity / N);
    sink->x_lo = guard_logit(score_thr);
    sink->thr = score_thr;
    return 0;
}

int rnpp::lazy_end(const float *bbox, const float *anchors, int64_t anchor_image_stride, const int32_t *im_hw, int N, int64_t A,
                   int C, float score_thr, double nms_thr, int max_det, const float *weights_host, int pre_nms_topk,
                   const int64_t *level_off_host, int num_levels, int64_t cand_capacity, float *out_boxes, float *out_scores,
                   int64_t *out_labels, int32_t *out_count, int32_t *out_status, void *workspace, cudaStream_t s,
                   const float *out_ratio_hw, int out_format) {
    RN_CHECK_ARG(bbox && anchors && im_hw && weights_host && out_boxes && out_scores && out_labels && out_count, RN_E_BADARG,
                 "rn_train_detect: null pointer");
    RN_CHECK_ARG(out_format == 0 || out_format == 1, RN_E_BADARG, "rn_train_detect: out_format must be 0 (xyxy) or 1 (xywh)");
    PPWorkspace w = carve(workspace, N, C, cand_capacity);
    FilterParams F;
    memset(&F, 0, sizeof(F));
    F.cap = (u32)cand_capacity;
    F.cap_n = (u32)max((int64_t)1, cand_capacity / N);
    F.w = w;
    F.wts = make_float4(weights_host[0], weights_host[1], weights_host[2], weights_host[3]);
    BoxSource box;
    memset(&box, 0, sizeof(box));
    box.nac = (const float4 *)bbox;
    return pp_tail(w, F, box, anchors, anchor_image_stride, im_hw, N, A, C, score_thr, nms_thr, max_det, pre_nms_topk,
                   level_off_host, num_levels, true, cand_capacity, out_boxes, out_scores, out_labels, out_count, out_status,
                   out_ratio_hw, out_format, s);
}

extern "C" int rn_postprocess(const float *logits, const float *bbox, const float *anchors,
                              int64_t anchor_image_stride, const int32_t *im_hw, int N,
                              int64_t A, int C, float score_thr, double nms_thr, int max_det,
                              const float *weights_host, int pre_nms_topk, const int64_t *level_off_host,
                              int num_levels, int algo, int64_t cand_capacity, float *out_boxes, float *out_scores,
                              int64_t *out_labels, int32_t *out_count, int32_t *out_status, void *workspace,
                              size_t workspace_bytes, rn_stream_t stream, const float *out_ratio_hw, int out_format) {
    RN_CHECK_ARG(out_format == 0 || out_format == 1, RN_E_BADARG, "rn_postprocess: out_format must be 0 (xyxy) or 1 (xywh)");
    RN_CHECK_ARG(logits && bbox && anchors && im_hw && weights_host && out_boxes && out_scores && out_labels &&
                     out_count && out_status && workspace,
                 RN_E_BADARG, "rn_postprocess: null pointer");
    RN_CHECK_ARG(N > 0 && A > 0 && C > 0, RN_E_BADARG, "rn_postprocess: N, A, C must be positive");
    RN_CHECK_ARG(N <= 65535, RN_E_TOOLARGE, "rn_postprocess: N=%d exceeds 65535", N);
    RN_CHECK_ARG(A < (1LL << 32), RN_E_TOOLARGE, "rn_postprocess: A exceeds 2^32");
    RN_CHECK_ARG(C <= 8192, RN_E_TOOLARGE, "rn_postprocess: C=%d exceeds 8192", C);
    RN_CHECK_ARG(max_det >= 1 && max_det <= MAX_DET_CAP, RN_E_TOOLARGE, "rn_postprocess: max_det=%d outside [1,%d]",
                 max_det, MAX_DET_CAP);
    RN_CHECK_ARG(cand_capacity >= 1 && cand_capacity < 0x7fffffffLL, RN_E_BADARG, "rn_postprocess: bad cand_capacity");
    RN_CHECK_ARG(pre_nms_topk >= 0, RN_E_BADARG, "rn_postprocess: negative pre_nms_topk");
    RN_CHECK_ARG(pre_nms_topk == 0 || (algo == RN_PP_LAZY && level_off_host && num_levels >= 1 && num_levels <= RN_MAX_LEVELS),
                 RN_E_BADARG, "rn_postprocess: pre_nms_topk needs algo = RN_PP_LAZY and 1..%d level offsets", RN_MAX_LEVELS);
    RN_CHECK_ARG((size_t)N * C < 0x7fffffffULL, RN_E_TOOLARGE, "rn_postprocess: N*C too large");
    RN_CHECK_ARG(algo == RN_PP_LAZY || algo == RN_PP_GENERAL, RN_E_BADARG, "rn_postprocess: unknown algo %d", algo);
    RN_CHECK_ARG(algo != RN_PP_LAZY || (unsigned long long)A * (unsigned long long)C < (1ULL << 32), RN_E_TOOLARGE,
                 "rn_postprocess: the lazy algorithm needs A*C < 2^32 (use RN_PP_GENERAL)");
    PPWorkspace w = carve(workspace, N, C, cand_capacity);
    RN_CHECK_ARG(workspace_bytes >= w.total_bytes, RN_E_WORKSPACE, "rn_postprocess: workspace too small (%zu < %zu)",
                 workspace_bytes, w.total_bytes);
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(workspace, 0, w.zero_bytes, s);
    if (e == cudaSuccess) e = cudaMemsetAsync(out_status, 0, 4 * sizeof(int32_t), s);
    if (e != cudaSuccess) { rn_set_error("rn_postprocess: memset failed: %s", cudaGetErrorString(e)); return (int)e; }

    // guard logit: every x <= x_lo has sigmoid(x) <= thr with a 1e-3 margin (SFU/libm error is ~1e-7)
    const float thr = score_thr;
    float x_lo;
    if (!(thr > 0.0f)) x_lo = -INFINITY;
    else if (thr >= 1.0f) x_lo = INFINITY;
    else {
        double xt = log((double)thr / (1.0 - (double)thr));
        x_lo = (float)(xt - 1e-3 * fmax(1.0, fabs(xt)));
    }
    const bool lazy = algo == RN_PP_LAZY;
    FilterParams F;
    F.logits = logits; F.bbox = (const float4 *)bbox; F.anchors = (const float4 *)anchors; F.im_hw = im_hw;
    F.A = A; F.anchor_stride = anchor_image_stride; F.C = C; F.x_lo = x_lo; F.thr = thr; F.cap = (u32)cand_capacity;
    F.cap_n = (u32)max((int64_t)1, cand_capacity / N); F.w = w;
    F.wts = make_float4(weights_host[0], weights_host[1], weights_host[2], weights_host[3]);
    const bool vec4 = (C % 4 == 0) && (((uintptr_t)logits & 15) == 0);
    const int CV = vec4 ? C / 4 : C;
    F.magic = CV == 1 ? 0u : (unsigned)((0x100000000ULL + (unsigned)CV - 1) / (unsigned)CV);
    const long long warp_tasks = (A + PP_WSPAN - 1) / PP_WSPAN;
    dim3 grid((unsigned)((warp_tasks + PP_BLOCK / 32 - 1) / (PP_BLOCK / 32)), (unsigned)N);
    if (lazy) {
        if (vec4) score_filter_kernel<4, true><<<grid, PP_BLOCK, 0, s>>>(F); else score_filter_kernel<1, true><<<grid, PP_BLOCK, 0, s>>>(F);
    } else {
        if (vec4) score_filter_kernel<4, false><<<grid, PP_BLOCK, 0, s>>>(F); else score_filter_kernel<1, false><<<grid, PP_BLOCK, 0, s>>>(F);
    }
    RN_CHECK_LAUNCH("rn_postprocess/score_filter");

    BoxSource box;
    memset(&box, 0, sizeof(box));
    box.nac = (const float4 *)bbox;
    return pp_tail(w, F, box, anchors, anchor_image_stride, im_hw, N, A, C, score_thr, nms_thr, max_det, pre_nms_topk,
                   level_off_host, num_levels, lazy, cand_capacity, out_boxes, out_scores, out_labels, out_count,
                   out_status, out_ratio_hw, out_format, s);
}

// ---- per-level NCHW inputs ------------------------------------------------------------
// The filter streams each level's conv output [N, na*C, H, W] as the flat array it is (a 128-bit load
// may straddle two channel planes: only the rare survivors decompose their flat index into
// (anchor, class)); the box activations are read in place by the NMS kernels (BoxSource), candidates only.
struct LevelFilterParams {
    const float *cls;        // one level, [N, na*C*HW] flat
    long long len;           // na*C*HW (< 2^31)
    int HW, na, C;
    unsigned magicHW, magicC;   // ceil(2^32 / d); 0 when d == 1
    long long lvl_off;
};

template <bool LAZY>
__device__ __noinline__ void emit_candidates_level(const FilterParams &P, const LevelFilterParams &Q, float4 v, int nvals,
                                                   int n, long long e0, u64 *st_key, u32 *st_seg, int *st_n) {
    const float vals[4] = {v.x, v.y, v.z, v.w};
    for (int k = 0; k < nvals; ++k) {
        if (!(vals[k] > P.x_lo)) continue;
        const float s = rn::sigmoid_ref(vals[k]);
        if (!(s > P.thr)) continue;
        const unsigned e = (unsigned)(e0 + k);                      // < 2^31: exact magic division below
        int ch = Q.magicHW ? (int)__umulhi(e, Q.magicHW) : (int)e;
        int pos = (int)e - ch * Q.HW;
        if (pos < 0) { --ch; pos += Q.HW; }                          // magic quotient can be one too large for big e
        const int a = Q.magicC ? (int)__umulhi((unsigned)ch, Q.magicC) : ch;
        const int c = ch - a * Q.C;
        const long long anchor = Q.lvl_off + (long long)pos * Q.na + a;
        const u32 lo = LAZY ? (u32)((long long)c * P.A + anchor) : (u32)anchor;
        const u64 key = ((u64)(~__float_as_uint(s)) << 32) | (u64)lo;
        const u32 seg = (u32)(n * P.C + c);
        const int slot = atomicAdd(st_n, 1);
        if (slot < PP_WSTAGE) {
            st_key[slot] = key;
            if (!LAZY) st_seg[slot] = seg;
        } else if (LAZY) {
            const u32 gp = atomicAdd(P.w.img_count + n, 1u);
            if (gp < P.cap_n) P.w.pool_key[(size_t)n * P.cap_n + gp] = key;
        } else {
            const u32 gp = atomicAdd(P.w.pool_count, 1u);
            if (gp < P.cap) {
                P.w.pool_key[gp] = key;
                P.w.pool_seg[gp] = seg;
                atomicAdd(P.w.seg_count + seg, 1);
            }
        }
    }
}

constexpr int LVF_SPAN = 8192;     // floats per warp task

struct AllLevelsFilterParams {
    LevelFilterParams lv[RN_MAX_LEVELS];
    int task_base[RN_MAX_LEVELS + 1];   // first warp task of each level (prefix sums), [num_levels] = total
    int vec4[RN_MAX_LEVELS];
    int num_levels;
};

template <int VEC, bool LAZY>
__device__ __forceinline__ void filter_level_span(const FilterParams &P, const LevelFilterParams &Q, const int n,
                                                  const long long e_begin, u64 *st_key, u32 *st_seg, int *st_n) {
    const int lane = threadIdx.x & 31;
    const int span = (int)min((long long)LVF_SPAN, Q.len - e_begin);
    const float *src = Q.cls + (long long)n * Q.len + e_begin;
    const int nvec = span / VEC;
    for (int base = 0; base < nvec; base += 32 * PP_U) {
        float4 v[PP_U];
#pragma unroll
        for (int u = 0; u < PP_U; ++u) {
            const int f = base + u * 32 + lane;
            v[u] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
            if (f < nvec) {
                if (VEC == 4) v[u] = rn::ld_stream_f4((const float4 *)src + f);
                else v[u].x = __ldg(src + f);
            }
        }
#pragma unroll
        for (int u = 0; u < PP_U; ++u) {
            const float vmax = VEC == 4 ? fmaxf(fmaxf(v[u].x, v[u].y), fmaxf(v[u].z, v[u].w)) : v[u].x;
            if (vmax > P.x_lo)
                emit_candidates_level<LAZY>(P, Q, v[u], VEC, n, e_begin + (long long)(base + u * 32 + lane) * VEC, st_key,
                                            st_seg, st_n);
        }
    }
}

// ALL pyramid levels in one launch: blockIdx.x enumerates the warp tasks (LVF_SPAN floats each) of every level, so the
// small P5-P7 levels — 20 us launches of their own, latency-bound on 16..200 CTAs — stream in the shadow of P3.
template <bool LAZY>
__global__ void __launch_bounds__(PP_BLOCK, 4) score_filter_levels_kernel(const __grid_constant__ FilterParams P,
                                                                          const __grid_constant__ AllLevelsFilterParams L) {
    __shared__ u64 s_key[PP_BLOCK / 32][PP_WSTAGE];
    __shared__ u32 s_seg[LAZY ? 1 : PP_BLOCK / 32][LAZY ? 1 : PP_WSTAGE];
    __shared__ int s_n[PP_BLOCK / 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = blockIdx.y;
    const int task = blockIdx.x * (PP_BLOCK / 32) + warp;
    if (task >= L.task_base[L.num_levels]) return;
    int l = 0;
#pragma unroll
    for (int k = 1; k < RN_MAX_LEVELS; ++k)
        if (k < L.num_levels && task >= L.task_base[k]) l = k;
    const LevelFilterParams &Q = L.lv[l];
    const long long e_begin = (long long)(task - L.task_base[l]) * LVF_SPAN;
    u64 *st_key = s_key[warp];
    u32 *st_seg = LAZY ? nullptr : s_seg[warp];
    int *st_n = &s_n[warp];
    if (lane == 0) *st_n = 0;
    __syncwarp();
    if (L.vec4[l]) filter_level_span<4, LAZY>(P, Q, n, e_begin, st_key, st_seg, st_n);
    else filter_level_span<1, LAZY>(P, Q, n, e_begin, st_key, st_seg, st_n);
    __syncwarp();
    const int staged = min(*st_n, PP_WSTAGE);
    if (staged == 0) return;
    u32 gbase = 0;
    if (lane == 0) gbase = atomicAdd(LAZY ? P.w.img_count + n : P.w.pool_count, (u32)staged);
    gbase = __shfl_sync(0xffffffffu, gbase, 0);
    for (int i = lane; i < staged; i += 32) {
        const u32 gp = gbase + (u32)i;
        if (LAZY) {
            if (gp < P.cap_n) P.w.pool_key[(size_t)n * P.cap_n + gp] = st_key[i];
        } else if (gp < P.cap) {
            P.w.pool_key[gp] = st_key[i];
            P.w.pool_seg[gp] = st_seg[i];
            atomicAdd(P.w.seg_count + st_seg[i], 1);
        }
    }
}

extern "C" size_t rn_postprocess_levels_workspace_bytes(int N, int64_t A, int C, int64_t cand_capacity, int max_det) {
    return rn_postprocess_workspace_bytes(N, A, C, cand_capacity, max_det);   // the box levels are indexed in place
}

extern "C" int rn_postprocess_levels(const float *const *cls_levels_host, const float *const *bbox_levels_host,
                                     const int32_t *level_desc_host, int num_levels, const float *anchors,
                                     int64_t anchor_image_stride, const int32_t *im_hw, int N, int64_t A, int C,
                                     float score_thr, double nms_thr, int max_det, const float *weights_host,
                                     int pre_nms_topk, int algo, int64_t cand_capacity, float *out_boxes,
                                     float *out_scores, int64_t *out_labels, int32_t *out_count, int32_t *out_status,
                                     void *workspace, size_t workspace_bytes, rn_stream_t stream,
                                     const float *out_ratio_hw, int out_format) {
    RN_CHECK_ARG(out_format == 0 || out_format == 1, RN_E_BADARG, "rn_postprocess_levels: out_format must be 0 (xyxy) or 1 (xywh)");
    RN_CHECK_ARG(cls_levels_host && bbox_levels_host && level_desc_host && anchors && im_hw && weights_host && out_boxes &&
                     out_scores && out_labels && out_count && out_status && workspace,
                 RN_E_BADARG, "rn_postprocess_levels: null pointer");
    RN_CHECK_ARG(num_levels >= 1 && num_levels <= RN_MAX_LEVELS, RN_E_TOOLARGE, "rn_postprocess_levels: bad num_levels");
    RN_CHECK_ARG(N > 0 && A > 0 && C > 0 && N <= 65535 && C <= 8192 && A < (1LL << 32), RN_E_BADARG, "rn_postprocess_levels: bad N/A/C");
    RN_CHECK_ARG(max_det >= 1 && max_det <= MAX_DET_CAP, RN_E_TOOLARGE, "rn_postprocess_levels: bad max_det");
    RN_CHECK_ARG(cand_capacity >= 1 && cand_capacity < 0x7fffffffLL, RN_E_BADARG, "rn_postprocess_levels: bad cand_capacity");
    RN_CHECK_ARG((size_t)N * C < 0x7fffffffULL, RN_E_TOOLARGE, "rn_postprocess_levels: N*C too large");
    RN_CHECK_ARG(algo == RN_PP_LAZY || algo == RN_PP_GENERAL, RN_E_BADARG, "rn_postprocess_levels: unknown algo %d", algo);
    RN_CHECK_ARG(algo != RN_PP_LAZY || (unsigned long long)A * (unsigned long long)C < (1ULL << 32), RN_E_TOOLARGE,
                 "rn_postprocess_levels: the lazy algorithm needs A*C < 2^32");
    RN_CHECK_ARG(pre_nms_topk >= 0 && (pre_nms_topk == 0 || algo == RN_PP_LAZY), RN_E_BADARG,
                 "rn_postprocess_levels: pre_nms_topk needs algo = RN_PP_LAZY");
    int64_t level_off[RN_MAX_LEVELS + 1];
    level_off[0] = 0;
    for (int l = 0; l < num_levels; ++l) {
        const int32_t *d = level_desc_host + 3 * l;
        RN_CHECK_ARG(d[0] >= 0 && d[1] >= 0 && d[2] >= 1, RN_E_BADARG, "rn_postprocess_levels: bad level %d descriptor", l);
        level_off[l + 1] = level_off[l] + (int64_t)d[0] * d[1] * d[2];
    }
    RN_CHECK_ARG(level_off[num_levels] == A, RN_E_BADARG, "rn_postprocess_levels: levels hold %lld anchors, A = %lld",
                 (long long)level_off[num_levels], (long long)A);
    PPWorkspace w = carve(workspace, N, C, cand_capacity);
    RN_CHECK_ARG(workspace_bytes >= w.total_bytes, RN_E_WORKSPACE, "rn_postprocess_levels: workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(workspace, 0, w.zero_bytes, s);
    if (e == cudaSuccess) e = cudaMemsetAsync(out_status, 0, 4 * sizeof(int32_t), s);
    if (e != cudaSuccess) { rn_set_error("rn_postprocess_levels: memset failed: %s", cudaGetErrorString(e)); return (int)e; }
    const float thr = score_thr;
    float x_lo;
    if (!(thr > 0.0f)) x_lo = -INFINITY;
    else if (thr >= 1.0f) x_lo = INFINITY;
    else {
        double xt = log((double)thr / (1.0 - (double)thr));
        x_lo = (float)(xt - 1e-3 * fmax(1.0, fabs(xt)));
    }
    const bool lazy = algo == RN_PP_LAZY;
    FilterParams F;
    F.logits = nullptr; F.bbox = nullptr; F.anchors = (const float4 *)anchors; F.im_hw = im_hw;
    F.A = A; F.anchor_stride = anchor_image_stride; F.C = C; F.magic = 0; F.x_lo = x_lo; F.thr = thr; F.cap = (u32)cand_capacity;
    F.cap_n = (u32)max((int64_t)1, cand_capacity / N); F.w = w;
    F.wts = make_float4(weights_host[0], weights_host[1], weights_host[2], weights_host[3]);
    AllLevelsFilterParams L;
    BoxSource box;
    memset(&L, 0, sizeof(L));
    memset(&box, 0, sizeof(box));
    L.num_levels = box.nlev = num_levels;
    long long tasks = 0;
    for (int l = 0; l < num_levels; ++l) {
        const int32_t *d = level_desc_host + 3 * l;
        const int HW = d[0] * d[1], na = d[2];
        RN_CHECK_ARG(HW == 0 || (cls_levels_host[l] && bbox_levels_host[l]), RN_E_BADARG, "rn_postprocess_levels: null level %d", l);
        LevelFilterParams &Q = L.lv[l];
        Q.cls = cls_levels_host[l]; Q.len = (long long)na * C * HW; Q.HW = HW; Q.na = na; Q.C = C; Q.lvl_off = level_off[l];
        RN_CHECK_ARG(Q.len < (1LL << 31), RN_E_TOOLARGE, "rn_postprocess_levels: level %d has more than 2^31 elements per image", l);
        Q.magicHW = HW <= 1 ? 0u : (unsigned)((0x100000000ULL + (unsigned)HW - 1) / (unsigned)HW);
        Q.magicC = C == 1 ? 0u : (unsigned)((0x100000000ULL + (unsigned)C - 1) / (unsigned)C);
        L.vec4[l] = (Q.len % 4 == 0) && (((uintptr_t)Q.cls & 15) == 0);
        L.task_base[l] = (int)tasks;
        tasks += (Q.len + LVF_SPAN - 1) / LVF_SPAN;
        box.lvl[l] = bbox_levels_host[l]; box.off[l] = level_off[l]; box.HW[l] = HW; box.na[l] = na;
        RN_CHECK_ARG(tasks < 0x7fffffffLL, RN_E_TOOLARGE, "rn_postprocess_levels: too many tasks");
    }
    for (int l = num_levels; l <= RN_MAX_LEVELS; ++l) { L.task_base[l] = (int)tasks; box.off[l] = A; }
    if (tasks > 0) {
        dim3 grid((unsigned)((tasks + PP_BLOCK / 32 - 1) / (PP_BLOCK / 32)), (unsigned)N);
        if (lazy) score_filter_levels_kernel<true><<<grid, PP_BLOCK, 0, s>>>(F, L);
        else score_filter_levels_kernel<false><<<grid, PP_BLOCK, 0, s>>>(F, L);
        RN_CHECK_LAUNCH("rn_postprocess_levels/score_filter");
    }
    return pp_tail(w, F, box, anchors, anchor_image_stride, im_hw, N, A, C, score_thr, nms_thr, max_det, pre_nms_topk,
                   level_off, num_levels, lazy, cand_capacity, out_boxes, out_scores, out_labels, out_count, out_status,
                   out_ratio_hw, out_format, s);
}

extern "C" int rn_nms_segments(const float *boxes, const int32_t *seg_off, int num_segments, int64_t total_boxes,
                               double nms_thr, uint8_t *keep_flags, void *workspace, size_t workspace_bytes,
                               rn_stream_t stream) {
    RN_CHECK_ARG(boxes && seg_off && keep_flags && workspace, RN_E_BADARG, "rn_nms_segments: null pointer");
    RN_CHECK_ARG(num_segments >= 0 && total_boxes >= 0, RN_E_BADARG, "rn_nms_segments: negative size");
    RN_CHECK_ARG(workspace_bytes >= (size_t)total_boxes * 16, RN_E_WORKSPACE, "rn_nms_segments: workspace too small");
    if (num_segments == 0 || total_boxes == 0) return 0;
    float thr_f = (float)nms_thr;
    if ((double)thr_f > nms_thr) thr_f = nextafterf(thr_f, -INFINITY);
    NmsParams M;
    memset(&M.box, 0, sizeof(M.box)); M.anchors = nullptr; M.im_hw = nullptr; M.boxes = (const float4 *)boxes;
    M.keep_flags = keep_flags; M.A = 0; M.anchor_stride = 0; M.C = 1; M.thr = thr_f; M.wts = make_float4(1.f, 1.f, 1.f, 1.f);
    M.seg_off = seg_off; M.sorted_key = nullptr; M.kept_key = nullptr; M.kept_box = (float4 *)workspace;
    M.kept_count = nullptr;
    nms_kernel<true><<<num_segments, NMS_CHUNK, 0, (cudaStream_t)stream>>>(M);
    RN_CHECK_LAUNCH("rn_nms_segments");
    return 0;
}
