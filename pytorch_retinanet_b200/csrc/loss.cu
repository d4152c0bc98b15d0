// Fused sigmoid-focal + smooth-L1 loss (forward, and gradients in the same pass).
// Replaces RetinaNetLosses.calc_loss / forward (retinanet/losses.py:19-145) given rn_match's codes.
//
// Roofline: HBM.  Algorithmic bytes per image: 4*A*C (logits) + 16*A (bbox) + 4*A (codes)
// [+ 4*A*C + 16*A gradient writes]; anchors/GT are tiny and L2-resident.  No one-hot, no gathered
// copy of the logits, no [A,C] temporaries: the per-anchor target column comes from the packed
// code written by the matcher.
//
// Arithmetic (per logit v, reference semantics):
//   x = v + 1                                   losses.py:84 (+1 on the LOGITS)
//   p = sigmoid(x) (detached)                   losses.py:42
//   t=0: w = alpha   * p^gamma      loss = w * softplus(x)       dL/dx = w * p
//   t=1: w = (1-alpha)*(1-p)^gamma  loss = w * softplus(-x)      dL/dx = w * (p - 1)
//        (alpha is applied inverted, losses.py:44; weights carry no gradient, losses.py:42)
// Every element is first treated as a negative in a branch-free 128-bit vector loop; the (at most
// one) positive column of a foreground anchor is then patched.  Anchors with code -2 (ignore)
// contribute nothing and get zero gradient.
// Math modes: FAST (default) uses ex2/rcp/lg2 SFU approximations, and a warp-uniform series path
// when every logit of the warp is small (x <= -2.77, the overwhelmingly common case) that needs
// only one ex2 + one rcp per element and is accurate to < 1e-7 relative — lg2.approx near 1 has
// too much ABSOLUTE error for log1p(e) of tiny e, so it is only used for e > 1/16.
// PRECISE uses expf/log1pf/IEEE division (test reference for the approximation error).
// Reductions: fp32 per thread (<= a few hundred terms), fp64 from the warp level on, per-CTA
// partials to the workspace and a fixed-order final kernel => bit-reproducible.
#include <algorithm>
#include <cstring>

#include "loss_math.cuh"
#include "pp_internal.cuh"
#include "rn_common.cuh"

namespace {

constexpr int LOSS_BLOCK = 256;
constexpr int LOSS_SPAN = 256;   // anchors per CTA
#ifndef LOSS_U_DEF
#define LOSS_U_DEF 4
#endif
#ifndef LOSS_MINB
#define LOSS_MINB 5       // fwd+grad: 48 registers, 5 CTAs/SM (measured r2: 386 us; 6 CTAs 388 us, 4 CTAs 393 us at config 2)
#endif
#ifndef LOSS_MINB_FWD
#define LOSS_MINB_FWD 6   // forward only: 40 registers, 6 CTAs/SM (201 us; 5 CTAs 209 us, 4 CTAs 222 us, 8 CTAs x 2 loads 203 us)
#endif
constexpr int LOSS_U = LOSS_U_DEF;   // 128-bit loads in flight per thread

static int g_math_mode = 0;      // 0 fast, 1 precise

// kernel-side form of rn_exchange_t (see peer_exchange below)
struct ExchangeDev {
    unsigned long long *peers[RN_MAX_PEERS];
    int rank, world;
};

struct LossParams {
    const float *logits;
    const float4 *bbox;
    const float4 *anchors;
    const float4 *gt;
    const int *gt_off;
    const int *codes;
    const int *fg_count;
    float *grad_logits;
    float4 *grad_bbox;
    double *partials;            // [N][chunks][2]
    unsigned *ticket;            // the final reduction's ticket: re-armed here, so no memset node is needed
    long long A;
    long long anchor_stride;
    int C;
    int chunks;
    float alpha, gamma, beta, batch_div;
    float4 wts;
    rnpp::LazySink sink;         // FILTER variants (rn_train_detect): candidate lists of the post-processing
    // FILTER variants also run the final reduction themselves (the last CTA of an image, by ticket): one node less
    unsigned *img_ticket;        // [N], zeroed by prep_kernel
    float *out_image, *out_total;
    double *tail;
    int N;
    ExchangeDev xch;
};

using namespace rnloss;

// rn_train_detect: the loss kernel is also the score filter of the post-processing — the logits are streamed ONCE for
// both halves of the path.  Rare path, deliberately not inlined (as in score_filter_kernel): exact sigmoid of the <= 4
// logits of one vector, strict threshold (models.py:196), key = (~score_bits << 32) | (class * A + anchor), appended to
// the image's candidate list with one atomic per survivor (~0.1 % of the elements).
constexpr int LOSS_STAGE = 192;   // candidate keys staged per CTA before ONE global append (a warp must not wait for an
                                  // L2 atomic's round trip per survivor: ~11 % of the warp-vectors hold one)
__device__ __noinline__ void emit_candidates_loss(const LossParams &P, float4 v, int f, int n, long long a0, int CV,
                                                  unsigned long long *s_cand, int *s_ncand) {
    const int al = f / CV;
    const int c0 = (f - al * CV) * 4;
    const long long anchor = a0 + al;
    const float vals[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (!(vals[k] > P.sink.x_lo)) continue;
        const float s = rn::sigmoid_ref(vals[k]);
        if (!(s > P.sink.thr)) continue;
        const unsigned lo = (unsigned)((long long)(c0 + k) * P.A + anchor);
        const unsigned long long key = ((unsigned long long)(~__float_as_uint(s)) << 32) | (unsigned long long)lo;
        const int pos = atomicAdd(s_ncand, 1);
        if (pos < LOSS_STAGE) {
            s_cand[pos] = key;
        } else {                                                // staging full (dense crowd): straight to the list
            const unsigned gp = atomicAdd(P.sink.img_count + n, 1u);
            if (gp < P.sink.cap_n) P.sink.pool_key[(size_t)n * P.sink.cap_n + gp] = key;
        }
    }
}
// all threads of the CTA: one global atomic for the staged candidates, then a coalesced copy
__device__ __forceinline__ void flush_candidates_loss(const LossParams &P, int n, const unsigned long long *s_cand,
                                                      const int *s_ncand, unsigned *s_gbase) {
    __syncthreads();
    const int staged = min(*s_ncand, LOSS_STAGE);
    if (staged == 0) return;                                    // block-uniform
    if (threadIdx.x == 0) *s_gbase = atomicAdd(P.sink.img_count + n, (unsigned)staged);
    __syncthreads();
    const unsigned gbase = *s_gbase;
    for (int i = threadIdx.x; i < staged; i += LOSS_BLOCK) {
        const unsigned gp = gbase + (unsigned)i;
        if (gp < P.sink.cap_n) P.sink.pool_key[(size_t)n * P.sink.cap_n + gp] = s_cand[i];
    }
}

__device__ __forceinline__ void block_sum3(double &a, double &b, double &c) {
    __shared__ double s[3][LOSS_BLOCK / 32];
    a = rn::warp_sum(a);
    b = rn::warp_sum(b);
    c = rn::warp_sum(c);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { s[0][w] = a; s[1][w] = b; s[2][w] = c; }
    __syncthreads();
    if (threadIdx.x == 0) {
        a = b = c = 0.0;
#pragma unroll
        for (int i = 0; i < LOSS_BLOCK / 32; ++i) { a += s[0][i]; b += s[1][i]; c += s[2][i]; }
    }
}

// One CTA's share of the loss: LOSS_SPAN anchors of image n.  Writes the CTA's partial sums.
// VEC = 4: C % 4 == 0 and 16-byte aligned rows (128-bit path); VEC = 1: any C (scalar path).
//
// The streaming loop knows nothing about targets: EVERY element is accumulated as a negative (and gets the negative's
// gradient) — no per-vector code load, no index arithmetic, no select.  The per-anchor epilogue then repairs the rows
// that are not plain negatives, all of them rare: the ONE positive column of a foreground anchor is re-read (L2 hit),
// its negative term taken back and its gradient element overwritten; the rows of ignore anchors (IoU in the
// [bg, fg] band) are re-read by their warp, 32 vectors at a time, their contribution subtracted and their gradient
// rows zeroed.  An image without GT boxes (every anchor is ignore, box_utils.py:70-71) is skipped outright, so its
// loss is exactly 0.  NB: a NaN / +inf logit inside an IGNORED anchor therefore poisons the sum (inf - inf), where
// the reference, which drops those rows before any arithmetic, stays finite; everywhere else non-finite logits
// propagate exactly as in the reference.
template <int VEC, bool WANT_GRAD, bool GAMMA2, bool PRECISE, bool FILTER>
__device__ __forceinline__ void loss_chunk(const LossParams &P, const int n, const int chunk) {
    const long long a0 = (long long)chunk * LOSS_SPAN;
    const int span = (int)min((long long)LOSS_SPAN, P.A - a0);
    const long long row0 = (long long)n * P.A + a0;
    const int CV = P.C / VEC;                       // vectors per anchor row
    const int nvec = span * CV;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const float *src = P.logits + row0 * P.C;
    float *dst = WANT_GRAD ? P.grad_logits + row0 * P.C : nullptr;
    __shared__ unsigned long long s_cand[FILTER ? LOSS_STAGE : 1];
    __shared__ int s_ncand;
    __shared__ unsigned s_gbase;
    if (FILTER) {
        if (t == 0) s_ncand = 0;
        __syncthreads();
    }

    if (__ldg(P.gt_off + n + 1) == __ldg(P.gt_off + n)) {      // no GT: nothing contributes, gradients are zero
        if (FILTER && VEC == 4) {                               // ... but the image still has detections
            for (int f = t; f < nvec; f += LOSS_BLOCK) {
                const float4 q = rn::ld_stream_f4((const float4 *)src + f);
                if (fmaxf(fmaxf(q.x, q.y), fmaxf(q.z, q.w)) > P.sink.x_lo) emit_candidates_loss(P, q, f, n, a0, CV, s_cand, &s_ncand);
            }
            flush_candidates_loss(P, n, s_cand, &s_ncand, &s_gbase);
        }
        if (WANT_GRAD) {
            for (int f = t; f < nvec; f += LOSS_BLOCK) {
                if (VEC == 4) rn::st_stream_f4((float4 *)dst + f, make_float4(0.f, 0.f, 0.f, 0.f));
                else dst[f] = 0.0f;
            }
            if (t < span) P.grad_bbox[row0 + t] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (t == 0) {
            double *o = P.partials + ((long long)n * P.chunks + chunk) * 2;
            o[0] = 0.0;
            o[1] = 0.0;
        }
        return;
    }
    const int F = __ldg(P.fg_count + n);
    const float inv = 1.0f / (fmaxf((float)F, 1.0f) * P.batch_div);   // gradient scale
    const float neg_gscale = P.alpha * inv;
    const float x_mid = kMidX - 1.0f;
    // this thread's anchor of the epilogue: fetched now, so that its dependent loads (GT box, logit of the positive
    // column) are one memory round trip after the streaming loop instead of two
    const int code = t < span ? __ldg(P.codes + row0 + t) : -1;
    const int gt0 = __ldg(P.gt_off + n);

    float acc_neg = 0.0f;   // sum p^g * softplus(x), every element treated as a negative
    float acc_pos = 0.0f;   // positive terms minus what the positives' columns added to acc_neg (alpha-weighted)

    for (int base = 0; base < nvec; base += LOSS_BLOCK * LOSS_U) {
        float v[LOSS_U][VEC];
#pragma unroll
        for (int u = 0; u < LOSS_U; ++u) {
            const int f = base + u * LOSS_BLOCK + t;
            if (VEC == 4) {
                float4 q = make_float4(-100.f, -100.f, -100.f, -100.f);    // padding: every term is exactly 0
                if (f < nvec) q = rn::ld_stream_f4((const float4 *)src + f);
                v[u][0] = q.x; v[u][VEC > 1 ? 1 : 0] = q.y; v[u][VEC > 2 ? 2 : 0] = q.z; v[u][VEC > 3 ? 3 : 0] = q.w;
            } else {
                v[u][0] = f < nvec ? __ldg(src + f) : -100.f;
            }
        }
#pragma unroll
        for (int u = 0; u < LOSS_U; ++u) {
            float g[VEC];
            if (FILTER && VEC == 4) {                            // the post-processing's score filter, on the same registers
                const float vmax = fmaxf(fmaxf(v[u][0], v[u][1]), fmaxf(v[u][VEC > 2 ? 2 : 0], v[u][VEC > 3 ? 3 : 0]));
                if (vmax > P.sink.x_lo && base + u * LOSS_BLOCK + t < nvec)     // rare
                    emit_candidates_loss(P, make_float4(v[u][0], v[u][1], v[u][VEC > 2 ? 2 : 0], v[u][VEC > 3 ? 3 : 0]),
                                         base + u * LOSS_BLOCK + t, n, a0, CV, s_cand, &s_ncand);
            }
            bool mid = !PRECISE;
#pragma unroll
            for (int k = 0; k < VEC; ++k) mid = mid && (v[u][k] <= x_mid);       // false for NaN
            float local = 0.0f;
            if (!PRECISE && __all_sync(0xffffffffu, mid)) {
#pragma unroll
                for (int k = 0; k < VEC; ++k) {
                    float wp;
                    focal_neg_mid<WANT_GRAD, GAMMA2>(v[u][k], P.gamma, local, wp);
                    g[k] = wp * neg_gscale;
                }
            } else {
#pragma unroll
                for (int k = 0; k < VEC; ++k) {
                    float pk, spk;
                    sigmoid_softplus<PRECISE>(v[u][k] + 1.0f, pk, spk);
                    const float w = pow_gamma<GAMMA2>(pk, P.gamma);
                    local = fmaf(w, spk, local);
                    g[k] = w * pk * neg_gscale;
                }
            }
            acc_neg += local;
            if (WANT_GRAD) {
                const int f = base + u * LOSS_BLOCK + t;
                if (f < nvec) {
                    if (VEC == 4) {
                        rn::st_stream_f4((float4 *)dst + f, make_float4(g[0], g[VEC > 1 ? 1 : 0], g[VEC > 2 ? 2 : 0],
                                                                        g[VEC > 3 ? 3 : 0]));
                    } else {
                        dst[f] = g[0];
                    }
                }
            }
        }
    }

    if (FILTER) flush_candidates_loss(P, n, s_cand, &s_ncand, &s_gbase);
    // ---- per-anchor epilogue (thread t <-> anchor t of the span); the gradient rows written above are finished ----
    if (WANT_GRAD) __syncthreads();
    // (1) ignore anchors: the warp takes their rows back, CV vectors over the 32 lanes
    unsigned ign = __ballot_sync(0xffffffffu, code == -2);
    while (ign) {
        const int arow = warp * 32 + __ffs(ign) - 1;
        ign &= ign - 1;
        const float *row = src + (long long)arow * P.C;
        float sub = 0.0f;
        for (int j = lane; j < CV; j += 32) {
            float q[VEC];
            if (VEC == 4) {
                const float4 w4 = __ldg((const float4 *)row + j);
                q[0] = w4.x; q[VEC > 1 ? 1 : 0] = w4.y; q[VEC > 2 ? 2 : 0] = w4.z; q[VEC > 3 ? 3 : 0] = w4.w;
            } else {
                q[0] = __ldg(row + j);
            }
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                float pk, spk;
                sigmoid_softplus<PRECISE>(q[k] + 1.0f, pk, spk);
                sub = fmaf(pow_gamma<GAMMA2>(pk, P.gamma), spk, sub);
            }
            if (WANT_GRAD) {
                float *grow = dst + (long long)arow * P.C;
                if (VEC == 4) ((float4 *)grow)[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                else grow[j] = 0.0f;
            }
        }
        acc_neg -= sub;
    }
    // (2) foreground anchors: the positive column (losses.py:96-105) and the regression term (losses.py:66-71, 19-27)
    float reg = 0.0f;
    if (t < span) {
        float4 gb = make_float4(0.f, 0.f, 0.f, 0.f);
        if (code >= 0) {
            const int cls = code >> 20;
            if (cls < P.C) {                       // cls = kNoClass (label outside 1..C): all-negative class targets
                const float x = __ldg(src + (long long)t * P.C + cls) + 1.0f;
                float p, sp;
                sigmoid_softplus<PRECISE>(x, p, sp);
                const float wn = pow_gamma<GAMMA2>(p, P.gamma);
                const float wpos = pow_gamma<GAMMA2>(1.0f - p, P.gamma) * (1.0f - P.alpha);
                acc_pos += wpos * (sp - x) - P.alpha * wn * sp;   // softplus(-x) = softplus(x) - x
                if (WANT_GRAD) dst[(long long)t * P.C + cls] = wpos * (p - 1.0f) * inv;
            }
            const float4 gtb = P.gt[gt0 + (code & 0xFFFFF)];
            const float4 an = P.anchors[(long long)n * P.anchor_stride + a0 + t];
            const float4 pr = P.bbox[row0 + t];
            const float4 tt = rn::encode_box(gtb, an, P.wts);
            const float d[4] = {pr.x - tt.x, pr.y - tt.y, pr.z - tt.z, pr.w - tt.w};
            float gr[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float nabs = fabsf(d[k]);
                if (P.beta < 1e-5f) {
                    reg += nabs;
                    gr[k] = d[k] > 0.0f ? inv : (d[k] < 0.0f ? -inv : 0.0f);
                } else if (nabs < P.beta) {
                    reg += 0.5f * nabs * nabs / P.beta;
                    gr[k] = d[k] / P.beta * inv;
                } else {
                    reg += nabs - 0.5f * P.beta;
                    gr[k] = d[k] > 0.0f ? inv : -inv;
                }
            }
            gb = make_float4(gr[0], gr[1], gr[2], gr[3]);
        }
        if (WANT_GRAD) P.grad_bbox[row0 + t] = gb;
    }

    double s_neg = acc_neg, s_pos = acc_pos, s_reg = reg;
    block_sum3(s_neg, s_pos, s_reg);
    if (threadIdx.x == 0) {
        double *o = P.partials + ((long long)n * P.chunks + chunk) * 2;
        o[0] = (double)P.alpha * s_neg + s_pos;
        o[1] = s_reg;
    }
}

__device__ void finalize_image_fused(const LossParams &P, int n);

template <int VEC, bool WANT_GRAD, bool GAMMA2, bool PRECISE, bool FILTER = false>
__global__ void __launch_bounds__(LOSS_BLOCK, PRECISE ? 1 : (WANT_GRAD ? LOSS_MINB : LOSS_MINB_FWD))
loss_kernel(const __grid_constant__ LossParams P) {
    if (!FILTER && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *P.ticket = 0u;   // loss_finalize_kernel runs after this grid
    loss_chunk<VEC, WANT_GRAD, GAMMA2, PRECISE, FILTER>(P, blockIdx.y, blockIdx.x);
    if (FILTER) {            // the CTA that completes an image reduces it; the one that completes the batch finishes the step
        __shared__ bool s_img_last;
        if (threadIdx.x == 0) {
            __threadfence();
            s_img_last = atomicAdd(P.img_ticket + blockIdx.y, 1u) == (unsigned)(P.chunks - 1);
        }
        __syncthreads();
        if (s_img_last) finalize_image_fused(P, blockIdx.y);
    }
}

// Fixed-order final reduction: block n reduces image n's chunk partials (thread-strided partial sums, then a
// fixed tree), the last block to finish (ticket) sums the images in index order.  The summation order
// depends only on (N, chunks) => bit-reproducible.  `tail` = [N][2] image sums + ticket, after the partials.
constexpr int FIN_BLOCK = 256;

// Image-sharded multi-GPU exchange folded into the final reduction (include/retinanet_b200.h,
// rn_exchange_t).  Receive buffer of a rank, in 8-byte words: slot[parity][sender][4] = (seq << 32) | float bits,
// then the rank's own step counter and error word.  An aligned 8-byte store is single-copy atomic, so a word whose
// upper half equals the expected sequence number carries a complete value (the LL idea: no flag, no fence).
// Two parities: a rank can run at most one step ahead of the slowest reader of its previous values.
constexpr int XCH_SLOT_WORDS = 2 * RN_MAX_PEERS * 4;
constexpr long long XCH_TIMEOUT_CYCLES = 10000000000LL;        // ~5 s at 1.9 GHz (callers align the ranks before the first step)
__device__ __forceinline__ void st_sys_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_sys_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Called by ALL threads of a block (>= 32 threads) after s_v / s_seq were written and a barrier: thread r stores the
// local vector into rank r's slots and collects rank r's vector from the local slots; out_total = sum in rank order.
__device__ __forceinline__ void peer_exchange(const float *s_v, const unsigned seq, float (*s_all)[4],
                                              float *__restrict__ out_total, const ExchangeDev &X) {
    const int t = threadIdx.x;
    if (t < X.world) {                                          // thread t talks to rank t
        const int par = (int)(seq & 1u);
        unsigned long long *dst = X.peers[t] + (par * RN_MAX_PEERS + X.rank) * 4;
#pragma unroll
        for (int k = 0; k < 4; ++k) st_sys_u64(dst + k, ((unsigned long long)seq << 32) | __float_as_uint(s_v[k]));
        const unsigned long long *src = X.peers[X.rank] + (par * RN_MAX_PEERS + t) * 4;
        const long long t0 = clock64();
        bool ok = true;
#pragma unroll 1
        for (int k = 0; k < 4; ++k) {
            unsigned long long w = 0;
            while (ok) {
                w = ld_sys_u64(src + k);
                if ((unsigned)(w >> 32) == seq) break;
                if (clock64() - t0 > XCH_TIMEOUT_CYCLES) ok = false;
            }
            s_all[t][k] = ok ? __uint_as_float((unsigned)w) : __int_as_float(0x7fc00000);
        }
        if (!ok) *((unsigned *)(X.peers[X.rank] + XCH_SLOT_WORDS) + 1) = 1u;    // error word: a peer never arrived
    }
    __syncthreads();
    if (t < 4) {                                                // rank order: the same bits on every rank
        double acc = 0.0;
        for (int rk = 0; rk < X.world; ++rk) acc += (double)s_all[rk][t];
        out_total[t] = (float)acc;
    }
}

// The exchange by itself (a rank whose shard is empty launches no loss kernel but must still take part).
__global__ void __launch_bounds__(32) exchange_kernel(float *__restrict__ total, const __grid_constant__ ExchangeDev X) {
    __shared__ float s_v[4];
    __shared__ unsigned s_seq;
    __shared__ float s_all[RN_MAX_PEERS][4];
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) s_v[k] = total[k];
        unsigned *seqp = (unsigned *)(X.peers[X.rank] + XCH_SLOT_WORDS);
        s_seq = *seqp + 1u;
        *seqp = s_seq;
    }
    __syncthreads();
    peer_exchange(s_v, s_seq, s_all, total, X);
}

// Called by all FIN_BLOCK threads of a CTA for image n.
__device__ __forceinline__ void finalize_image(const double *partials, const int *fg_count, const int n, const int N,
                                               const int chunks, const float batch_div, float *__restrict__ out_image,
                                               float *__restrict__ out_total, double *tail, const ExchangeDev &X) {
    __shared__ double s_c[FIN_BLOCK / 32], s_r[FIN_BLOCK / 32];
    __shared__ bool s_last;
    __shared__ float s_v[4];
    __shared__ unsigned s_seq;
    __shared__ float s_all[RN_MAX_PEERS][4];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const double2 *p = (const double2 *)partials + (long long)n * chunks;
    double c = 0.0, r = 0.0;
    for (int k = t; k < chunks; k += 4 * FIN_BLOCK) {          // 4 independent loads in flight; fixed summation order
        double2 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = k + u * FIN_BLOCK < chunks ? __ldcg(p + k + u * FIN_BLOCK) : make_double2(0.0, 0.0);
#pragma unroll
        for (int u = 0; u < 4; ++u) { c += v[u].x; r += v[u].y; }
    }
    c = rn::warp_sum(c);
    r = rn::warp_sum(r);
    if (lane == 0) { s_c[warp] = c; s_r[warp] = r; }
    __syncthreads();
    unsigned *ticket = (unsigned *)(tail + 2 * (long long)N);
    if (t == 0) {
        double cs = 0.0, rs = 0.0;
#pragma unroll
        for (int w = 0; w < FIN_BLOCK / 32; ++w) { cs += s_c[w]; rs += s_r[w]; }
        const int F = __ldcg(fg_count + n);
        const double den = F > 0 ? (double)F : 1.0;             // clamp(F, min=1)  losses.py:108-109
        cs /= den;
        rs /= den;
        tail[2 * n] = cs;
        tail[2 * n + 1] = rs;
        if (out_image) {
            out_image[3 * n + 0] = (float)cs;
            out_image[3 * n + 1] = (float)rs;
            out_image[3 * n + 2] = (float)F;
        }
        __threadfence();
        s_last = atomicAdd(ticket, 1u) == (unsigned)(N - 1);
    }
    __syncthreads();
    if (!s_last) return;                                        // block-uniform
    // the last block sums the images in index order: FIN_BLOCK of them are fetched at a time (one round trip), thread 0
    // adds them up serially from shared memory
    __shared__ double s_ic[FIN_BLOCK], s_ir[FIN_BLOCK];
    __shared__ int s_if[FIN_BLOCK];
    double cs = 0.0, rs = 0.0;
    long long fsum = 0;
    __threadfence();
    for (int i0 = 0; i0 < N; i0 += FIN_BLOCK) {
        const int i = i0 + t;
        if (i < N) {
            const volatile double *vt = tail;
            s_ic[t] = vt[2 * i];
            s_ir[t] = vt[2 * i + 1];
            s_if[t] = __ldcg(fg_count + i);
        }
        __syncthreads();
        if (t == 0) {
            const int m = min(FIN_BLOCK, N - i0);
            for (int q = 0; q < m; ++q) { cs += s_ic[q]; rs += s_ir[q]; fsum += s_if[q]; }
        }
        __syncthreads();
    }
    if (t == 0) {
        s_v[0] = (float)(cs / (double)batch_div);               // losses.py:138-140
        s_v[1] = (float)(rs / (double)batch_div);
        s_v[2] = (float)fsum;                                   // sum_n F_n   } carried for the image-sharded
        s_v[3] = (float)N;                                      // local images } exchange
        if (X.world <= 1) {
#pragma unroll
            for (int k = 0; k < 4; ++k) out_total[k] = s_v[k];
        } else {
            unsigned *seqp = (unsigned *)(X.peers[X.rank] + XCH_SLOT_WORDS);
            s_seq = *seqp + 1u;                                 // this rank's step counter (same on every rank)
            *seqp = s_seq;
        }
    }
    if (X.world <= 1) return;
    __syncthreads();
    peer_exchange(s_v, s_seq, s_all, out_total, X);
}

__global__ void __launch_bounds__(FIN_BLOCK) loss_finalize_kernel(const double *__restrict__ partials,
                                                                  const int *__restrict__ fg_count, int N, int chunks,
                                                                  float batch_div, float *__restrict__ out_image,
                                                                  float *__restrict__ out_total, double *__restrict__ tail,
                                                                  const __grid_constant__ ExchangeDev X) {
    finalize_image(partials, fg_count, blockIdx.x, N, chunks, batch_div, out_image, out_total, tail, X);
}

__device__ void finalize_image_fused(const LossParams &P, int n) {
    finalize_image(P.partials, P.fg_count, n, P.N, P.chunks, P.batch_div, P.out_image, P.out_total, P.tail, P.xch);
}

// rn_train_detect: everything the step's kernels expect to find zeroed, in ONE launch instead of three memset nodes.
__global__ void __launch_bounds__(256) prep_kernel(int *fg_count, unsigned *img_ticket, unsigned *ticket, int *status,
                                                   unsigned *img_count, int N) {
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        fg_count[i] = 0;
        img_ticket[i] = 0u;
        img_count[i] = 0u;
    }
    if (threadIdx.x < 4) status[threadIdx.x] = 0;
    if (threadIdx.x == 0) *ticket = 0u;
}

// host rn_exchange_t -> kernel parameter; returns false on a malformed descriptor
inline bool make_exchange(const rn_exchange_t *x, ExchangeDev &X) {
    memset(&X, 0, sizeof(X));
    X.world = 1;
    if (!x || x->world <= 1) return true;
    if (x->world > RN_MAX_PEERS || x->rank < 0 || x->rank >= x->world) return false;
    for (int r = 0; r < x->world; ++r) {
        if (!x->peers[r]) return false;
        X.peers[r] = (unsigned long long *)x->peers[r];
    }
    X.rank = x->rank;
    X.world = x->world;
    return true;
}

__global__ void __launch_bounds__(256) scale_kernel(float *__restrict__ buf, long long n, const float *__restrict__ scale) {
    const float s = __ldg(scale);
    if (s == 1.0f) return;
    const long long n4 = n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    float4 *b4 = (float4 *)buf;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 v = b4[i];
        v.x *= s; v.y *= s; v.z *= s; v.w *= s;
        b4[i] = v;
    }
    for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) buf[i] *= s;
}


template <int VEC, bool WANT_GRAD, bool GAMMA2>
void launch_loss(const LossParams &P, dim3 grid, cudaStream_t s, bool precise, bool filter) {
    if (filter) {
        if (VEC == 4 && !precise) loss_kernel<4, WANT_GRAD, GAMMA2, false, true><<<grid, LOSS_BLOCK, 0, s>>>(P);
    } else if (precise)
        loss_kernel<VEC, WANT_GRAD, GAMMA2, true><<<grid, LOSS_BLOCK, 0, s>>>(P);
    else
        loss_kernel<VEC, WANT_GRAD, GAMMA2, false><<<grid, LOSS_BLOCK, 0, s>>>(P);
}
template <int VEC, bool WANT_GRAD>
void launch_loss_g(const LossParams &P, dim3 grid, cudaStream_t s, bool precise, bool filter) {
    if (P.gamma == 2.0f)
        launch_loss<VEC, WANT_GRAD, true>(P, grid, s, precise, filter);
    else
        launch_loss<VEC, WANT_GRAD, false>(P, grid, s, precise, filter);
}

// (host helpers)
inline int loss_chunks(int64_t A) { return (int)((A + LOSS_SPAN - 1) / LOSS_SPAN); }

}  // namespace

extern "C" int rn_loss_set_math_mode(int mode) {
    int old = g_math_mode;
    g_math_mode = mode ? 1 : 0;
    return old;
}

extern "C" int rn_exchange_total(float *total, const rn_exchange_t *exchange_host, rn_stream_t stream) {
    RN_CHECK_ARG(total && exchange_host, RN_E_BADARG, "rn_exchange_total: null pointer");
    ExchangeDev X;
    RN_CHECK_ARG(make_exchange(exchange_host, X), RN_E_BADARG, "rn_exchange_total: malformed exchange descriptor");
    if (X.world <= 1) return 0;
    exchange_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(total, X);
    RN_CHECK_LAUNCH("rn_exchange_total");
    return 0;
}

extern "C" size_t rn_loss_workspace_bytes(int N, int64_t A, int C) {
    (void)C;
    if (N <= 0 || A <= 0) return 16;
    return ((size_t)N * (size_t)loss_chunks(A) * 2 + (size_t)N * 2 + 2 + ((size_t)N + 1) / 2) * sizeof(double);
}

static int loss_impl(const float *logits, const float *bbox, const float *anchors, int64_t anchor_image_stride,
                     const float *gt_boxes,
                     const int32_t *gt_off, const int32_t *codes, const int32_t *fg_count, int N, int64_t A, int C,
                     float alpha, float gamma, float beta, const float *weights_host, float batch_div,
                     float *out_image, float *out_total, float *grad_logits, float *grad_bbox, void *workspace,
                     size_t workspace_bytes, rn_stream_t stream, const rn_exchange_t *exchange_host,
                     const rnpp::LazySink *sink) {
    RN_CHECK_ARG(logits && bbox && anchors && gt_off && codes && fg_count && out_total && weights_host, RN_E_BADARG,
                 "rn_loss: null pointer");
    ExchangeDev X;
    RN_CHECK_ARG(make_exchange(exchange_host, X), RN_E_BADARG, "rn_loss: malformed exchange descriptor");
    RN_CHECK_ARG(N > 0 && A > 0 && C > 0, RN_E_BADARG, "rn_loss: N, A, C must be positive (got %d, %lld, %d)", N,
                 (long long)A, C);
    RN_CHECK_ARG(C <= 2047, RN_E_TOOLARGE, "rn_loss: C=%d exceeds 2047 classes", C);
    RN_CHECK_ARG(N <= 65535, RN_E_TOOLARGE, "rn_loss: N=%d exceeds 65535 images per call", N);
    RN_CHECK_ARG((grad_logits == nullptr) == (grad_bbox == nullptr), RN_E_BADARG,
                 "rn_loss: grad_logits and grad_bbox must be given together");
    RN_CHECK_ARG(batch_div > 0.0f, RN_E_BADARG, "rn_loss: batch_div must be positive");
    RN_CHECK_ARG(workspace && workspace_bytes >= rn_loss_workspace_bytes(N, A, C), RN_E_WORKSPACE,
                 "rn_loss: workspace too small (%zu < %zu)", workspace_bytes, rn_loss_workspace_bytes(N, A, C));
    cudaStream_t s = (cudaStream_t)stream;
    LossParams P;
    P.logits = logits; P.bbox = (const float4 *)bbox; P.anchors = (const float4 *)anchors;
    P.gt = (const float4 *)gt_boxes; P.gt_off = gt_off; P.codes = codes; P.fg_count = fg_count;
    P.grad_logits = grad_logits; P.grad_bbox = (float4 *)grad_bbox; P.partials = (double *)workspace;
    P.ticket = (unsigned *)((double *)workspace + (size_t)N * loss_chunks(A) * 2 + 2 * (size_t)N);
    P.img_ticket = (unsigned *)((double *)workspace + (size_t)N * loss_chunks(A) * 2 + 2 * (size_t)N + 2);
    P.tail = (double *)workspace + (size_t)N * loss_chunks(A) * 2;
    P.out_image = out_image; P.out_total = out_total; P.N = N; P.xch = X;
    P.A = A; P.anchor_stride = anchor_image_stride; P.C = C; P.chunks = loss_chunks(A);
    P.alpha = alpha; P.gamma = gamma; P.beta = beta; P.batch_div = batch_div;
    P.wts = make_float4(weights_host[0], weights_host[1], weights_host[2], weights_host[3]);
    const bool vec4 = (C % 4 == 0) && (((uintptr_t)logits & 15) == 0) && (!grad_logits || ((uintptr_t)grad_logits & 15) == 0);
    dim3 grid((unsigned)P.chunks, (unsigned)N);
    const bool precise = g_math_mode == 1;
    const bool filter = sink != nullptr;
    memset(&P.sink, 0, sizeof(P.sink));
    if (filter) {
        RN_CHECK_ARG(vec4 && !precise, RN_E_BADARG, "rn_train_detect needs C %% 4 == 0, 16-byte aligned logits and the default math "
                                                    "mode (call rn_train_loss and rn_postprocess separately otherwise)");
        P.sink = *sink;
    }
    if (vec4) {
        if (grad_logits) launch_loss_g<4, true>(P, grid, s, precise, filter); else launch_loss_g<4, false>(P, grid, s, precise, filter);
    } else {
        if (grad_logits) launch_loss_g<1, true>(P, grid, s, precise, false); else launch_loss_g<1, false>(P, grid, s, precise, false);
    }
    RN_CHECK_LAUNCH("rn_loss");
    if (!filter) {           // (the FILTER variant of the kernel finishes the reduction itself)
        double *tail = (double *)workspace + (size_t)N * P.chunks * 2;
        loss_finalize_kernel<<<N, FIN_BLOCK, 0, s>>>((const double *)workspace, fg_count, N, P.chunks, batch_div, out_image,
                                                     out_total, tail, X);
        RN_CHECK_LAUNCH("rn_loss_finalize");
    }
    return 0;
}

extern "C" int rn_loss(const float *logits, const float *bbox, const float *anchors, int64_t anchor_image_stride,
                       const float *gt_boxes,
                       const int32_t *gt_off, const int32_t *codes, const int32_t *fg_count, int N, int64_t A, int C,
                       float alpha, float gamma, float beta, const float *weights_host, float batch_div,
                       float *out_image, float *out_total, float *grad_logits, float *grad_bbox, void *workspace,
                       size_t workspace_bytes, rn_stream_t stream, const rn_exchange_t *exchange_host) {
    return loss_impl(logits, bbox, anchors, anchor_image_stride, gt_boxes, gt_off, codes, fg_count, N, A, C, alpha, gamma, beta,
                     weights_host, batch_div, out_image, out_total, grad_logits, grad_bbox, workspace, workspace_bytes, stream,
                     exchange_host, nullptr);
}

// ---- rn_train_loss: the training half of the path behind ONE C call (rn_match, then rn_loss + final reduction on
// the same stream).  Two single-launch fusions of the matcher into the loss stream were built and measured on
// B200 (ticket-lagged CTAs; persistent warp-specialised CTAs with a dedicated matcher warp): both bit-identical,
// both SLOWER than this sequence (427 / 473 us vs 392 us at config 2) because the loss kernel already runs at
// ~0.95 of the measured HBM peak with 74 % of its issue slots busy — the matcher's instructions do not fit in the
// remaining slots and every in-kernel dependency (ticket, acquire, election) costs the streamers a round trip.
// See profiles/r01_notes.md ("single-launch training kernel") and the two ncu summaries next to it.
extern "C" size_t rn_train_loss_workspace_bytes(int N, int64_t A, int C) { return rn_loss_workspace_bytes(N, A, C); }

extern "C" int rn_train_loss(const float *logits, const float *bbox, const float *anchors, int64_t anchor_image_stride,
                             const float *gt_boxes, const int64_t *gt_labels, const int32_t *gt_off, int N,
                             int64_t gt_total, int64_t A, int C,
                             float fg_thr, float bg_thr, float alpha, float gamma, float beta, const float *weights_host,
                             float batch_div, int32_t *codes, int32_t *fg_count, float *out_image, float *out_total,
                             float *grad_logits, float *grad_bbox, void *workspace, size_t workspace_bytes,
                             rn_stream_t stream, const rn_exchange_t *exchange_host) {
    RN_CHECK_ARG(gt_labels && codes && fg_count, RN_E_BADARG, "rn_train_loss: null pointer");
    int rc = rn_match(anchors, A, anchor_image_stride, gt_boxes, gt_labels, gt_off, N, gt_total, fg_thr, bg_thr, nullptr,
                      codes, fg_count, stream);
    if (rc) return rc;
    return rn_loss(logits, bbox, anchors, anchor_image_stride, gt_boxes, gt_off, codes, fg_count, N, A, C, alpha, gamma, beta,
                   weights_host, batch_div, out_image, out_total, grad_logits, grad_bbox, workspace, workspace_bytes, stream,
                   exchange_host);
}

// ---- rn_train_detect: BOTH halves of the path on the same head outputs with ONE pass over the logits.  The training
// step of the reference reads cls_preds for the loss (losses.py:113-145) and, when detections of the same batch are
// wanted too (evaluation during training, models.py:160-243), a second time for the score threshold.  Here the loss
// kernel applies the post-processing's score filter to the registers it is streaming anyway and appends the survivors to
// the lazy NMS's candidate lists: per step 4*A*C bytes per image less HBM traffic (1.03 GB of 3.2 GB at config 2) and
// one kernel less; the NMS then runs behind the loss instead of beside it.  Same results as rn_train_loss +
// rn_postprocess(RN_PP_LAZY): the candidate SET is identical, the lists' order is irrelevant (the NMS ranks by key).
static size_t tdet_loss_bytes(int N, int64_t A, int C) { return (rn_loss_workspace_bytes(N, A, C) + 255) / 256 * 256; }

extern "C" size_t rn_train_detect_workspace_bytes(int N, int64_t A, int C, int64_t cand_capacity, int max_det) {
    return tdet_loss_bytes(N, A, C) + rn_postprocess_workspace_bytes(N, A, C, cand_capacity, max_det);
}

extern "C" int rn_train_detect(const float *logits, const float *bbox, const float *anchors, int64_t anchor_image_stride,
                               const float *gt_boxes, const int64_t *gt_labels, const int32_t *gt_off, int N, int64_t gt_total,
                               int64_t A, int C, float fg_thr, float bg_thr, float alpha, float gamma, float beta,
                               const float *weights_host, float batch_div, int32_t *codes, int32_t *fg_count, float *out_image,
                               float *out_total, float *grad_logits, float *grad_bbox, const int32_t *im_hw, float score_thr,
                               double nms_thr, int max_det, int pre_nms_topk, const int64_t *level_off_host, int num_levels,
                               int64_t cand_capacity, float *out_boxes, float *out_scores, int64_t *out_labels,
                               int32_t *out_count, int32_t *out_status, const float *out_ratio_hw, int out_format,
                               void *workspace, size_t workspace_bytes, rn_stream_t stream,
                               const rn_exchange_t *exchange_host, int phases) {
    RN_CHECK_ARG(gt_labels && codes && fg_count && workspace, RN_E_BADARG, "rn_train_detect: null pointer");
    RN_CHECK_ARG(phases >= 1 && phases <= 7, RN_E_BADARG, "rn_train_detect: phases is a mask of 1 (matcher), 4 (loss), 2 (NMS)");
    RN_CHECK_ARG(N > 0 && A > 0 && C > 0, RN_E_BADARG, "rn_train_detect: N, A, C must be positive");
    const size_t lb = tdet_loss_bytes(N, A, C);
    RN_CHECK_ARG(workspace_bytes >= lb, RN_E_WORKSPACE, "rn_train_detect: workspace too small");
    void *pp_ws = (char *)workspace + lb;
    cudaStream_t s = (cudaStream_t)stream;
    rnpp::LazySink sink;
    RN_CHECK_ARG(out_status && lb >= rn_loss_workspace_bytes(N, A, C), RN_E_BADARG, "rn_train_detect: null status");
    int rc = rnpp::lazy_begin(N, A, C, score_thr, max_det, pre_nms_topk, level_off_host, num_levels, cand_capacity, out_status,
                              pp_ws, workspace_bytes - lb, s, &sink, false);
    if (rc) return rc;
    if (phases & 1) {
        {   // one launch zeroes what the step's kernels accumulate into (instead of three memset nodes)
            double *base = (double *)workspace + (size_t)N * loss_chunks(A) * 2 + 2 * (size_t)N;
            prep_kernel<<<1, 256, 0, s>>>(fg_count, (unsigned *)(base + 2), (unsigned *)base, out_status, sink.img_count, N);
            RN_CHECK_LAUNCH("rn_train_detect/prep");
        }
        rc = rnpp::match_impl(anchors, A, anchor_image_stride, gt_boxes, gt_labels, gt_off, N, gt_total, fg_thr, bg_thr, nullptr,
                              codes, fg_count, stream, false);
        if (rc) return rc;
    }
    if (phases & 4) {
        rc = loss_impl(logits, bbox, anchors, anchor_image_stride, gt_boxes, gt_off, codes, fg_count, N, A, C, alpha, gamma, beta,
                       weights_host, batch_div, out_image, out_total, grad_logits, grad_bbox, workspace, lb, stream,
                       exchange_host, &sink);
        if (rc) return rc;
    }
    if (!(phases & 2)) return 0;
    return rnpp::lazy_end(bbox, anchors, anchor_image_stride, im_hw, N, A, C, score_thr, nms_thr, max_det, weights_host,
                          pre_nms_topk, level_off_host, num_levels, cand_capacity, out_boxes, out_scores, out_labels, out_count,
                          out_status, pp_ws, s, out_ratio_hw, out_format);
}

// ---- per-level NCHW layout -----------------------------------------------------------
// The reference's head permutes every level's conv output [N, na*C, H, W] to [N, H*W*na, C] and
// concatenates the levels (retinanet/layers.py:189-195, 253-259): a full extra read+write of the logits
// right before this path.  loss_levels_kernel consumes the conv outputs directly (index math only) and
// writes the gradients back in NCHW, so that pass disappears.  Element (n, a, c, y, x) of a level lives
// at ((n*na + a)*C + c)*H*W + y*W + x and belongs to anchor lvl_off + (y*W + x)*na + a.
// One CTA = 128 consecutive positions of one level x one cell-anchor index a x its C class planes; each warp
// walks class planes, LV_CU in flight, each lane owning 4 positions (a 128-bit load when H*W % 4 == 0, else
// coalesced scalar loads).  32 registers per thread -> 8 CTAs per SM (full occupancy), which is worth more here
// than loads in flight per thread: the CTAs are short (C/8 planes per warp) and many.
namespace {
#ifndef LV_PU
#define LV_PU 1      // 128-position chunks of one plane in flight per lane
#endif
#ifndef LV_CU
#define LV_CU 2      // channel planes in flight per warp (2 planes x 8 CTAs/SM measured best: 0.43 vs 0.56 ms with 4 x 4)
#endif
constexpr int LV_TILE = 128 * LV_PU;
constexpr int LV_BLOCK = 256;

struct LvlDesc {             // one pyramid level
    const float *cls;
    const float *box;
    float *gcls;
    float *gbox;
    long long lvl_off;       // anchor offset of the level
    int HW, na, tile_base, chunk_base, vec4;
};

struct LvlLossParams {
    LvlDesc lv[RN_MAX_LEVELS];
    int num_levels, total_tiles;
    const float4 *anchors;
    const float4 *gt;
    const int *gt_off;
    const int *codes;
    const int *fg_count;
    double *partials;
    long long A, anchor_stride;
    int C, chunks_total;
    float alpha, gamma, beta, batch_div;
    float4 wts;
};

// One CTA = (tile of 128 positions of one level, image n, cell-anchor index a): C class planes of 512 B.
// No shared memory and no barrier before the streaming loop: each lane keeps the ignore flags of its 4
// positions in registers (they are the same for all C planes) and the first plane loads are issued together
// with the code loads.  The plane walk treats every element as a negative; the one positive class element of a
// foreground anchor is patched afterwards, in the per-position regression loop.
template <int VEC, bool WANT_GRAD, bool GAMMA2>
__device__ __forceinline__ void loss_levels_body(const LvlLossParams &P, const LvlDesc &D, int tile, int n, int a) {
    const int p0 = tile * LV_TILE;
    const int np = min(LV_TILE, D.HW - p0);
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const long long anchor0 = D.lvl_off + (long long)p0 * D.na + a;   // anchor of position p: anchor0 + p*na
    const int *codes = P.codes + (long long)n * P.A + anchor0;
    // per-lane constants of the 4 positions this lane owns in every class plane
    bool ign[LV_PU][4];        // ignore anchor (or out of range): logit replaced by -100 -> p = sp = grad = 0 exactly
    bool ok[LV_PU];            // lane's vector lies inside the tile (VEC == 4)
#pragma unroll
    for (int pu = 0; pu < LV_PU; ++pu) {
        ok[pu] = pu * 128 + lane * 4 < np;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int p = VEC == 4 ? pu * 128 + lane * 4 + k : pu * 128 + k * 32 + lane;
            const int cd = p < np ? __ldg(codes + (long long)p * D.na) : -2;
            ign[pu][k] = cd == -2;
        }
    }
    const int F = P.fg_count[n];
    const float inv = 1.0f / (fmaxf((float)F, 1.0f) * P.batch_div);
    const float neg_gscale = P.alpha * inv;

    float acc_neg = 0.0f, acc_pos = 0.0f;
    const long long plane_stride = (long long)(LV_BLOCK / 32) * D.HW;
    const long long first = ((long long)n * D.na + a) * P.C * D.HW + (long long)warp * D.HW + p0;
    const float *src = D.cls + first;
    float *dst = WANT_GRAD ? D.gcls + first : nullptr;
    for (int c0 = warp; c0 < P.C; c0 += (LV_BLOCK / 32) * LV_CU, src += LV_CU * plane_stride, dst += WANT_GRAD ? LV_CU * plane_stride : 0) {
        float v[LV_CU][LV_PU][4];
#pragma unroll
        for (int cu = 0; cu < LV_CU; ++cu) {
            const bool live = c0 + cu * (LV_BLOCK / 32) < P.C;         // warp-uniform
            const float *plane = src + cu * plane_stride;
#pragma unroll
            for (int pu = 0; pu < LV_PU; ++pu) {
#pragma unroll
                for (int k = 0; k < 4; ++k) v[cu][pu][k] = -100.0f;
                if (live) {
                    if (VEC == 4) {
                        if (ok[pu]) {
                            const float4 q = rn::ld_stream_f4((const float4 *)(plane + pu * 128 + lane * 4));
                            v[cu][pu][0] = q.x; v[cu][pu][1] = q.y; v[cu][pu][2] = q.z; v[cu][pu][3] = q.w;
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int p = pu * 128 + k * 32 + lane;
                            if (p < np) v[cu][pu][k] = __ldg(plane + p);
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int cu = 0; cu < LV_CU; ++cu) {
            const int c = c0 + cu * (LV_BLOCK / 32);
            if (c >= P.C) continue;                                   // warp-uniform
#pragma unroll
            for (int pu = 0; pu < LV_PU; ++pu) {
                float x[4], g[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) x[k] = ign[pu][k] ? -100.0f : v[cu][pu][k];
                bool mid = true;
#pragma unroll
                for (int k = 0; k < 4; ++k) mid = mid && (x[k] <= kMidX - 1.0f);       // false for NaN
                float local = 0.0f;
                if (__all_sync(0xffffffffu, mid)) {            // every element as a negative; positives are patched after the walk
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        float wp;
                        focal_neg_mid<WANT_GRAD, GAMMA2>(x[k], P.gamma, local, wp);
                        g[k] = wp * neg_gscale;
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        float pk, spk;
                        sigmoid_softplus<false>(x[k] + 1.0f, pk, spk);
                        const float w = pow_gamma<GAMMA2>(pk, P.gamma);
                        local = fmaf(w, spk, local);
                        g[k] = w * pk * neg_gscale;
                    }
                }
                acc_neg += local;
                if (WANT_GRAD) {
                    float *gplane = dst + cu * plane_stride;
                    if (VEC == 4) {
                        if (ok[pu]) rn::st_stream_f4((float4 *)(gplane + pu * 128 + lane * 4), make_float4(g[0], g[1], g[2], g[3]));
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int p = pu * 128 + k * 32 + lane;
                            if (p < np) gplane[p] = g[k];
                        }
                    }
                }
            }
        }
    }

    // ---- per position: regression (box channel (a*4 + k) plane, stride HW) and, for a foreground anchor, the patch
    // of its ONE positive class element — its logit is re-read (L2), the negative term the walk added is taken back
    // and the gradient element the walk wrote is overwritten (hence the barrier) ----
    if (WANT_GRAD) __syncthreads();
    float reg = 0.0f;
    for (int p = t; p < np; p += LV_BLOCK) {
        const int cd = __ldg(codes + (long long)p * D.na);
        const long long b0 = ((long long)(n * D.na + a) * 4) * D.HW + p0 + p;
        float gr[4] = {0.f, 0.f, 0.f, 0.f};
        if (cd >= 0) {
            if ((cd >> 20) < P.C) {               // a label outside 1..C matches no class plane (as in loss_kernel)
                const long long e = (((long long)n * D.na + a) * P.C + (cd >> 20)) * D.HW + p0 + p;
                const float xx = __ldg(D.cls + e) + 1.0f;
                float pp, sp;
                sigmoid_softplus<false>(xx, pp, sp);
                const float wn = pow_gamma<GAMMA2>(pp, P.gamma);
                const float wp = pow_gamma<GAMMA2>(1.0f - pp, P.gamma) * (1.0f - P.alpha);
                acc_pos += wp * (sp - xx) - P.alpha * wn * sp;
                if (WANT_GRAD) D.gcls[e] = wp * (pp - 1.0f) * inv;
            }
            const long long anchor = anchor0 + (long long)p * D.na;
            const float4 gtb = P.gt[P.gt_off[n] + (cd & 0xFFFFF)];
            const float4 an = P.anchors[(long long)n * P.anchor_stride + anchor];
            const float4 tt = rn::encode_box(gtb, an, P.wts);
            const float tv[4] = {tt.x, tt.y, tt.z, tt.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float d = __ldg(D.box + b0 + (long long)k * D.HW) - tv[k];
                const float nabs = fabsf(d);
                if (P.beta < 1e-5f) {
                    reg += nabs;
                    gr[k] = d > 0.0f ? inv : (d < 0.0f ? -inv : 0.0f);
                } else if (nabs < P.beta) {
                    reg += 0.5f * nabs * nabs / P.beta;
                    gr[k] = d / P.beta * inv;
                } else {
                    reg += nabs - 0.5f * P.beta;
                    gr[k] = d > 0.0f ? inv : -inv;
                }
            }
        }
        if (WANT_GRAD) {
#pragma unroll
            for (int k = 0; k < 4; ++k) D.gbox[b0 + (long long)k * D.HW] = gr[k];
        }
    }

    double s_neg = acc_neg, s_pos = acc_pos, s_reg = reg;
    block_sum3(s_neg, s_pos, s_reg);
    if (t == 0) {
        double *o = P.partials + ((long long)n * P.chunks_total + D.chunk_base + (long long)tile * D.na + a) * 2;
        o[0] = (double)P.alpha * s_neg + s_pos;
        o[1] = s_reg;
    }
}

// All pyramid levels in ONE launch: blockIdx.x enumerates the tiles of every level (the small P5-P7 levels
// would otherwise be latency-bound launches of their own), blockIdx.y = image, blockIdx.z = cell anchor.
#ifndef LV_MINB
#define LV_MINB 8    // resident CTAs per SM asked of ptxas: 32 registers/thread, full occupancy (8 bytes of spills)
#endif
template <bool WANT_GRAD, bool GAMMA2>
__global__ void __launch_bounds__(LV_BLOCK, LV_MINB) loss_levels_kernel(const __grid_constant__ LvlLossParams P) {
    int l = 0;
#pragma unroll
    for (int k = 1; k < RN_MAX_LEVELS; ++k)
        if (k < P.num_levels && (int)blockIdx.x >= P.lv[k].tile_base) l = k;
    const LvlDesc &D = P.lv[l];
    const int a = blockIdx.z;
    if (a >= D.na) return;                                           // levels may differ in anchors per cell
    const int tile = blockIdx.x - D.tile_base;
    if (D.vec4) loss_levels_body<4, WANT_GRAD, GAMMA2>(P, D, tile, blockIdx.y, a);
    else loss_levels_body<1, WANT_GRAD, GAMMA2>(P, D, tile, blockIdx.y, a);
}

inline int lv_tiles(int HW) { return (HW + LV_TILE - 1) / LV_TILE; }
}  // namespace

extern "C" size_t rn_loss_levels_workspace_bytes(int N, const int32_t *level_desc_host, int num_levels) {
    if (N <= 0 || !level_desc_host || num_levels <= 0) return 16;
    size_t chunks = 0;
    for (int l = 0; l < num_levels; ++l)
        chunks += (size_t)lv_tiles(level_desc_host[3 * l] * level_desc_host[3 * l + 1]) * (size_t)level_desc_host[3 * l + 2];
    return ((size_t)N * chunks * 2 + (size_t)N * 2 + 2) * sizeof(double);
}

extern "C" int rn_loss_levels(const float *const *cls_levels_host, const float *const *bbox_levels_host,
                              const int32_t *level_desc_host, int num_levels, const float *anchors,
                              int64_t anchor_image_stride, const float *gt_boxes, const int32_t *gt_off,
                              const int32_t *codes, const int32_t *fg_count, int N, int64_t A, int C, float alpha,
                              float gamma, float beta, const float *weights_host, float batch_div, float *out_image,
                              float *out_total, float *const *grad_cls_levels_host, float *const *grad_bbox_levels_host,
                              void *workspace, size_t workspace_bytes, rn_stream_t stream,
                              const rn_exchange_t *exchange_host) {
    ExchangeDev X;
    RN_CHECK_ARG(make_exchange(exchange_host, X), RN_E_BADARG, "rn_loss_levels: malformed exchange descriptor");
    RN_CHECK_ARG(cls_levels_host && bbox_levels_host && level_desc_host && anchors && gt_off && codes && fg_count &&
                     out_total && weights_host, RN_E_BADARG, "rn_loss_levels: null pointer");
    RN_CHECK_ARG(num_levels >= 1 && num_levels <= RN_MAX_LEVELS, RN_E_TOOLARGE, "rn_loss_levels: bad num_levels %d", num_levels);
    RN_CHECK_ARG(N > 0 && A > 0 && C > 0 && N <= 65535 && C <= 2047, RN_E_BADARG, "rn_loss_levels: bad N/A/C");
    RN_CHECK_ARG((grad_cls_levels_host == nullptr) == (grad_bbox_levels_host == nullptr), RN_E_BADARG,
                 "rn_loss_levels: gradient level arrays must be given together");
    RN_CHECK_ARG(batch_div > 0.0f, RN_E_BADARG, "rn_loss_levels: batch_div must be positive");
    RN_CHECK_ARG(workspace && workspace_bytes >= rn_loss_levels_workspace_bytes(N, level_desc_host, num_levels),
                 RN_E_WORKSPACE, "rn_loss_levels: workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    const bool want = grad_cls_levels_host != nullptr;
    LvlLossParams P;
    memset(&P, 0, sizeof(P));
    int tiles = 0, chunks = 0, max_na = 1;
    long long lvl_off = 0;
    for (int l = 0; l < num_levels; ++l) {
        const int32_t *d = level_desc_host + 3 * l;
        RN_CHECK_ARG(d[0] >= 0 && d[1] >= 0 && d[2] >= 1 && d[2] <= 64, RN_E_BADARG,
                     "rn_loss_levels: bad level %d descriptor {%d,%d,%d}", l, d[0], d[1], d[2]);
        const int HW = d[0] * d[1];
        RN_CHECK_ARG(HW == 0 || (cls_levels_host[l] && bbox_levels_host[l]), RN_E_BADARG, "rn_loss_levels: null level %d", l);
        LvlDesc &D = P.lv[l];
        D.cls = cls_levels_host[l]; D.box = bbox_levels_host[l];
        D.gcls = want ? grad_cls_levels_host[l] : nullptr; D.gbox = want ? grad_bbox_levels_host[l] : nullptr;
        D.lvl_off = lvl_off; D.HW = HW; D.na = d[2]; D.tile_base = tiles; D.chunk_base = chunks;
        D.vec4 = (HW % 4 == 0) && (((uintptr_t)D.cls & 15) == 0) && (!want || ((uintptr_t)D.gcls & 15) == 0);
        tiles += lv_tiles(HW);
        chunks += lv_tiles(HW) * d[2];
        lvl_off += (long long)HW * d[2];
        max_na = d[2] > max_na ? d[2] : max_na;
    }
    RN_CHECK_ARG(lvl_off == A, RN_E_BADARG, "rn_loss_levels: levels hold %lld anchors, A = %lld", lvl_off, (long long)A);
    P.num_levels = num_levels; P.total_tiles = tiles;
    P.anchors = (const float4 *)anchors; P.gt = (const float4 *)gt_boxes; P.gt_off = gt_off; P.codes = codes;
    P.fg_count = fg_count; P.partials = (double *)workspace; P.A = A; P.anchor_stride = anchor_image_stride;
    P.C = C; P.chunks_total = chunks; P.alpha = alpha; P.gamma = gamma; P.beta = beta; P.batch_div = batch_div;
    P.wts = make_float4(weights_host[0], weights_host[1], weights_host[2], weights_host[3]);
    if (tiles > 0) {
        dim3 grid((unsigned)tiles, (unsigned)N, (unsigned)max_na);
        const bool g2 = gamma == 2.0f;
        if (want) { if (g2) loss_levels_kernel<true, true><<<grid, LV_BLOCK, 0, s>>>(P); else loss_levels_kernel<true, false><<<grid, LV_BLOCK, 0, s>>>(P); }
        else      { if (g2) loss_levels_kernel<false, true><<<grid, LV_BLOCK, 0, s>>>(P); else loss_levels_kernel<false, false><<<grid, LV_BLOCK, 0, s>>>(P); }
        RN_CHECK_LAUNCH("rn_loss_levels");
    }
    {
        double *tail = (double *)workspace + (size_t)N * chunks * 2;
        cudaError_t e = cudaMemsetAsync(tail + 2 * (size_t)N, 0, sizeof(unsigned), s);
        if (e != cudaSuccess) { rn_set_error("rn_loss_levels: memset failed: %s", cudaGetErrorString(e)); return (int)e; }
        loss_finalize_kernel<<<N, FIN_BLOCK, 0, s>>>((const double *)workspace, fg_count, N, chunks, batch_div, out_image,
                                                     out_total, tail, X);
    }
    RN_CHECK_LAUNCH("rn_loss_levels/finalize");
    return 0;
}

// ---- dense element-wise losses (API parity with RetinaNetLosses.focal_loss / smooth_l1_loss) ----------
// focal_loss(clas_pred, clas_tgt), retinanet/losses.py:29-47, for arbitrary float targets:
//   p = sigmoid(x) (detached); w = (t(1-p) + (1-t)p)^gamma * ((1-t)alpha + t(1-alpha));
//   loss = sum w * ((1-t) x - log_sigmoid(x));  dloss/dx = w * (p - t)
// smooth_l1_loss(input, target), retinanet/losses.py:19-27.
// HBM-bound streaming reductions: grid-stride 128-bit loads, per-CTA fp64 partials, one-block finish.
namespace {
constexpr int DENSE_BLOCK = 256;
constexpr int DENSE_GRID = RN_SM_COUNT_B200 * 8;

template <bool FOCAL>
__global__ void __launch_bounds__(DENSE_BLOCK) dense_loss_kernel(const float *__restrict__ x, const float *__restrict__ t,
                                                                 long long n, float p0, float p1, float *__restrict__ grad,
                                                                 double *__restrict__ partials) {
    float acc = 0.0f;
    const long long stride = (long long)gridDim.x * DENSE_BLOCK;
    for (long long i = (long long)blockIdx.x * DENSE_BLOCK + threadIdx.x; i < n; i += stride) {
        const float xv = __ldg(x + i), tv = __ldg(t + i);
        float l, g;
        if (FOCAL) {
            const float alpha = p0, gamma = p1;
            const float e = expf(-fabsf(xv));
            const float r = __fdiv_rn(1.0f, 1.0f + e);
            const float p = xv >= 0.0f ? r : e * r;
            float w = tv * (1.0f - p) + (1.0f - tv) * p;
            w = (gamma == 2.0f ? w * w : (w > 0.0f ? powf(w, gamma) : (gamma == 0.0f ? 1.0f : 0.0f)));
            w *= (1.0f - tv) * alpha + tv * (1.0f - alpha);
            const float logsig = fminf(xv, 0.0f) - log1pf(e);
            l = w * ((1.0f - tv) * xv - logsig);
            g = w * (p - tv);
        } else {
            const float beta = p0;
            const float d = xv - tv, a = fabsf(d);
            if (beta < 1e-5f) { l = a; g = d > 0.0f ? 1.0f : (d < 0.0f ? -1.0f : 0.0f); }
            else if (a < beta) { l = 0.5f * a * a / beta; g = d / beta; }
            else { l = a - 0.5f * beta; g = d > 0.0f ? 1.0f : -1.0f; }
        }
        acc += l;
        if (grad) grad[i] = g;
    }
    double a = acc, b = 0.0, c = 0.0;
    block_sum3(a, b, c);
    if (threadIdx.x == 0) partials[blockIdx.x] = a;
}

__global__ void __launch_bounds__(256) dense_finish_kernel(const double *__restrict__ partials, int n, float *__restrict__ out) {
    __shared__ double s[8];
    double a = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) a += partials[i];
    a = rn::warp_sum(a);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        double r = 0.0;
        for (int i = 0; i < 8; ++i) r += s[i];
        out[0] = (float)r;
    }
}

int dense_launch(bool focal, const float *x, const float *t, int64_t n, float p0, float p1, float *out, float *grad,
                 void *workspace, size_t workspace_bytes, rn_stream_t stream, const char *name) {
    RN_CHECK_ARG(n >= 0 && out, RN_E_BADARG, "%s: bad argument", name);
    RN_CHECK_ARG(n == 0 || (x && t), RN_E_BADARG, "%s: null input", name);
    RN_CHECK_ARG(workspace && workspace_bytes >= DENSE_GRID * sizeof(double), RN_E_WORKSPACE, "%s: workspace too small", name);
    cudaStream_t s = (cudaStream_t)stream;
    const int grid = (int)std::max<long long>(1, std::min<long long>(DENSE_GRID, (n + DENSE_BLOCK - 1) / DENSE_BLOCK));
    if (focal) dense_loss_kernel<true><<<grid, DENSE_BLOCK, 0, s>>>(x, t, n, p0, p1, grad, (double *)workspace);
    else dense_loss_kernel<false><<<grid, DENSE_BLOCK, 0, s>>>(x, t, n, p0, p1, grad, (double *)workspace);
    RN_CHECK_LAUNCH(name);
    dense_finish_kernel<<<1, 256, 0, s>>>((const double *)workspace, grid, out);
    RN_CHECK_LAUNCH(name);
    return 0;
}
}  // namespace

extern "C" size_t rn_dense_loss_workspace_bytes(void) { return DENSE_GRID * sizeof(double); }

extern "C" int rn_focal_loss_dense(const float *logits, const float *targets, int64_t n, float alpha, float gamma,
                                   float *out_sum, float *grad, void *workspace, size_t workspace_bytes,
                                   rn_stream_t stream) {
    return dense_launch(true, logits, targets, n, alpha, gamma, out_sum, grad, workspace, workspace_bytes, stream,
                        "rn_focal_loss_dense");
}

extern "C" int rn_smooth_l1_dense(const float *input, const float *target, int64_t n, float beta, float *out_sum,
                                  float *grad, void *workspace, size_t workspace_bytes, rn_stream_t stream) {
    return dense_launch(false, input, target, n, beta, 0.0f, out_sum, grad, workspace, workspace_bytes, stream,
                        "rn_smooth_l1_dense");
}

extern "C" int rn_scale_by_device_scalar(float *buf, int64_t n, const float *scale, rn_stream_t stream) {
    RN_CHECK_ARG(buf && scale && n >= 0, RN_E_BADARG, "rn_scale_by_device_scalar: bad argument");
    if (n == 0) return 0;
    RN_CHECK_ARG(((uintptr_t)buf & 15) == 0, RN_E_BADARG, "rn_scale_by_device_scalar: buffer not 16-byte aligned");
    scale_kernel<<<RN_SM_COUNT_B200 * 8, 256, 0, (cudaStream_t)stream>>>(buf, (long long)n, scale);
    RN_CHECK_LAUNCH("rn_scale_by_device_scalar");
    return 0;
}
