// Streaming kernels for NCHW-contiguous feature maps: every (n,c) plane is M = H*W contiguous
// elements = nvec vector accesses.  The tensor is swept as one flat space of planes*nvec
// vectors (see plan.h): a CTA owns a contiguous slice cut into pieces at plane boundaries
// (CTA mode), or a warp owns whole small planes (warp mode).  Per piece the group reduces once;
// planes shared by several CTAs are merged from per-CTA partials in a fixed order by whoever
// arrives last (weighted tickets, no float atomics), so results are run-to-run deterministic.
#pragma once
#include "common.cuh"
#include "plan.h"

#define MS_HD __host__ __device__ __forceinline__

namespace ms {

// Style tables are addressed as table[row * ld + c]: `ld` (>= C) lets mu and sig live interleaved in
// one [N_global, 2C] buffer, which is what a single all-gather of per-rank [N, 2C] blocks produces.
struct TableRef {
    int C;            // channels
    int ld;           // floats between consecutive rows of mu_all / sig_all
    int row_offset;   // first global row owned by this call
    __device__ __forceinline__ int64_t at(int64_t plane) const {     // local plane index -> offset in mu_all/sig_all
        const int64_t n = plane / C;
        return (row_offset + n) * ld + (plane - n * C);
    }
};

// The style arithmetic of maxstyle.py:172-185 for one (n,c): returns scale = A/sig and shift = B.
// Shared by tables_kernel and the fused forward so the two forward paths agree to the last bit.
__device__ __forceinline__ void style_coeffs(float sg, float m, float sg_partner, float mu_partner, bool mix, bool no_noise,
                                             float lmda_raw, float gamma_noise, float beta_noise, float gs, float bs,
                                             float& scale, float& shift, bool clamp = true) {
    float sg_mix = sg, mu_mix = m;
    if (mix) {
        // MaxStyle clamps the mixing weight (maxstyle.py:173); MixStyle does not (mixstyle.py:95-96: a fixed lmda > 1
        // or a Gaussian-sampled one extrapolates) -- MAXSTYLE_NO_CLAMP
        const float l = clamp ? fminf(fmaxf(lmda_raw, 0.f), 1.f) : lmda_raw;
        sg_mix = sg * (1.f - l) + sg_partner * l;
        mu_mix = m * (1.f - l) + mu_partner * l;
    }
    float A = sg_mix, B = mu_mix;
    if (!no_noise) {
        A = sg_mix + gamma_noise * gs;
        B = mu_mix + beta_noise * bs;
    }
    scale = A / sg;
    shift = B;
}

// Geometry of one sweep (device copy of Plan + per-call options).
struct Sweep {
    int64_t M;        // elements per plane
    int64_t nvec;     // vectors per plane
    int64_t planes;   // N*C
    int64_t total;    // planes*nvec
    int64_t per;      // CTA mode: vectors per CTA; warp mode: planes per warp
    int slots;        // partial slots per plane
    int reverse;      // walk the slice from its high end (so that the kernel that follows a forward
                      // sweep over the same tensor meets the lines that are still in L2 first)
    int in_policy;    // L2 policy of the loads of the tensor that other kernels of the layer re-read (x)
    int io_policy;    // L2 policy of everything else (dy loads, y / dx stores)
    int pre_op;       // activation applied to x as it is loaded (kPreNone / kPreLeakyRelu / kPreSigmoid): x = act(z)
    float pre_param;  // negative slope of the leaky relu
    int vslices;      // backward, CTA mode: number of slices of `per` vectors; the CTAs take them from a ticket counter (0: one per CTA)
};

template <int G> struct GroupIdx {
    static constexpr int kPerBlock = kThreads / G;
    __device__ static __forceinline__ int lane() { return G > 32 ? threadIdx.x : (threadIdx.x & 31); }
    __device__ static __forceinline__ int64_t index() {
        return (int64_t)blockIdx.x * kPerBlock + (G > 32 ? 0 : (threadIdx.x >> 5));
    }
};

struct Piece {
    int64_t plane;    // plane index
    int v0, v1;       // vector range inside the plane (planes hold < 2^31 vectors)
};

// The pieces of one group's slice, in sweep order.
template <int G> struct PieceIter {
    int64_t lo, hi, nvec;
    bool reverse;
    // gi: index of the group (CTA index in CTA mode, global warp index in warp mode)
    MS_HD PieceIter(const Sweep& g, int64_t gi) {
        const int64_t unit = G > 32 ? 1 : g.nvec;             // warp mode: the slice is whole planes
        lo = gi * g.per * unit;
        hi = lo + g.per * unit < g.total ? lo + g.per * unit : g.total;
        nvec = g.nvec;
        reverse = g.reverse != 0;
    }
    MS_HD bool next(Piece& p) {
        if (lo >= hi) return false;
        if (!reverse) {
            p.plane = lo / nvec;
            const int64_t v0 = lo - p.plane * nvec;
            const int64_t v1 = v0 + (hi - lo) < nvec ? v0 + (hi - lo) : nvec;
            p.v0 = (int)v0; p.v1 = (int)v1;
            lo += v1 - v0;
        } else {
            p.plane = (hi - 1) / nvec;
            const int64_t v1 = hi - p.plane * nvec;
            const int64_t v0 = v1 - (hi - lo) > 0 ? v1 - (hi - lo) : 0;
            p.v0 = (int)v0; p.v1 = (int)v1;
            hi -= v1 - v0;
        }
        return true;
    }
};

// A piece [v0, v1) is walked in batches of G*VPT vectors: `full` whole batches (no bounds checks in
// the hot loop) and one ragged batch at the far end of the walk.  batch_begin() gives the first
// vector of batch i for this sweep direction; the ragged batch is i == full.
template <int G, int VPT> struct Batches {
    static constexpr int kStep = G * VPT;
    int v0, v1, full, rem;
    bool reverse;
    MS_HD Batches(const Piece& p, bool rev) : v0(p.v0), v1(p.v1), reverse(rev) {
        const int len = p.v1 - p.v0;
        full = len / kStep;
        rem = len - full * kStep;
    }
    MS_HD int begin(int i) const { return reverse ? v1 - (i + 1) * kStep : v0 + i * kStep; }
    // ragged batch: vectors [lo, hi)
    MS_HD int ragged_lo() const { return reverse ? v0 : v0 + full * kStep; }
    MS_HD int ragged_hi() const { return reverse ? v0 + rem : v1; }
};

// Which CTAs share a plane (CTA mode): CTA b covers vectors [b*per, (b+1)*per).
struct PlaneShare {
    int64_t first;    // first CTA touching the plane
    int count;        // number of CTAs touching it
};
MS_HD PlaneShare plane_share(const Sweep& g, int64_t plane) {
    PlaneShare s;
    s.first = (plane * g.nvec) / g.per;
    s.count = (int)(((plane + 1) * g.nvec - 1) / g.per - s.first) + 1;
    return s;
}

// ---------------------------------------------------------------------------------------------
// Kernel 1: instance statistics.  Replaces x.mean(dim=[2,3]) / x.var(dim=[2,3]) / sqrt(var+eps)
// of the reference (src/advanced/maxstyle.py:157-159) with ONE read of x.
// Numerics: every value is first shifted by K = the plane's first element (exact for data whose
// spread is small against its mean -- the case where fp32 moments lose digits), then per thread
// batches of VPT vectors are reduced two-pass in registers (sum -> mean -> squared deviations) and
// folded into a running (n, mean, M2) with the Chan/Welford merge; warp shuffles, shared memory
// across the CTA's warps, and for planes shared by several CTAs a fixed-order merge of the
// per-CTA partials by the last arriver.
// ---------------------------------------------------------------------------------------------
template <typename T, int VEC, int G, int VPT>
__global__ void __launch_bounds__(kThreads, kBlocksPerSM)
stats_nchw_kernel(const T* __restrict__ x, float* __restrict__ mu, float* __restrict__ sig, TableRef tr,
                  float4* __restrict__ partials, unsigned long long* __restrict__ plane_tickets, Sweep g, float eps) {
    __shared__ Scratch scratch;
    const int t = GroupIdx<G>::lane();
    const uint64_t pol = make_policy(g.in_policy);
    const float inv_m1 = 1.0f / (float)(g.M - 1);
    constexpr int kTail = VPT > 1 ? VPT / 2 : 1;
    PieceIter<G> it(g, GroupIdx<G>::index());
    Piece pc;
    while (it.next(pc)) {
        const T* base = x + pc.plane * g.M;
        const float K = pre_apply(to_f32<T>(__ldg(base)), g.pre_op, g.pre_param);
        Moments acc{0.f, 0.f, 0.f};
        const int len = pc.v1 - pc.v0;
        const Batches<G, VPT> bt(pc, g.reverse != 0);
        for (int i = 0; i < bt.full; ++i) {            // whole batches: constant count, no bounds checks
            const T* p = base + (int64_t)(bt.begin(i) + t) * VEC;
            float val[VPT][VEC];
#pragma unroll
            for (int j = 0; j < VPT; ++j) { Vec<T, VEC>::load(p + (int64_t)j * G * VEC, val[j], pol); pre_apply_vec(val[j], g.pre_op, g.pre_param); }
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < VPT; ++j)
#pragma unroll
                for (int k = 0; k < VEC; ++k) { val[j][k] -= K; s += val[j][k]; }
            Moments b;
            b.n = (float)(VPT * VEC);
            b.mean = s * (1.0f / (float)(VPT * VEC));
            float q = 0.f;
#pragma unroll
            for (int j = 0; j < VPT; ++j)
#pragma unroll
                for (int k = 0; k < VEC; ++k) { const float d = val[j][k] - b.mean; q = fmaf(d, d, q); }
            b.m2 = q;
            acc = merge_fast(acc, b);
        }
        if (bt.rem) {                                  // ragged end of the piece: < G*VPT vectors, kTail loads in flight
            const int hi = bt.ragged_hi();
            for (int lo = bt.ragged_lo() + t; lo < hi; lo += kTail * G) {
                float val[kTail][VEC];
#pragma unroll
                for (int j = 0; j < kTail; ++j)
                    if (lo + j * G < hi) { Vec<T, VEC>::load(base + (int64_t)(lo + j * G) * VEC, val[j], pol); pre_apply_vec(val[j], g.pre_op, g.pre_param); }
#pragma unroll
                for (int j = 0; j < kTail; ++j) {
                    if (lo + j * G < hi) {             // each vector is its own mini-batch
                        float s = 0.f;
#pragma unroll
                        for (int k = 0; k < VEC; ++k) { val[j][k] -= K; s += val[j][k]; }
                        Moments b;
                        b.n = (float)VEC;
                        b.mean = s * (1.0f / (float)VEC);
                        float q = 0.f;
#pragma unroll
                        for (int k = 0; k < VEC; ++k) { const float d = val[j][k] - b.mean; q = fmaf(d, d, q); }
                        b.m2 = q;
                        acc = merge_fast(acc, b);
                    }
                }
            }
        }
        const Moments tot = group_merge<G>(acc, scratch);
        if (len == g.nvec) {                           // the piece is the whole plane
            if (t == 0) {
                const int64_t o = tr.at(pc.plane);
                mu[o] = K + tot.mean;
                sig[o] = sqrtf(tot.m2 * inv_m1 + eps);
            }
        } else if (t < 32) {                           // CTA mode only: warp 0 publishes, the last arriver merges
            const PlaneShare sh = plane_share(g, pc.plane);
            float4* slot = partials + pc.plane * g.slots;
            bool last = false;
            if (t == 0) {
                slot[blockIdx.x - sh.first] = make_float4(tot.n, tot.mean, tot.m2, 0.f);
                last = ticket_add(&plane_tickets[pc.plane], (unsigned long long)len, (unsigned long long)g.nvec);
            }
            last = __shfl_sync(0xffffffffu, (int)last, 0) != 0;
            __syncwarp();
            if (last) {
                Moments m{0.f, 0.f, 0.f};
                for (int k0 = 0; k0 < sh.count; k0 += 32) {           // fixed order: slot index, then a shuffle tree
                    Moments p{0.f, 0.f, 0.f};
                    if (k0 + t < sh.count) {
                        const float4 v = __ldcg(&slot[k0 + t]);
                        p = Moments{v.x, v.y, v.z};
                    }
                    m = merge(m, warp_merge(p));
                }
                if (t == 0) {
                    const int64_t o = tr.at(pc.plane);
                    mu[o] = K + m.mean;
                    sig[o] = sqrtf(m.m2 * inv_m1 + eps);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Kernel 2: apply.  y = (x - mu) * scale + shift with scale = A/sig, shift = B, i.e. the
// normalise + mix + perturb + affine chain of maxstyle.py:161,172-185 folded into one FMA per
// element (the [N,C] tables come from tables_kernel).
// ---------------------------------------------------------------------------------------------
template <typename T, int VEC, int G, int VPT>
__global__ void __launch_bounds__(kThreads, kBlocksPerSM)
apply_nchw_kernel(const T* __restrict__ x, T* __restrict__ y, const float* __restrict__ mu, TableRef tr,
                  const float* __restrict__ scale, const float* __restrict__ shift, Sweep g, unsigned int* __restrict__ ymin,
                  unsigned int* __restrict__ ymax) {
    const int t = GroupIdx<G>::lane();
    const uint64_t pol_in = make_policy(g.in_policy), pol_out = make_policy(g.io_policy);
    PieceIter<G> it(g, GroupIdx<G>::index());
    Piece pc;
    while (it.next(pc)) {
        const T* src = x + pc.plane * g.M;
        T* dst = y + pc.plane * g.M;
        const float m = __ldg(mu + tr.at(pc.plane)), a = __ldg(scale + pc.plane), b = __ldg(shift + pc.plane);
        float lo_y = INFINITY, hi_y = -INFINITY;
        const Batches<G, VPT> bt(pc, g.reverse != 0);
        for (int i = 0; i < bt.full; ++i) {
            const int64_t o = (int64_t)(bt.begin(i) + t) * VEC;
            float val[VPT][VEC];
#pragma unroll
            for (int j = 0; j < VPT; ++j) { Vec<T, VEC>::load(src + o + (int64_t)j * G * VEC, val[j], pol_in); pre_apply_vec(val[j], g.pre_op, g.pre_param); }
#pragma unroll
            for (int j = 0; j < VPT; ++j) {
#pragma unroll
                for (int k = 0; k < VEC; ++k) { val[j][k] = fmaf(val[j][k] - m, a, b); lo_y = fminf(lo_y, val[j][k]); hi_y = fmaxf(hi_y, val[j][k]); }
                Vec<T, VEC>::store(dst + o + (int64_t)j * G * VEC, val[j], pol_out);
            }
        }
        if (bt.rem) {
            const int lo = bt.ragged_lo() + t, hi = bt.ragged_hi();
            float val[VPT][VEC];
#pragma unroll
            for (int j = 0; j < VPT; ++j)
                if (lo + j * G < hi) { Vec<T, VEC>::load(src + (int64_t)(lo + j * G) * VEC, val[j], pol_in); pre_apply_vec(val[j], g.pre_op, g.pre_param); }
#pragma unroll
            for (int j = 0; j < VPT; ++j) {
                if (lo + j * G < hi) {
#pragma unroll
                    for (int k = 0; k < VEC; ++k) { val[j][k] = fmaf(val[j][k] - m, a, b); lo_y = fminf(lo_y, val[j][k]); hi_y = fmaxf(hi_y, val[j][k]); }
                    Vec<T, VEC>::store(dst + (int64_t)(lo + j * G) * VEC, val[j], pol_out);
                }
            }
        }
        if (ymin != nullptr) warp_minmax_publish(lo_y, hi_y, ymin + pc.plane, ymax + pc.plane);      // the values are rounded to T on store: exact for fp32
    }
}

// ---------------------------------------------------------------------------------------------
// Optimiser step arithmetic (shared by the backward epilogue and the stand-alone step kernel).
// Adam follows torch.optim.Adam's single-tensor update order (lerp, addcmul, bias corrections,
// addcdiv), the optimiser the reference caller uses (advanced_triplet_recon_segmentation_model.py
// :537,562).  Sign mode is BASELINE.json's sign-gradient step.
// ---------------------------------------------------------------------------------------------
struct StepArgs {
    int mode, maximize, update_noise, update_mix;
    double lr, beta1, beta2, eps;
    int t;
    int* step_dev;
    float *gamma_noise, *beta_noise, *lmda;
    float *gamma_m, *gamma_v, *beta_m, *beta_v, *lmda_m, *lmda_v;
};

struct StepCoef {
    float one_minus_b1, b2, one_minus_b2, step_size, bc2_sqrt, eps, lr;
};

// 1 - beta^t without cancellation: -expm1(t * log1p(-(1 - beta))), in fp32 (relative error ~1e-7;
// torch evaluates the same quantities in double on the host and rounds the results to fp32).
__device__ __forceinline__ float one_minus_pow(float one_minus_beta, int t) {
    return -expm1f((float)t * log1pf(-one_minus_beta));
}

__device__ __forceinline__ StepCoef step_coef(const StepArgs& s) {
    StepCoef c{};
    if (s.mode == 1) {
        const int t = s.step_dev ? (*(volatile int*)s.step_dev + 1) : s.t;
        c.one_minus_b1 = (float)(1.0 - s.beta1);
        c.b2 = (float)s.beta2;
        c.one_minus_b2 = (float)(1.0 - s.beta2);
        const float bc1 = one_minus_pow(c.one_minus_b1, t);
        const float bc2 = one_minus_pow(c.one_minus_b2, t);
        c.step_size = (float)s.lr / bc1;
        c.bc2_sqrt = sqrtf(bc2);
        c.eps = (float)s.eps;
    }
    c.lr = (float)s.lr;
    return c;
}

__device__ __forceinline__ void step_update(int mode, int maximize, const StepCoef& c, float g, float* p, float* m,
                                            float* v) {
    if (maximize) g = -g;
    if (mode == 1) {
        float mm = *m, vv = *v;
        mm = mm + c.one_minus_b1 * (g - mm);
        vv = vv * c.b2 + c.one_minus_b2 * g * g;
        *m = mm;
        *v = vv;
        const float denom = sqrtf(vv) / c.bc2_sqrt + c.eps;
        *p = *p - c.step_size * (mm / denom);
    } else if (mode == 2) {
        const float sgn = g > 0.f ? 1.f : (g < 0.f ? -1.f : 0.f);
        *p = *p - c.lr * sgn;
    }
}

// ---------------------------------------------------------------------------------------------
// Kernel 3: backward.  mu/sig are detached in the reference (maxstyle.py:160), so the autograd
// graph of :161-185 collapses to   dx = dy * A/sig,   dA = sum dy*(x-mu)/sig,   dB = sum dy
// per plane, followed by tiny per-sample reductions.  One sweep reads
// dy and x once and writes dx once (template DX=false when x does not require grad).  The group
// that completes a sample (weighted ticket over its C*nvec vectors) runs that sample's epilogue:
// parameter gradients, the channel reduction for d_lmda in fixed order (no float atomics) and
// the fused optimiser step.
// ---------------------------------------------------------------------------------------------
struct BwdTables {
    const float* mu_all;
    const float* sig_all;
    const float* scale;         // local rows: A/sig
    const int64_t* perm;
    const float* lmda;          // local rows
    const float* gamma_std;
    const float* beta_std;
    float* d_gamma;
    float* d_beta;
    float* d_lmda;
    int row_offset, N, C, flags, ld;
};

template <int G>
__device__ __forceinline__ void bwd_finalize_sample(int n, const BwdTables& tb, const StepArgs& st, const Sweep& g,
                                                    const float4* partials, int* done_counter, Scratch& scratch,
                                                    int fixed_count = -1) {
    // `fixed_count` >= 0: every channel of the sample has that many partials (NHWC: the sample, not the plane, is
    // what CTAs share); otherwise the count comes from the plane's share of the flat NCHW sweep.
    // One thread per channel: every channel's chain of dependent loads (partials -> tables -> Adam state)
    // runs in parallel with the others; the partials of a shared plane are summed in slot order.
    const int t = GroupIdx<G>::lane();
    const int C = tb.C;
    const int64_t row = (int64_t)tb.row_offset + n;
    const bool mix = tb.flags & 1, no_noise = tb.flags & 2;
    const int64_t prow = mix ? tb.perm[row] : row;
    const StepCoef coef = step_coef(st);
    const int ld = tb.ld;
    float lam_acc = 0.f, unused = 0.f;
    for (int c = t; c < C; c += G) {
        const int64_t plane = (int64_t)n * C + c;
        int count = 1;
        if (fixed_count >= 0) count = fixed_count;
        else if constexpr (G > 32) count = plane_share(g, plane).count;
        const float4* slot = partials + plane * g.slots;
        float s1 = 0.f, s2 = 0.f;
        for (int k = 0; k < count; ++k) {
            const float4 p = __ldcg(&slot[k]);
            s1 += p.x;
            s2 += p.y;
        }
        const float sg = tb.sig_all[row * ld + c], m = tb.mu_all[row * ld + c];
        const float dA = s2 / sg, dB = s1;
        const float gg = no_noise ? 0.f : dA * tb.gamma_std[c];
        const float gb = no_noise ? 0.f : dB * tb.beta_std[c];
        if (tb.d_gamma) tb.d_gamma[plane] = gg;
        if (tb.d_beta) tb.d_beta[plane] = gb;
        if (mix) lam_acc += dA * (tb.sig_all[prow * ld + c] - sg) + dB * (tb.mu_all[prow * ld + c] - m);
        if (st.mode != 0 && st.update_noise) {
            step_update(st.mode, st.maximize, coef, gg, st.gamma_noise + plane, st.gamma_m + plane, st.gamma_v + plane);
            step_update(st.mode, st.maximize, coef, gb, st.beta_noise + plane, st.beta_m + plane, st.beta_v + plane);
        }
    }
    group_sum2<G>(lam_acc, unused, scratch);
    if (t == 0) {
        float dl = 0.f;
        if (mix) {
            const float l = tb.lmda[n];
            dl = ((tb.flags & 8) || (l >= 0.f && l <= 1.f)) ? lam_acc : 0.f;      // clamp backward: closed interval
        }
        if (tb.d_lmda) tb.d_lmda[n] = dl;
        if (st.mode != 0 && st.update_mix && mix)
            step_update(st.mode, st.maximize, coef, dl, st.lmda + n, st.lmda_m + n, st.lmda_v + n);
        if (st.mode != 0 && st.step_dev) {                     // device-side step counter
            __threadfence();
            if (atomicAdd(done_counter, 1) == tb.N - 1) {
                *done_counter = 0;
                *st.step_dev = *(volatile int*)st.step_dev + 1;
            }
        }
    }
}

template <typename T, int VEC, int G, int VPT, bool DX>
__global__ void __launch_bounds__(kThreads, kBlocksPerSM)
bwd_nchw_kernel(const T* __restrict__ dy, const T* __restrict__ x, T* __restrict__ dx, float4* __restrict__ partials,
                unsigned long long* __restrict__ sample_tickets, int* __restrict__ done_counter, Sweep g, BwdTables tb,
                StepArgs st) {
    __shared__ Scratch scratch;
    const int t = GroupIdx<G>::lane();
    const uint64_t pol_x = make_policy(g.in_policy), pol_io = make_policy(g.io_policy);
    const unsigned long long sample_total = (unsigned long long)tb.C * (unsigned long long)g.nvec;
    // CTA mode with g.vslices > gridDim.x: the tensor is cut into more slices than there are CTAs; a CTA starts on slice
    // blockIdx.x and takes further ones from a ticket counter (fetched a slice ahead), so the CTAs that the memory system
    // serves faster do more of them and all finish together -- with one static slice per CTA the last 15 % of the kernel ran at
    // half the DRAM throughput (profiles/r02_pm_series.txt).  Partial slots are indexed by slice, not by CTA: results do not
    // depend on who computed what.
    __shared__ int next_slice[2];
    int par = 0;
    unsigned int* slice_ticket = reinterpret_cast<unsigned int*>(done_counter) + 16;      // bytes 64 / 128 of the counter block
    unsigned int* slice_exits = reinterpret_cast<unsigned int*>(done_counter) + 32;
    const bool dynamic = G > 32 && g.vslices > (int)gridDim.x;
    int64_t gi = GroupIdx<G>::index();
  for (;;) {
    if (dynamic && t == 0) next_slice[par] = (int)(gridDim.x + atomicAdd(slice_ticket, 1u));
    PieceIter<G> it(g, gi);
    Piece pc;
    int pending_n = -1;                    // sample whose finished vectors have not been ticketed yet
    unsigned long long pending = 0;
    bool more = it.next(pc);
    while (more) {
        const T* gsrc = dy + pc.plane * g.M;
        const T* xsrc = x + pc.plane * g.M;
        T* dst = DX ? dx + pc.plane * g.M : nullptr;
        const int n = (int)(pc.plane / tb.C);
        const float m = __ldg(tb.mu_all + ((int64_t)tb.row_offset + n) * tb.ld + (pc.plane - (int64_t)n * tb.C));
        const float a = __ldg(tb.scale + pc.plane);
        float s1 = 0.f, s2 = 0.f;
        const int len = pc.v1 - pc.v0;
        const Batches<G, VPT> bt(pc, g.reverse != 0);
        for (int i = 0; i < bt.full; ++i) {
            const int64_t o = (int64_t)(bt.begin(i) + t) * VEC;
            float gv[VPT][VEC], xv[VPT][VEC];
#pragma unroll
            for (int j = 0; j < VPT; ++j) {
                Vec<T, VEC>::load(gsrc + o + (int64_t)j * G * VEC, gv[j], pol_io);
                Vec<T, VEC>::load(xsrc + o + (int64_t)j * G * VEC, xv[j], pol_x);
                pre_apply_vec(xv[j], g.pre_op, g.pre_param);
            }
            float b1 = 0.f, b2 = 0.f;
#pragma unroll
            for (int j = 0; j < VPT; ++j) {
#pragma unroll
                for (int k = 0; k < VEC; ++k) {
                    b1 += gv[j][k];
                    b2 = fmaf(gv[j][k], xv[j][k] - m, b2);
                }
                if constexpr (DX) {
#pragma unroll
                    for (int k = 0; k < VEC; ++k) gv[j][k] *= a;
                    if (g.pre_op != kPreNone) {                  // dz = dx * act'(z), through x = act(z)
#pragma unroll
                        for (int k = 0; k < VEC; ++k) gv[j][k] *= pre_grad(xv[j][k], g.pre_op, g.pre_param);
                    }
                    Vec<T, VEC>::store(dst + o + (int64_t)j * G * VEC, gv[j], pol_io);
                }
            }
            s1 += b1;
            s2 += b2;
        }
        if (bt.rem) {
            const int hi = bt.ragged_hi();
            for (int lo = bt.ragged_lo() + t; lo < hi; lo += G) {
                float gv[VEC], xv[VEC];
                Vec<T, VEC>::load(gsrc + (int64_t)lo * VEC, gv, pol_io);
                Vec<T, VEC>::load(xsrc + (int64_t)lo * VEC, xv, pol_x);
                pre_apply_vec(xv, g.pre_op, g.pre_param);
                float b1 = 0.f, b2 = 0.f;
#pragma unroll
                for (int k = 0; k < VEC; ++k) {
                    b1 += gv[k];
                    b2 = fmaf(gv[k], xv[k] - m, b2);
                }
                if constexpr (DX) {
#pragma unroll
                    for (int k = 0; k < VEC; ++k) gv[k] *= a * pre_grad(xv[k], g.pre_op, g.pre_param);
                    Vec<T, VEC>::store(dst + (int64_t)lo * VEC, gv, pol_io);
                }
                s1 += b1;
                s2 += b2;
            }
        }
        group_sum2<G>(s1, s2, scratch);
        if (t == 0) {
            int k = 0;
            if constexpr (G > 32) k = (int)(gi - plane_share(g, pc.plane).first);
            partials[pc.plane * g.slots + k] = make_float4(s1, s2, 0.f, 0.f);
        }
        pending_n = n;
        pending += (unsigned long long)len;
        more = it.next(pc);
        // ticket the finished vectors once per (group, sample): when the sample changes or the slice ends
        if (!more || (int)(pc.plane / tb.C) != pending_n) {
            bool last = false;
            if (t == 0) last = ticket_add(&sample_tickets[pending_n], pending, sample_total);
            if (group_bcast<G>(last, scratch)) bwd_finalize_sample<G>(pending_n, tb, st, g, partials, done_counter, scratch);
            pending = 0;
        }
    }
    if (!dynamic) break;
    __syncthreads();
    gi = next_slice[par];                  // rewritten two slices later, behind this barrier
    par ^= 1;
    if (gi >= g.vslices) {
        if (t == 0 && atomicAdd(slice_exits, 1u) == gridDim.x - 1u) { *slice_ticket = 0u; *slice_exits = 0u; }      // last CTA out
        break;
    }
  }
}

}  // namespace ms
