// Streaming kernels for NCHW-contiguous feature maps: every (n,c) plane is M = H*W contiguous
// elements.  A work item is (plane, split): `chunk` 16-byte vectors of one plane, handled by a
// group of G threads (a warp for small planes, the whole CTA otherwise).  Grids are persistent:
// 148 SMs x kBlocksPerSM CTAs stride over the item list, so neighbouring CTAs touch neighbouring
// memory at the same time and there is no per-item CTA launch cost.
#pragma once
#include "common.cuh"
#include "plan.h"

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

struct ItemGeom {
    int64_t M;        // elements per plane
    int64_t nvec;     // vectors per plane
    int64_t chunk;    // vectors per item
    int64_t items;    // planes * splits
    int splits;
};

template <int G> struct GroupIdx {
    static constexpr int kPerBlock = kThreads / G;
    __device__ static __forceinline__ int lane() { return G > 32 ? threadIdx.x : (threadIdx.x & 31); }
    __device__ static __forceinline__ int64_t first() {
        return (int64_t)blockIdx.x * kPerBlock + (G > 32 ? 0 : (threadIdx.x >> 5));
    }
    __device__ static __forceinline__ int64_t stride() { return (int64_t)gridDim.x * kPerBlock; }
};

// ---------------------------------------------------------------------------------------------
// Kernel 1: instance statistics.  Replaces x.mean(dim=[2,3]) / x.var(dim=[2,3]) / sqrt(var+eps)
// of the reference (src/advanced/maxstyle.py:157-159) with ONE read of x.
// Per thread: batches of VPT vectors are reduced two-pass in registers (sum -> mean -> squared
// deviations) and folded into a running (n, mean, M2) with the Chan/Welford merge; then warp
// shuffles, then shared memory across the CTA's warps, then (splits > 1) a last-arriver merge
// of the per-item partials in fixed order, so results are run-to-run deterministic.
// ---------------------------------------------------------------------------------------------
template <typename T, int VEC, int G, int VPT, Hint LOAD>
__global__ void __launch_bounds__(kThreads, kBlocksPerSM)
stats_nchw_kernel(const T* __restrict__ x, float* __restrict__ mu, float* __restrict__ sig, TableRef tr,
                  float4* __restrict__ partials, int* __restrict__ plane_counters, ItemGeom g, float eps) {
    __shared__ Scratch scratch;
    const int t = GroupIdx<G>::lane();
    const float inv_m1 = 1.0f / (float)(g.M - 1);
    for (int64_t item = GroupIdx<G>::first(); item < g.items; item += GroupIdx<G>::stride()) {
        const int64_t plane = item / g.splits;
        const int split = (int)(item - plane * g.splits);
        const int64_t v_begin = (int64_t)split * g.chunk;
        const int64_t v_end = min(g.nvec, v_begin + g.chunk);
        const T* base = x + plane * g.M;
        Moments acc{0.f, 0.f, 0.f};
        for (int64_t v0 = v_begin + t; v0 < v_end; v0 += (int64_t)G * VPT) {
            float val[VPT][VEC];
            bool ok[VPT];
#pragma unroll
            for (int j = 0; j < VPT; ++j) {
                const int64_t idx = v0 + (int64_t)j * G;
                ok[j] = idx < v_end;
                if (ok[j]) {
                    Vec<T, VEC>::template load<LOAD>(base + idx * VEC, val[j]);
                } else {
#pragma unroll
                    for (int k = 0; k < VEC; ++k) val[j][k] = 0.f;
                }
            }
            Moments b;
            if (ok[VPT - 1]) {                         // full batch: constant count
                float s = 0.f;
#pragma unroll
                for (int j = 0; j < VPT; ++j)
#pragma unroll
                    for (int k = 0; k < VEC; ++k) s += val[j][k];
                b.n = (float)(VPT * VEC);
                b.mean = s * (1.0f / (float)(VPT * VEC));
                float q = 0.f;
#pragma unroll
                for (int j = 0; j < VPT; ++j)
#pragma unroll
                    for (int k = 0; k < VEC; ++k) { const float d = val[j][k] - b.mean; q = fmaf(d, d, q); }
                b.m2 = q;
            } else {                                   // ragged tail of the item
                float s = 0.f, cnt = 0.f;
#pragma unroll
                for (int j = 0; j < VPT; ++j) {
                    cnt += ok[j] ? (float)VEC : 0.f;
#pragma unroll
                    for (int k = 0; k < VEC; ++k) s += val[j][k];
                }
                b.n = cnt;
                b.mean = s / cnt;
                float q = 0.f;
#pragma unroll
                for (int j = 0; j < VPT; ++j)
#pragma unroll
                    for (int k = 0; k < VEC; ++k) { const float d = ok[j] ? val[j][k] - b.mean : 0.f; q = fmaf(d, d, q); }
                b.m2 = q;
            }
            acc = merge(acc, b);
        }
        const Moments tot = group_merge<G>(acc, scratch);
        if (g.splits == 1) {
            if (t == 0) {
                const int64_t o = tr.at(plane);
                mu[o] = tot.mean;
                sig[o] = sqrtf(tot.m2 * inv_m1 + eps);
            }
        } else {
            if (t == 0) partials[item] = make_float4(tot.n, tot.mean, tot.m2, 0.f);
            if (arrive_is_last<G>(&plane_counters[plane], g.splits, scratch)) {
                if (t == 0) {
                    Moments m{0.f, 0.f, 0.f};
                    for (int s = 0; s < g.splits; ++s) {
                        const float4 p = __ldcg(&partials[plane * g.splits + s]);
                        m = merge(m, Moments{p.x, p.y, p.z});
                    }
                    const int64_t o = tr.at(plane);
                    mu[o] = m.mean;
                    sig[o] = sqrtf(m.m2 * inv_m1 + eps);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Kernel 2: apply.  y = (x - mu) * scale + shift with scale = A/sig, shift = B, i.e. the
// normalise + mix + perturb + affine chain of maxstyle.py:161,172-185 folded into one FMA per
// element (the [N,C] tables come from tables_kernel).  x is read for the last time (evict-first),
// y is written with streaming stores.
// ---------------------------------------------------------------------------------------------
template <typename T, int VEC, int G, int VPT>
__global__ void __launch_bounds__(kThreads, kBlocksPerSM)
apply_nchw_kernel(const T* __restrict__ x, T* __restrict__ y, const float* __restrict__ mu, TableRef tr,
                  const float* __restrict__ scale, const float* __restrict__ shift, ItemGeom g) {
    const int t = GroupIdx<G>::lane();
    for (int64_t item = GroupIdx<G>::first(); item < g.items; item += GroupIdx<G>::stride()) {
        const int64_t plane = item / g.splits;
        const int split = (int)(item - plane * g.splits);
        const int64_t v_begin = (int64_t)split * g.chunk;
        const int64_t v_end = min(g.nvec, v_begin + g.chunk);
        const T* src = x + plane * g.M;
        T* dst = y + plane * g.M;
        const float m = __ldg(mu + tr.at(plane)), a = __ldg(scale + plane), b = __ldg(shift + plane);
        for (int64_t v0 = v_begin + t; v0 < v_end; v0 += (int64_t)G * VPT) {
            float val[VPT][VEC];
#pragma unroll
            for (int j = 0; j < VPT; ++j) {
                const int64_t idx = v0 + (int64_t)j * G;
                if (idx < v_end) Vec<T, VEC>::template load<Hint::kStream>(src + idx * VEC, val[j]);
            }
#pragma unroll
            for (int j = 0; j < VPT; ++j) {
                const int64_t idx = v0 + (int64_t)j * G;
                if (idx < v_end) {
#pragma unroll
                    for (int k = 0; k < VEC; ++k) val[j][k] = fmaf(val[j][k] - m, a, b);
                    Vec<T, VEC>::store(dst + idx * VEC, val[j]);
                }
            }
        }
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

__device__ __forceinline__ StepCoef step_coef(const StepArgs& s) {
    StepCoef c{};
    if (s.mode == 1) {
        const int t = s.step_dev ? (*(volatile int*)s.step_dev + 1) : s.t;
        const double bc1 = 1.0 - pow(s.beta1, (double)t);
        const double bc2 = 1.0 - pow(s.beta2, (double)t);
        c.one_minus_b1 = (float)(1.0 - s.beta1);
        c.b2 = (float)s.beta2;
        c.one_minus_b2 = (float)(1.0 - s.beta2);
        c.step_size = (float)(s.lr / bc1);
        c.bc2_sqrt = (float)sqrt(bc2);
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
// finishing the last item of a sample runs the epilogue for that sample: parameter gradients,
// the channel reduction for d_lmda in fixed order (no float atomics) and the fused optimiser step.
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
__device__ __forceinline__ void bwd_finalize_sample(int n, const BwdTables& tb, const StepArgs& st,
                                                    const float4* partials, int splits, int* done_counter,
                                                    Scratch& scratch) {
    const int t = GroupIdx<G>::lane();
    const int C = tb.C;
    const int64_t row = (int64_t)tb.row_offset + n;
    const bool mix = tb.flags & 1, no_noise = tb.flags & 2;
    const int64_t prow = mix ? tb.perm[row] : row;
    StepCoef coef = step_coef(st);
    float lam_acc = 0.f, unused = 0.f;
    for (int c = t; c < C; c += G) {
        const int64_t plane = (int64_t)n * C + c;
        float s1 = 0.f, s2 = 0.f;
        for (int s = 0; s < splits; ++s) {
            const float4 p = __ldcg(&partials[plane * splits + s]);
            s1 += p.x;
            s2 += p.y;
        }
        const int ld = tb.ld;
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
            dl = (l >= 0.f && l <= 1.f) ? lam_acc : 0.f;      // clamp backward: closed interval
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
                int* __restrict__ sample_counters, int* __restrict__ done_counter, ItemGeom g, BwdTables tb,
                StepArgs st) {
    __shared__ Scratch scratch;
    const int t = GroupIdx<G>::lane();
    for (int64_t item = GroupIdx<G>::first(); item < g.items; item += GroupIdx<G>::stride()) {
        const int64_t plane = item / g.splits;
        const int split = (int)(item - plane * g.splits);
        const int64_t v_begin = (int64_t)split * g.chunk;
        const int64_t v_end = min(g.nvec, v_begin + g.chunk);
        const T* gsrc = dy + plane * g.M;
        const T* xsrc = x + plane * g.M;
        T* dst = DX ? dx + plane * g.M : nullptr;
        const int n = (int)(plane / tb.C);
        const float m = __ldg(tb.mu_all + ((int64_t)tb.row_offset + n) * tb.ld + (plane - (int64_t)n * tb.C));
        const float a = __ldg(tb.scale + plane);
        float s1 = 0.f, s2 = 0.f;
        for (int64_t v0 = v_begin + t; v0 < v_end; v0 += (int64_t)G * VPT) {
            float gv[VPT][VEC], xv[VPT][VEC];
#pragma unroll
            for (int j = 0; j < VPT; ++j) {
                const int64_t idx = v0 + (int64_t)j * G;
                if (idx < v_end) {
                    Vec<T, VEC>::template load<Hint::kStream>(gsrc + idx * VEC, gv[j]);
                    Vec<T, VEC>::template load<Hint::kStream>(xsrc + idx * VEC, xv[j]);
                } else {
#pragma unroll
                    for (int k = 0; k < VEC; ++k) { gv[j][k] = 0.f; xv[j][k] = m; }
                }
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
                    const int64_t idx = v0 + (int64_t)j * G;
                    if (idx < v_end) {
#pragma unroll
                        for (int k = 0; k < VEC; ++k) gv[j][k] *= a;
                        Vec<T, VEC>::store(dst + idx * VEC, gv[j]);
                    }
                }
            }
            s1 += b1;
            s2 += b2;
        }
        group_sum2<G>(s1, s2, scratch);
        if (t == 0) partials[item] = make_float4(s1, s2, 0.f, 0.f);
        if (arrive_is_last<G>(&sample_counters[n], tb.C * g.splits, scratch))
            bwd_finalize_sample<G>(n, tb, st, partials, g.splits, done_counter, scratch);
    }
}

}  // namespace ms
