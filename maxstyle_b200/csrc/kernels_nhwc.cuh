// Streaming kernels for NHWC (channels-last) feature maps: memory is [N][H*W][C], a pixel is a row of
// C contiguous elements = CV vector accesses, a sample is M*CV contiguous vectors (plan.h: PlanNhwc).
// Same flat sweep as the NCHW kernels, but the unit that CTAs share is the SAMPLE: the tensor is one
// index space of N*M*CV vectors, CTA b owns the contiguous slice [b*per, (b+1)*per) cut into pieces at
// sample boundaries.  Inside a piece `active` = floor(256/CV)*CV threads stream rows of `active`
// consecutive vectors, so every access is fully coalesced AND a thread meets the same VEC channels at
// every step: the per-channel statistics (mean/M2, or the two backward sums) live in registers for the
// whole piece.  At the end of a piece the threads that hold the same channels are merged (warp shuffles
// when CV divides 32, then shared memory, fixed order), the CTA publishes one partial per channel, and the
// last CTA to finish a sample merges the partials in slot order -- no float atomics, deterministic.
#pragma once
#include "common.cuh"
#include "kernels_nchw.cuh"

namespace ms {

struct RowGeom {
    int C;            // channels
    int cv;           // vectors per pixel
    int active;       // streaming threads per CTA (multiple of cv)
    int shuffle;      // cv divides 32 (and is < 32): merge inside the warp first
};

template <int VEC> struct ChanScratch {
    float n[kThreads];                      // entry e = slot*cv + channel-vector
    float a[VEC][kThreads + 1];
    float b[VEC][kThreads + 1];
    Scratch s;
};

// ---- per-channel reductions across the threads of a CTA ---------------------------------------------
// Stage 1 (optional) folds the lanes of a warp that hold the same channel vector; stage 2 stores one entry
// per (slot, channel vector); the caller then walks the `nslots` entries of a channel in order.
template <int VEC>
__device__ __forceinline__ int chan_publish_moments(float cnt, float (&mean)[VEC], float (&m2)[VEC], int cvv, const RowGeom& rg,
                                                    ChanScratch<VEC>& sh) {
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (rg.shuffle) {
        for (int o = 16; o >= rg.cv; o >>= 1) {
            const float cnt_o = __shfl_xor_sync(0xffffffffu, cnt, o);
            const float nn = cnt + cnt_o;
            const float w = nn > 0.f ? cnt_o / nn : 0.f;
            const float cw = cnt * w;
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                const float mean_o = __shfl_xor_sync(0xffffffffu, mean[k], o);
                const float m2_o = __shfl_xor_sync(0xffffffffu, m2[k], o);
                const float d = mean_o - mean[k];
                mean[k] = fmaf(d, w, mean[k]);
                m2[k] = m2[k] + m2_o + d * d * cw;
            }
            cnt = nn;
        }
    }
    __syncthreads();                                            // the scratch may still be read from the previous piece
    const bool writer = rg.shuffle ? lane < rg.cv : t < rg.active;
    if (writer) {
        const int slot = rg.shuffle ? warp : t / rg.cv;
        const int e = slot * rg.cv + cvv;
        sh.n[e] = cnt;
#pragma unroll
        for (int k = 0; k < VEC; ++k) { sh.a[k][e] = mean[k]; sh.b[k][e] = m2[k]; }
    }
    __syncthreads();
    return rg.shuffle ? kWarps : rg.active / rg.cv;
}

template <int VEC>
__device__ __forceinline__ Moments chan_moments(int c, int nslots, const RowGeom& rg, const ChanScratch<VEC>& sh) {
    const int cvv = c / VEC, k = c - cvv * VEC;
    Moments m{0.f, 0.f, 0.f};
    for (int s = 0; s < nslots; ++s) {
        const int e = s * rg.cv + cvv;
        m = merge(m, Moments{sh.n[e], sh.a[k][e], sh.b[k][e]});
    }
    return m;
}

template <int VEC>
__device__ __forceinline__ int chan_publish_sums(float (&s1)[VEC], float (&s2)[VEC], int cvv, const RowGeom& rg, ChanScratch<VEC>& sh) {
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (rg.shuffle) {
        for (int o = 16; o >= rg.cv; o >>= 1) {
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                s1[k] += __shfl_xor_sync(0xffffffffu, s1[k], o);
                s2[k] += __shfl_xor_sync(0xffffffffu, s2[k], o);
            }
        }
    }
    __syncthreads();
    const bool writer = rg.shuffle ? lane < rg.cv : t < rg.active;
    if (writer) {
        const int slot = rg.shuffle ? warp : t / rg.cv;
        const int e = slot * rg.cv + cvv;
#pragma unroll
        for (int k = 0; k < VEC; ++k) { sh.a[k][e] = s1[k]; sh.b[k][e] = s2[k]; }
    }
    __syncthreads();
    return rg.shuffle ? kWarps : rg.active / rg.cv;
}

// ---------------------------------------------------------------------------------------------
// Kernel 1 (NHWC): instance statistics, maxstyle.py:157-159, one read of x.
// Per thread and channel a running (mean, M2) with a count common to the thread's VEC channels: batches of
// VPT pixels are reduced two-pass in registers (sum -> mean -> squared deviations) and folded in with the
// Chan/Welford merge (one fast reciprocal per batch, shared by the VEC channels).
// ---------------------------------------------------------------------------------------------
template <typename T, int VEC, int VPT>
__global__ void __launch_bounds__(kThreads, kBlocksPerSMNhwc)
stats_nhwc_kernel(const T* __restrict__ x, float* __restrict__ mu, float* __restrict__ sig, TableRef tr,
                  float4* __restrict__ partials, unsigned long long* __restrict__ sample_tickets, Sweep g, RowGeom rg, float eps) {
    __shared__ ChanScratch<VEC> sh;
    const int t = threadIdx.x;
    const uint64_t pol = make_policy(g.in_policy);
    const float inv_m1 = 1.0f / (float)(g.M - 1);
    const int C = rg.C, A = rg.active;
    PieceIter<kThreads> it(g, blockIdx.x);
    Piece pc;
    while (it.next(pc)) {
        const int64_t n = pc.plane;
        const T* base = x + n * g.nvec * VEC;
        const int cvv = (int)((pc.v0 + (int64_t)t) % rg.cv);
        float mean[VEC], m2[VEC], K[VEC], cnt = 0.f;
#pragma unroll
        for (int k = 0; k < VEC; ++k) { mean[k] = 0.f; m2[k] = 0.f; K[k] = 0.f; }
        if (t < A && pc.v0 + t < pc.v1) {
            int v = pc.v0 + t;
            // Fixed shift per (sample, channel): the sample's FIRST pixel.  Every value has K subtracted before
            // anything is summed, so data whose spread is tiny against its mean loses no digits, and because all
            // threads and all CTAs that share the sample use the same K, every merge (registers, shuffles, shared
            // memory, per-CTA partials) happens between small shifted means; K is added back once, at the end.
            Vec<T, VEC>::load(base + (int64_t)cvv * VEC, K, pol);
            for (; v + (VPT - 1) * A < pc.v1; v += VPT * A) {
                float val[VPT][VEC];
#pragma unroll
                for (int j = 0; j < VPT; ++j) Vec<T, VEC>::load(base + (int64_t)(v + j * A) * VEC, val[j], pol);
                const float nn = cnt + (float)VPT;
                const float w = __fdividef((float)VPT, nn);
                const float cw = cnt * w;
#pragma unroll
                for (int k = 0; k < VEC; ++k) {
                    float s = 0.f;
#pragma unroll
                    for (int j = 0; j < VPT; ++j) { val[j][k] -= K[k]; s += val[j][k]; }
                    const float bm = s * (1.0f / (float)VPT);      // batch mean (shifted)
                    float q = 0.f;
#pragma unroll
                    for (int j = 0; j < VPT; ++j) { const float d = val[j][k] - bm; q = fmaf(d, d, q); }
                    const float dm = bm - mean[k];
                    mean[k] = fmaf(dm, w, mean[k]);
                    m2[k] += fmaf(dm * dm, cw, q);
                }
                cnt = nn;
            }
            for (; v < pc.v1; v += A) {                       // ragged end: single pixels
                float val[VEC];
                Vec<T, VEC>::load(base + (int64_t)v * VEC, val, pol);
                const float nn = cnt + 1.f;
                const float w = __fdividef(1.f, nn);
                const float cw = cnt * w;
#pragma unroll
                for (int k = 0; k < VEC; ++k) {
                    const float dd = (val[k] - K[k]) - mean[k];
                    mean[k] = fmaf(dd, w, mean[k]);
                    m2[k] = fmaf(dd * dd, cw, m2[k]);
                }
                cnt = nn;
            }
        }
        const int nslots = chan_publish_moments<VEC>(cnt, mean, m2, cvv, rg, sh);
        const int len = pc.v1 - pc.v0;
        if (len == g.nvec) {                                   // the piece is the whole sample
            for (int c = t; c < C; c += kThreads) {
                const Moments m = chan_moments<VEC>(c, nslots, rg, sh);
                const int64_t o = ((int64_t)tr.row_offset + n) * tr.ld + c;
                mu[o] = to_f32<T>(__ldg(base + c)) + m.mean;
                sig[o] = sqrtf(m.m2 * inv_m1 + eps);
            }
        } else {
            const PlaneShare shr = plane_share(g, n);
            const int slot = (int)(blockIdx.x - shr.first);
            for (int c = t; c < C; c += kThreads) {
                const Moments m = chan_moments<VEC>(c, nslots, rg, sh);
                partials[(n * C + c) * g.slots + slot] = make_float4(m.n, m.mean, m.m2, 0.f);
            }
            __threadfence();
            __syncthreads();
            bool last = false;
            if (t == 0) last = ticket_add(&sample_tickets[n], (unsigned long long)len, (unsigned long long)g.nvec);
            if (group_bcast<kThreads>(last, sh.s)) {
                for (int c = t; c < C; c += kThreads) {
                    const float4* slot0 = partials + (n * C + c) * g.slots;
                    Moments m{0.f, 0.f, 0.f};
                    for (int k = 0; k < shr.count; ++k) {
                        const float4 v = __ldcg(&slot0[k]);
                        m = merge(m, Moments{v.x, v.y, v.z});
                    }
                    const int64_t o = ((int64_t)tr.row_offset + n) * tr.ld + c;
                    mu[o] = to_f32<T>(__ldg(base + c)) + m.mean;
                    sig[o] = sqrtf(m.m2 * inv_m1 + eps);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Kernel 2 (NHWC): y = (x - mu[n,c]) * scale[n,c] + shift[n,c]; the thread's VEC channels' coefficients
// are loaded once per piece.
// ---------------------------------------------------------------------------------------------
template <typename T, int VEC, int VPT>
__global__ void __launch_bounds__(kThreads, kBlocksPerSMNhwc)
apply_nhwc_kernel(const T* __restrict__ x, T* __restrict__ y, const float* __restrict__ mu, TableRef tr,
                  const float* __restrict__ scale, const float* __restrict__ shift, Sweep g, RowGeom rg) {
    const int t = threadIdx.x;
    const int A = rg.active;
    if (t >= A) return;
    const uint64_t pol_in = make_policy(g.in_policy), pol_out = make_policy(g.io_policy);
    PieceIter<kThreads> it(g, blockIdx.x);
    Piece pc;
    while (it.next(pc)) {
        const int64_t n = pc.plane;
        const T* src = x + n * g.nvec * VEC;
        T* dst = y + n * g.nvec * VEC;
        const int c0 = (int)((pc.v0 + (int64_t)t) % rg.cv) * VEC;
        float m[VEC], a[VEC], b[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            m[k] = __ldg(mu + ((int64_t)tr.row_offset + n) * tr.ld + c0 + k);
            a[k] = __ldg(scale + n * rg.C + c0 + k);
            b[k] = __ldg(shift + n * rg.C + c0 + k);
        }
        int v = pc.v0 + t;
        for (; v + (VPT - 1) * A < pc.v1; v += VPT * A) {
            float val[VPT][VEC];
#pragma unroll
            for (int j = 0; j < VPT; ++j) Vec<T, VEC>::load(src + (int64_t)(v + j * A) * VEC, val[j], pol_in);
#pragma unroll
            for (int j = 0; j < VPT; ++j) {
#pragma unroll
                for (int k = 0; k < VEC; ++k) val[j][k] = fmaf(val[j][k] - m[k], a[k], b[k]);
                Vec<T, VEC>::store(dst + (int64_t)(v + j * A) * VEC, val[j], pol_out);
            }
        }
        for (; v < pc.v1; v += A) {
            float val[VEC];
            Vec<T, VEC>::load(src + (int64_t)v * VEC, val, pol_in);
#pragma unroll
            for (int k = 0; k < VEC; ++k) val[k] = fmaf(val[k] - m[k], a[k], b[k]);
            Vec<T, VEC>::store(dst + (int64_t)v * VEC, val, pol_out);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Kernel 3 (NHWC): backward.  dx = dy*scale;  per (n,c): S1 = sum dy, S2 = sum dy*(x-mu) kept in registers
// per channel; the CTA that completes a sample runs the same epilogue as the NCHW kernel
// (bwd_finalize_sample: parameter gradients, d_lmda in fixed order, fused optimiser step).
// ---------------------------------------------------------------------------------------------
template <typename T, int VEC, int VPT, bool DX>
__global__ void __launch_bounds__(kThreads, kBlocksPerSMNhwc)
bwd_nhwc_kernel(const T* __restrict__ dy, const T* __restrict__ x, T* __restrict__ dx, float4* __restrict__ partials,
                unsigned long long* __restrict__ sample_tickets, int* __restrict__ done_counter, Sweep g, RowGeom rg,
                BwdTables tb, StepArgs st) {
    __shared__ ChanScratch<VEC> sh;
    const int t = threadIdx.x;
    const uint64_t pol_x = make_policy(g.in_policy), pol_io = make_policy(g.io_policy);
    const int C = rg.C, A = rg.active;
    PieceIter<kThreads> it(g, blockIdx.x);
    Piece pc;
    while (it.next(pc)) {
        const int64_t n = pc.plane;
        const T* gsrc = dy + n * g.nvec * VEC;
        const T* xsrc = x + n * g.nvec * VEC;
        T* dst = DX ? dx + n * g.nvec * VEC : nullptr;
        const int cvv = (int)((pc.v0 + (int64_t)t) % rg.cv);
        float m[VEC], a[VEC], s1[VEC], s2[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) { m[k] = 0.f; a[k] = 0.f; s1[k] = 0.f; s2[k] = 0.f; }
        if (t < A) {
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                m[k] = __ldg(tb.mu_all + ((int64_t)tb.row_offset + n) * tb.ld + cvv * VEC + k);
                if constexpr (DX) a[k] = __ldg(tb.scale + n * C + cvv * VEC + k);
            }
            int v = pc.v0 + t;
            for (; v + (VPT - 1) * A < pc.v1; v += VPT * A) {
                float gv[VPT][VEC], xv[VPT][VEC];
#pragma unroll
                for (int j = 0; j < VPT; ++j) {
                    Vec<T, VEC>::load(gsrc + (int64_t)(v + j * A) * VEC, gv[j], pol_io);
                    Vec<T, VEC>::load(xsrc + (int64_t)(v + j * A) * VEC, xv[j], pol_x);
                }
#pragma unroll
                for (int j = 0; j < VPT; ++j) {
#pragma unroll
                    for (int k = 0; k < VEC; ++k) {
                        s1[k] += gv[j][k];
                        s2[k] = fmaf(gv[j][k], xv[j][k] - m[k], s2[k]);
                    }
                    if constexpr (DX) {
#pragma unroll
                        for (int k = 0; k < VEC; ++k) gv[j][k] *= a[k];
                        Vec<T, VEC>::store(dst + (int64_t)(v + j * A) * VEC, gv[j], pol_io);
                    }
                }
            }
            for (; v < pc.v1; v += A) {
                float gv[VEC], xv[VEC];
                Vec<T, VEC>::load(gsrc + (int64_t)v * VEC, gv, pol_io);
                Vec<T, VEC>::load(xsrc + (int64_t)v * VEC, xv, pol_x);
#pragma unroll
                for (int k = 0; k < VEC; ++k) {
                    s1[k] += gv[k];
                    s2[k] = fmaf(gv[k], xv[k] - m[k], s2[k]);
                }
                if constexpr (DX) {
#pragma unroll
                    for (int k = 0; k < VEC; ++k) gv[k] *= a[k];
                    Vec<T, VEC>::store(dst + (int64_t)v * VEC, gv, pol_io);
                }
            }
        }
        const int nslots = chan_publish_sums<VEC>(s1, s2, cvv, rg, sh);
        const PlaneShare shr = plane_share(g, n);
        const int slot = (int)(blockIdx.x - shr.first);
        for (int c = t; c < C; c += kThreads) {
            const int cv2 = c / VEC, k = c - cv2 * VEC;
            float u = 0.f, w = 0.f;
            for (int s = 0; s < nslots; ++s) { u += sh.a[k][s * rg.cv + cv2]; w += sh.b[k][s * rg.cv + cv2]; }
            partials[(n * C + c) * g.slots + slot] = make_float4(u, w, 0.f, 0.f);
        }
        __threadfence();
        __syncthreads();
        bool last = false;
        if (t == 0) last = ticket_add(&sample_tickets[n], (unsigned long long)(pc.v1 - pc.v0), (unsigned long long)g.nvec);
        if (group_bcast<kThreads>(last, sh.s)) bwd_finalize_sample<kThreads>((int)n, tb, st, g, partials, done_counter, sh.s, shr.count);
    }
}

}  // namespace ms
