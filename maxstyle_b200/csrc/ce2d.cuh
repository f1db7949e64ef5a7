// Pixel-wise cross entropy on NCHW logits: replaces the reference's cross_entropy_2D
// (src/models/custom_loss.py:1043-1105, label-map branch :1069-1078) -- log_softmax over C, two transposes + a contiguous
// copy to [N*H*W, C], nll_loss, mask multiply, sum, divide: ~8 kernels and three materialised [N*H*W, C] tensors forward,
// as many backward -- by ONE kernel each way that reads the logits where they are.
//   forward : loss = -(1/D) * sum_p mask_p * w[t_p] * log_softmax(l_p)[t_p],   D = N*H*W if size_average else 1
//   backward: dl[n,c,q] = g/D * mask_p * w[t_p] * (softmax(l_p)[c] - [c == t_p])
// A thread owns a pixel: its C logits sit H*W elements apart, so a warp's loads of one class are coalesced.  Pixels whose
// label is -100 (F.nll_loss's default ignore_index, which the reference inherits) contribute nothing.  The forward's sum is
// reduced per CTA and the per-CTA partials are added in index order by the last CTA to finish: run-to-run deterministic.
#pragma once
#include "common.cuh"

namespace ms {

constexpr int kCeThreads = 256;
constexpr long long kCeIgnore = -100;

// A label outside [0, C) that is not the ignore index: the reference's F.nll_loss stops with a device-side assert (the CUDA
// error reaches the host at its next synchronisation); silently skipping the pixel would train on a smaller loss.
__device__ __forceinline__ void ce_bad_label() { __trap(); }

template <typename T>
__device__ __forceinline__ float ce_logp(const T* __restrict__ base, int64_t hw, int C, int t, float& mx, float& lse) {
    mx = -INFINITY;
    for (int c = 0; c < C; ++c) mx = fmaxf(mx, to_f32<T>(base[(int64_t)c * hw]));
    float s = 0.f, lt = 0.f;
    for (int c = 0; c < C; ++c) {
        const float l = to_f32<T>(base[(int64_t)c * hw]);
        s += expf(l - mx);
        if (c == t) lt = l;
    }
    lse = logf(s);
    return lt - mx - lse;
}

template <typename T>
__global__ void __launch_bounds__(kCeThreads)
ce2d_fwd_kernel(const T* __restrict__ logits, const long long* __restrict__ target, const float* __restrict__ weight,
                const float* __restrict__ mask, float* __restrict__ loss, float* __restrict__ partials,
                unsigned int* __restrict__ counter, int64_t P, int64_t hw, int C, float inv_denom) {
    __shared__ float red[kCeThreads / 32];
    __shared__ int is_last;
    float acc = 0.f;
    for (int64_t p = (int64_t)blockIdx.x * kCeThreads + threadIdx.x; p < P; p += (int64_t)gridDim.x * kCeThreads) {
        const long long t = target[p];
        if (t == kCeIgnore) continue;
        if (t < 0 || t >= C) ce_bad_label();                 // F.nll_loss raises a device assert here; so do we
        const int64_t n = p / hw, q = p - n * hw;
        float mx, lse;
        const float lp = ce_logp<T>(logits + n * C * hw + q, hw, C, (int)t, mx, lse);
        float w = weight ? weight[t] : 1.f;
        if (mask) w *= mask[p];
        acc -= w * lp;
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < kCeThreads / 32; ++i) s += red[i];
        partials[blockIdx.x] = s;
        __threadfence();
        is_last = atomicAdd(counter, 1u) == gridDim.x - 1u;
    }
    __syncthreads();
    if (is_last && threadIdx.x == 0) {
        __threadfence();
        float s = 0.f;
        for (unsigned int i = 0; i < gridDim.x; ++i) s += __ldcg(partials + i);      // fixed order
        loss[0] = s * inv_denom;
        *counter = 0u;
    }
}

template <typename T>
__global__ void __launch_bounds__(kCeThreads)
ce2d_bwd_kernel(const T* __restrict__ logits, const long long* __restrict__ target, const float* __restrict__ weight,
                const float* __restrict__ mask, const float* __restrict__ dloss, T* __restrict__ dlogits, int64_t P, int64_t hw,
                int C, float inv_denom) {
    const float g = dloss[0] * inv_denom;
    for (int64_t p = (int64_t)blockIdx.x * kCeThreads + threadIdx.x; p < P; p += (int64_t)gridDim.x * kCeThreads) {
        const long long t = target[p];
        const int64_t n = p / hw, q = p - n * hw;
        const T* base = logits + n * C * hw + q;
        T* out = dlogits + n * C * hw + q;
        if (t == kCeIgnore) {
            for (int c = 0; c < C; ++c) out[(int64_t)c * hw] = from_f32<T>(0.f);
            continue;
        }
        if (t < 0 || t >= C) ce_bad_label();
        float mx, lse;
        ce_logp<T>(base, hw, C, (int)t, mx, lse);
        float w = (weight ? weight[t] : 1.f) * g;
        if (mask) w *= mask[p];
        for (int c = 0; c < C; ++c) {
            const float sm = expf(to_f32<T>(base[(int64_t)c * hw]) - mx - lse);
            out[(int64_t)c * hw] = from_f32<T>(w * (sm - (c == (int)t ? 1.f : 0.f)));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Loss AND its gradient in one sweep (the inner loop always back-propagates this loss: model:555-561), both branches of
// cross_entropy_2D:
//   label map   (custom_loss.py:1069-1078): loss = -(1/D) sum_p m_p w[t_p] logp_p[t_p];      dl_pc = (1/D) m_p w[t_p] (p_pc - [c == t_p])
//   soft target (custom_loss.py:1079-1102): loss = -(1/D) sum_p m_p sum_c w_c q_pc logp_pc,  q = softmax(target) or target (is_gt)
//                                           dl_pc = -(1/D) m_p (w_c q_pc - p_pc sum_k w_k q_pk)
//                                           dt_pj = -(1/D) m_p q_pj (w_j logp_pj - sum_c w_c q_pc logp_pc)   [!is_gt]
//                                           dt_pj = -(1/D) m_p w_j logp_pj                                  [is_gt]
// `dlogits` (and `dtarget`, optional) receive the gradient for an upstream gradient of 1; the autograd glue scales them by the
// actual upstream gradient with ce2d_scale_kernel (one more pass over 1/4 of the traffic, no re-evaluation of the softmax).
// kCeMaxC classes are kept in registers; more take the two-kernel path.
// ---------------------------------------------------------------------------------------------
constexpr int kCeMaxC = 8;

template <typename T, int CMAX>
__global__ void __launch_bounds__(kCeThreads)
ce2d_fused_kernel(const T* __restrict__ logits, const long long* __restrict__ labels, const T* __restrict__ soft, int is_gt,
                  const float* __restrict__ weight, const float* __restrict__ mask, float* __restrict__ loss,
                  T* __restrict__ dlogits, T* __restrict__ dsoft, float* __restrict__ partials, unsigned int* __restrict__ counter,
                  int64_t P, int64_t hw, int C, float inv_denom) {
    __shared__ float red[kCeThreads / 32];
    __shared__ int is_last;
    float acc = 0.f;
    for (int64_t p = (int64_t)blockIdx.x * kCeThreads + threadIdx.x; p < P; p += (int64_t)gridDim.x * kCeThreads) {
        const int64_t n = p / hw, q = p - n * hw;
        const T* base = logits + n * C * hw + q;
        float l[CMAX], mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < CMAX; ++c) {
            l[c] = c < C ? to_f32<T>(base[(int64_t)c * hw]) : -INFINITY;
            mx = fmaxf(mx, l[c]);
        }
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < CMAX; ++c) s += c < C ? expf(l[c] - mx) : 0.f;
        const float lse = logf(s);
        const float mk = (mask ? mask[p] : 1.f) * inv_denom;
        T* out = dlogits ? dlogits + n * C * hw + q : nullptr;
        if (labels != nullptr) {
            const long long t = labels[p];
            if (t == kCeIgnore) {
                if (out)
#pragma unroll
                    for (int c = 0; c < CMAX; ++c) if (c < C) out[(int64_t)c * hw] = from_f32<T>(0.f);
                continue;
            }
            if (t < 0 || t >= C) ce_bad_label();
            const float w = (weight ? weight[t] : 1.f) * mk;
            float lt = 0.f;
#pragma unroll
            for (int c = 0; c < CMAX; ++c) if (c == (int)t) lt = l[c];
            acc -= w * (lt - mx - lse);
            if (out)
#pragma unroll
                for (int c = 0; c < CMAX; ++c)
                    if (c < C) out[(int64_t)c * hw] = from_f32<T>(w * (expf(l[c] - mx - lse) - (c == (int)t ? 1.f : 0.f)));
        } else {
            const T* tb = soft + n * C * hw + q;
            float qv[CMAX];
            if (is_gt) {
#pragma unroll
                for (int c = 0; c < CMAX; ++c) qv[c] = c < C ? to_f32<T>(tb[(int64_t)c * hw]) : 0.f;
            } else {
                float tmx = -INFINITY, ts = 0.f;
#pragma unroll
                for (int c = 0; c < CMAX; ++c) { qv[c] = c < C ? to_f32<T>(tb[(int64_t)c * hw]) : -INFINITY; tmx = fmaxf(tmx, qv[c]); }
#pragma unroll
                for (int c = 0; c < CMAX; ++c) { qv[c] = c < C ? expf(qv[c] - tmx) : 0.f; ts += qv[c]; }
                const float inv = 1.f / ts;
#pragma unroll
                for (int c = 0; c < CMAX; ++c) qv[c] *= inv;
            }
            float wq = 0.f, wql = 0.f;                                  // sum_c w_c q_c  and  sum_c w_c q_c logp_c
#pragma unroll
            for (int c = 0; c < CMAX; ++c) {
                if (c < C) {
                    const float w = weight ? weight[c] : 1.f;
                    wq += w * qv[c];
                    wql += w * qv[c] * (l[c] - mx - lse);
                }
            }
            acc -= mk * wql;
            if (out)
#pragma unroll
                for (int c = 0; c < CMAX; ++c)
                    if (c < C) out[(int64_t)c * hw] = from_f32<T>(-mk * ((weight ? weight[c] : 1.f) * qv[c] - expf(l[c] - mx - lse) * wq));
            if (dsoft) {
                T* dt = dsoft + n * C * hw + q;
#pragma unroll
                for (int c = 0; c < CMAX; ++c) {
                    if (c < C) {
                        const float wl = (weight ? weight[c] : 1.f) * (l[c] - mx - lse);
                        dt[(int64_t)c * hw] = from_f32<T>(is_gt ? -mk * wl : -mk * qv[c] * (wl - wql));
                    }
                }
            }
        }
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < kCeThreads / 32; ++i) s += red[i];
        partials[blockIdx.x] = s;
        __threadfence();
        is_last = atomicAdd(counter, 1u) == gridDim.x - 1u;
    }
    __syncthreads();
    if (is_last && threadIdx.x == 0) {
        __threadfence();
        float s = 0.f;
        for (unsigned int i = 0; i < gridDim.x; ++i) s += __ldcg(partials + i);      // fixed order
        loss[0] = s;
        *counter = 0u;
    }
}

// g[i] *= *scale (the upstream gradient of the scalar loss, read on the device).
template <typename T>
__global__ void __launch_bounds__(256)
ce2d_scale_kernel(T* __restrict__ g, const float* __restrict__ scale, int64_t count) {
    const float s = scale[0];
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < count; i += (int64_t)gridDim.x * 256)
        g[i] = from_f32<T>(to_f32<T>(g[i]) * s);
}

}  // namespace ms
