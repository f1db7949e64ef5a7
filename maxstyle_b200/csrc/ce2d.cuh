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
        if (t == kCeIgnore || t < 0 || t >= C) continue;
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
        if (t == kCeIgnore || t < 0 || t >= C) {
            for (int c = 0; c < C; ++c) out[(int64_t)c * hw] = from_f32<T>(0.f);
            continue;
        }
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

}  // namespace ms
