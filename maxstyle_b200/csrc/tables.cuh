// Small-table kernels: everything the reference does on the [N,C] style tables
// (src/advanced/maxstyle.py:165-185) and the stand-alone optimiser step.
#pragma once
#include "common.cuh"
#include "kernels_nchw.cuh"

namespace ms {

constexpr int kTableThreads = 128;

__device__ __forceinline__ float block_sum_128(float v, float* red) {
    v = warp_sum(v);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    return red[0] + red[1] + red[2] + red[3];
}

// One CTA per channel c.
//  (1) first forward only: gamma_std[c] = std_n(sig[:,c]), beta_std[c] = std_n(mu[:,c]), unbiased over
//      the GLOBAL batch, two-pass (maxstyle.py:165-168);
//  (2) for the local rows: l = clamp(lmda,0,1); partner row = perm[row]; sig_mix / mu_mix lerp
//      (maxstyle.py:173-176); A = sig_mix + gamma_noise*gamma_std, B = mu_mix + beta_noise*beta_std
//      (:184-185); scale = A/sig, shift = B.
__global__ void __launch_bounds__(kTableThreads)
tables_kernel(const float* __restrict__ mu_all, const float* __restrict__ sig_all, int ld, int n_global, int row_offset,
              int n_local, int C, const int64_t* __restrict__ perm, const float* __restrict__ lmda,
              const float* __restrict__ gamma_noise, const float* __restrict__ beta_noise,
              float* __restrict__ gamma_std, float* __restrict__ beta_std, int flags,
              float* __restrict__ scale, float* __restrict__ shift) {
    __shared__ float red[4];
    const int c = blockIdx.x;
    const bool mix = flags & 1, no_noise = flags & 2, compute_std = flags & 4;
    float gs = 0.f, bs = 0.f;
    if (compute_std && gamma_std != nullptr) {     // the reference fills the cache whatever no_noise says (:165-168)
        float s_sig = 0.f, s_mu = 0.f;
        for (int n = threadIdx.x; n < n_global; n += kTableThreads) {
            s_sig += sig_all[(int64_t)n * ld + c];
            s_mu += mu_all[(int64_t)n * ld + c];
        }
        const float mean_sig = block_sum_128(s_sig, red) / (float)n_global;
        const float mean_mu = block_sum_128(s_mu, red) / (float)n_global;
        float q_sig = 0.f, q_mu = 0.f;
        for (int n = threadIdx.x; n < n_global; n += kTableThreads) {
            const float ds = sig_all[(int64_t)n * ld + c] - mean_sig;
            const float dm = mu_all[(int64_t)n * ld + c] - mean_mu;
            q_sig = fmaf(ds, ds, q_sig);
            q_mu = fmaf(dm, dm, q_mu);
        }
        gs = sqrtf(block_sum_128(q_sig, red) / (float)(n_global - 1));
        bs = sqrtf(block_sum_128(q_mu, red) / (float)(n_global - 1));
        if (threadIdx.x == 0) { gamma_std[c] = gs; beta_std[c] = bs; }
    } else if (!no_noise) {
        gs = gamma_std[c];
        bs = beta_std[c];
    }
    for (int n = threadIdx.x; n < n_local; n += kTableThreads) {
        const int64_t row = (int64_t)row_offset + n;
        const float sg = sig_all[row * ld + c], m = mu_all[row * ld + c];
        const int64_t pr = mix ? perm[row] : row;
        float sc, sh;
        style_coeffs(sg, m, sig_all[pr * ld + c], mu_all[pr * ld + c], mix, no_noise, mix ? lmda[n] : 0.f,
                     no_noise ? 0.f : gamma_noise[(int64_t)n * C + c], no_noise ? 0.f : beta_noise[(int64_t)n * C + c], gs, bs,
                     sc, sh, !(flags & 8));
        scale[(int64_t)n * C + c] = sc;
        shift[(int64_t)n * C + c] = sh;
    }
}

// Stand-alone optimiser step over the three parameter tensors (gradients supplied by the caller).
__global__ void __launch_bounds__(256)
step_kernel(const float* __restrict__ d_gamma, const float* __restrict__ d_beta, const float* __restrict__ d_lmda,
            int N, int C, StepArgs st) {
    const StepCoef coef = step_coef(st);
    const int64_t nc = (int64_t)N * C;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nc && st.update_noise) {
        if (d_gamma) step_update(st.mode, st.maximize, coef, d_gamma[i], st.gamma_noise + i, st.gamma_m + i, st.gamma_v + i);
        if (d_beta) step_update(st.mode, st.maximize, coef, d_beta[i], st.beta_noise + i, st.beta_m + i, st.beta_v + i);
    }
    if (i < N && st.update_mix && d_lmda)
        step_update(st.mode, st.maximize, coef, d_lmda[i], st.lmda + i, st.lmda_m + i, st.lmda_v + i);
}

// Bumps the device-side step counter after step_kernel has finished (same stream).
__global__ void step_count_kernel(int* step_dev) { *step_dev += 1; }

}  // namespace ms
