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
__device__ __forceinline__ void tables_body(const float* mu_all, const float* sig_all, int ld, int n_global, int row_offset,
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

__global__ void __launch_bounds__(kTableThreads)
tables_kernel(const float* __restrict__ mu_all, const float* __restrict__ sig_all, int ld, int n_global, int row_offset,
              int n_local, int C, const int64_t* __restrict__ perm, const float* __restrict__ lmda,
              const float* __restrict__ gamma_noise, const float* __restrict__ beta_noise,
              float* __restrict__ gamma_std, float* __restrict__ beta_std, int flags,
              float* __restrict__ scale, float* __restrict__ shift) {
    tables_body(mu_all, sig_all, ld, n_global, row_offset, n_local, C, perm, lmda, gamma_noise, beta_noise, gamma_std, beta_std, flags,
                scale, shift);
}

// ---------------------------------------------------------------------------------------------
// Exchange + tables in ONE kernel over NVLink peer memory (multi-GPU forward): replaces the NCCL all-gather of the
// (mu | sig) rows AND the table kernel that follows it.  Every rank owns a small buffer in symmetric memory
// (torch.distributed._symmetric_memory: the same allocation mapped into every peer's address space):
//     inbox [2 parities][N_global][2][C] x {value bits, epoch}   8-byte words other ranks PUSHED here
// CTA c (one per channel): pushes this rank's rows of channel c into every peer's inbox as 8-byte {value, epoch} words
// with single stores over NVLink (posted writes; an 8-byte store is indivisible, so the epoch tag IS the arrival flag --
// no fence, no separate flag, no round trip: the low-latency scheme NCCL's LL protocol uses); then every thread spins on
// the words it needs in its OWN inbox (local memory) until they carry this epoch, writes them into the local
// [N_global, ld] table and the CTA goes on with the table arithmetic of tables_body on the now complete channel.  No
// grid-wide or device-wide barrier: a channel only waits for the same channel on the other ranks, and every CTA pushes
// BEFORE it waits, so ranks cannot wait on each other in a cycle.  The epoch lives in device memory (the last CTA out
// increments it), parity = epoch & 1, so the kernel's arguments never change and it replays from a CUDA graph.  Two
// parities are enough: a rank reaches epoch e+2 only after its peers pushed e+1, which they do after finishing their
// kernel of epoch e, i.e. after they consumed the words of epoch e.
// ---------------------------------------------------------------------------------------------
struct PeerTables {
    const unsigned long long* peers;   // [R] device array: base address of every rank's symmetric buffer (own entry included)
    int rank, world;
    unsigned int* epoch;               // local: number of exchanges completed so far
    unsigned int* done;                // local: CTAs finished in this launch (zero between launches)
    int* error;                        // local: set to 1 if a wait timed out
};

__device__ __forceinline__ void st_ll(void* p, float v, unsigned int tag) {
    asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(v)), "r"(tag) : "memory");
}
__device__ __forceinline__ bool ld_ll(const void* p, unsigned int tag, float& v) {
    unsigned int bits, got;
    asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(bits), "=r"(got) : "l"(p) : "memory");
    v = __uint_as_float(bits);
    return got == tag;
}

__global__ void __launch_bounds__(kTableThreads)
tables_p2p_kernel(PeerTables pt, float* __restrict__ mu_all, float* __restrict__ sig_all, int ld, int n_global, int row_offset,
                  int n_local, int C, const int64_t* __restrict__ perm, const float* __restrict__ lmda,
                  const float* __restrict__ gamma_noise, const float* __restrict__ beta_noise,
                  float* __restrict__ gamma_std, float* __restrict__ beta_std, int flags,
                  float* __restrict__ scale, float* __restrict__ shift) {
    const int c = blockIdx.x, t = threadIdx.x;
    const unsigned int epoch = *(volatile unsigned int*)pt.epoch + 1u;
    const int p = (int)(epoch & 1u);
    const size_t words_per_parity = (size_t)n_global * 2 * C;  // 8-byte words
    const int others = pt.world - 1;
    // (1) push this rank's rows of channel c into every peer's inbox: {value, epoch} words
    for (int i = t; i < others * n_local; i += kTableThreads) {
        int r = i / n_local;
        const int n = i - r * n_local;
        if (r >= pt.rank) ++r;
        const int64_t row = (int64_t)row_offset + n;
        uint2* dst = reinterpret_cast<uint2*>(pt.peers[r]) + p * words_per_parity + (size_t)row * 2 * C + c;
        st_ll(dst, mu_all[row * ld + c], epoch);
        st_ll(dst + C, sig_all[row * ld + c], epoch);
    }
    // (2) received words (local memory) -> local table, as they arrive
    const uint2* inbox = reinterpret_cast<const uint2*>(pt.peers[pt.rank]) + p * words_per_parity;
    for (int i = t; i < others * n_local; i += kTableThreads) {
        int r = i / n_local;
        const int n = i - r * n_local;
        if (r >= pt.rank) ++r;
        const int64_t row = (int64_t)r * n_local + n;
        const uint2* src = inbox + (size_t)row * 2 * C + c;
        float m = 0.f, sg = 0.f;
        const long long t0 = clock64();
        while (!(ld_ll(src, epoch, m) & ld_ll(src + C, epoch, sg))) {
            if (clock64() - t0 > 40000000000LL) wait_timed_out(pt.error);      // ~20 s: another rank may be late, not dead
        }
        mu_all[row * ld + c] = m;
        sig_all[row * ld + c] = sg;
    }
    __syncthreads();
    // (3) the table arithmetic on the complete channel
    tables_body(mu_all, sig_all, ld, n_global, row_offset, n_local, C, perm, lmda, gamma_noise, beta_noise, gamma_std, beta_std, flags,
                scale, shift);
    // (4) the last CTA out closes the epoch
    __syncthreads();
    if (t == 0) {
        __threadfence();
        if (atomicAdd(pt.done, 1u) == gridDim.x - 1u) {
            *pt.done = 0u;
            *pt.epoch = epoch;
            __threadfence();
        }
    }
}

// Rank barrier over the same peer buffers (one warp): every rank pushes an epoch-tagged word to every peer and spins on its own
// slots.  Launched right before the one-kernel multi-GPU forward so that the ranks START it within a microsecond or two of
// each other: the launch skew between ranks (they replay their graphs independently) is then spent idle here, before any
// streaming, instead of inside the forward where a channel finaliser waiting for a late rank lets x fall out of the L2 window.
// The barrier words live behind the inboxes: [2 parities][world] x {rank, epoch}.
__global__ void __launch_bounds__(32)
rank_barrier_kernel(PeerTables pt, size_t barrier_offset_words, unsigned int* bar_epoch) {
    const int t = threadIdx.x;
    const unsigned int epoch = *(volatile unsigned int*)bar_epoch + 1u;
    const int p = (int)(epoch & 1u);
    if (t < pt.world && t != pt.rank) {
        uint2* dst = reinterpret_cast<uint2*>(pt.peers[t]) + barrier_offset_words + (size_t)p * pt.world + pt.rank;
        st_ll(dst, __int_as_float(pt.rank), epoch);
        const uint2* src = reinterpret_cast<const uint2*>(pt.peers[pt.rank]) + barrier_offset_words + (size_t)p * pt.world + t;
        float who;
        const long long t0 = clock64();
        while (!ld_ll(src, epoch, who)) {
            if (clock64() - t0 > 40000000000LL) wait_timed_out(pt.error);      // ~20 s: another rank may be late, not dead
        }
    }
    __syncwarp();
    if (t == 0) *bar_epoch = epoch;
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

// rescale_intensity of the reference (src/common_utils/basic_operations.py:257-281) given the per-plane min / max that the forward
// kernels collected while writing y:  out = (y - min) / (max - min + eps) * (new_max - new_min) + new_min.  One pass (read y,
// write out) instead of the reference's two reductions + four elementwise kernels.
template <typename T>
__global__ void __launch_bounds__(256)
rescale_kernel(const T* __restrict__ y, const unsigned int* __restrict__ ymin, const unsigned int* __restrict__ ymax, T* __restrict__ out,
               float new_min, float new_max, float eps, int64_t planes, int64_t M) {
    const int64_t tiles = (M + 1023) / 1024;
    for (int64_t w = blockIdx.x; w < planes * tiles; w += gridDim.x) {
        const int64_t plane = w / tiles, tile = w - plane * tiles;
        const float lo = ordered_float(ymin[plane]), hi = ordered_float(ymax[plane]);
        const float denom = hi - lo + eps, span = new_max - new_min;
        const int64_t base = plane * M + tile * 1024;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int64_t e = tile * 1024 + threadIdx.x + 256 * i;
            if (e < M) out[base + threadIdx.x + 256 * i] = from_f32<T>((to_f32<T>(y[base + threadIdx.x + 256 * i]) - lo) / denom * span + new_min);
        }
    }
}

// Bumps the device-side step counter after step_kernel has finished (same stream).
__global__ void step_count_kernel(int* step_dev) { *step_dev += 1; }

}  // namespace ms
