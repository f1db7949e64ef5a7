// Resident forward: the whole MaxStyle forward (maxstyle.py:157-185) in ONE persistent kernel in which every
// (n,c) plane crosses the SM boundary exactly once in each direction -- x is read from HBM once INTO SHARED
// MEMORY, its moments and the affine map are computed there, and y is written from there.
//
// A plane of the layer is needed twice: for its moments and, once the moments of its mixing partner
// (n' = perm[n], same channel) are known, for y = (x - mu)*A/sig + B.  The window-in-L2 kernel (fused_fwd.cuh)
// streams the plane twice and lets L2 serve the second read; here the plane (<= ~220 KB: 224x224 fp32 is 196 KB)
// simply stays in the SM's shared memory between the two uses:
//   * items = planes, channel-major (id -> c = id / N, n = id % N), handed out by an atomic ticket, so the
//     planes of one channel -- the only ones that depend on each other -- are in flight on neighbouring CTAs
//     at the same time;
//   * a PRODUCER warp (one elected thread) takes the ticket and moves the plane in with cp.async.bulk (TMA
//     1-D bulk copies) in <= 16 chunks of ~15 KB, each completing on its own mbarrier;
//   * the CONSUMER warps accumulate shifted moments chunk by chunk as the chunks land (so the reduction is
//     hidden under the load), merge them in fixed order, publish (mu, sig) + a ready flag for the plane, wait
//     for the partner's flag (first forward: for the whole channel, to take the batch std of maxstyle.py:165-168),
//     compute the style coefficients, then sweep the chunks again out of shared memory and store y with 16-byte
//     coalesced stores;  each warp releases a chunk (mbarrier `empty`) as soon as it has read it for the last
//     time, and the producer immediately refills it with the same chunk of the CTA's NEXT plane -- the load of
//     plane i+1 runs under the stores of plane i.
// Deadlock freedom: an item only waits for items of its own channel, i.e. at most N-1 positions away in the
// ticket order, and a CTA takes a ticket only when it can load that item at once (its previous plane is past
// its wait).  Take the lowest item that is waiting: everything before it is finished, every ticket handed out
// is loaded and published without waiting for anything, so if the item it waits for had no ticket yet, all
// CTAs would be holding waiting items that lie between the two -- fewer than N of them.  With a grid of >= N
// CTAs (all co-resident: the grid is sized from the occupancy query) some CTA is therefore free to take the
// next ticket.  The host
// only selects this kernel when 2 <= N <= grid; waits carry a clock64() bound that raises the workspace error
// flag instead of hanging.
#pragma once
#include "common.cuh"
#include "fused_fwd.cuh"
#include "kernels_nchw.cuh"

namespace ms {

constexpr int kResMaxChunks = 16;
constexpr int kResMaxN = 304;            // >= 148 SMs x 2 CTAs; rows a CTA stages for the batch std
constexpr int kResCtrlBytes = 4096;      // control block in front of the plane buffer

struct ResidentArgs {
    int N, C;
    int64_t M;                 // elements per plane
    int plane_bytes;
    int chunk_bytes;           // consumers * 16 * kUnroll
    int chunks;                // ceil(plane_bytes / chunk_bytes) <= kResMaxChunks
    int64_t total_items;       // N * C
    int flags;
    float eps;
    int in_policy, io_policy;
    int stagger_cycles, slot_div;                    // CTA b starts (b / slot_div) * stagger_cycles late (the CTAs of an SM out of phase)
    float *mu, *sig, *scale, *shift;                 // [N, C]
    const int64_t* perm;
    const float *lmda, *gamma_noise, *beta_noise;
    float *gamma_std, *beta_std;                     // [C]
    unsigned int* ready;                             // [N*C] plane flags, zero between calls
    unsigned long long* queue;                       // ticket counter, zero between calls
    unsigned int* done;                              // CTAs that have left, zero between calls
    int* error;
};

// ---- mbarrier / bulk-copy PTX ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: returns false (after raising *error) if the phase did not complete within the spin limit.
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, int* error) {
    if (mbar_try_wait(bar, parity)) return true;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > kFusedSpinLimit) wait_timed_out(error);
    }
    return true;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ Words4 lds128(const void* p) {
    Words4 r;
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]) : "r"(smem_u32(p)));
    return r;
}

template <int CW> struct ResidentCtrl {
    uint64_t full[kResMaxChunks];      // producer/TMA -> consumers: chunk landed
    uint64_t empty[kResMaxChunks];     // consumers -> producer: chunk may be overwritten (count = consumer warps)
    long long item_id[2];              // ticket of the item with parity p (-1: no more work)
    float red_n[CW], red_mean[CW], red_m2[CW];
    float coef[4];                     // mu, scale, shift of the current plane
    int flag;
    float fin_mu[kResMaxN], fin_sig[kResMaxN];
};

// One 16-byte vector of the plane -> fp32 values.
template <typename T> struct ResVec {
    static constexpr int kElems = 16 / (int)sizeof(T);
    static __device__ __forceinline__ void load(const char* p, float (&v)[kElems]) {
        const Words4 r = lds128(p);
        Unpack<T, 4>::to(r.w, v);
    }
};

template <typename T, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
fwd_resident_kernel(const T* __restrict__ x, T* __restrict__ y, ResidentArgs a) {
    constexpr int CW = THREADS / 32 - 1;          // consumer warps
    constexpr int TC = CW * 32;                   // consumer threads
    constexpr int VE = ResVec<T>::kElems;         // elements per 16-byte vector
    constexpr int U = THREADS >= 512 ? 2 : 4;     // vectors a consumer thread handles per full chunk
    extern __shared__ __align__(128) unsigned char smem_raw[];
    auto& sh = *reinterpret_cast<ResidentCtrl<CW>*>(smem_raw);
    static_assert(sizeof(ResidentCtrl<CW>) <= kResCtrlBytes, "control block too large");
    char* plane_buf = reinterpret_cast<char*>(smem_raw) + kResCtrlBytes;
    const int t = threadIdx.x;

    if (t == 0) {
        for (int i = 0; i < kResMaxChunks; ++i) { mbar_init(&sh.full[i], 1); mbar_init(&sh.empty[i], CW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (t >= TC) {
        // =============================== producer warp ===============================
        if (t == TC) {
            const uint64_t pol = make_policy(a.in_policy);
            if (a.stagger_cycles > 0) {                        // before the first ticket: nothing waits on a sleeping CTA
                const long long wait = (long long)(blockIdx.x / a.slot_div) * a.stagger_cycles, t0 = clock64();
                while (clock64() - t0 < wait) __nanosleep(200);
            }
            for (int p = 0;; ++p) {
                // The next ticket is taken only once chunk 0 of the current plane has been released, i.e. after the
                // consumers are past their wait: a CTA never holds a ticket it cannot start loading at once (a ticket
                // parked behind a waiting plane could be the very plane another CTA waits for -- a cycle).
                if (p > 0) mbar_wait(&sh.empty[0], (p - 1) & 1, a.error);
                const long long id = (long long)atomicAdd(a.queue, 1ull);
                const bool more = id < a.total_items;
                sh.item_id[p & 1] = more ? id : -1ll;
                if (!more) {                                   // stop signal rides on full[0]
                    mbar_arrive(&sh.full[0]);
                    break;
                }
                const int c = (int)(id / a.N), n = (int)(id - (long long)c * a.N);
                const char* src = reinterpret_cast<const char*>(x + ((int64_t)n * a.C + c) * a.M);
                for (int ch = 0; ch < a.chunks; ++ch) {
                    if (p > 0 && ch > 0) mbar_wait(&sh.empty[ch], (p - 1) & 1, a.error);
                    const int off = ch * a.chunk_bytes;
                    const uint32_t bytes = (uint32_t)min(a.chunk_bytes, a.plane_bytes - off);
                    mbar_arrive_expect_tx(&sh.full[ch], bytes);
                    bulk_g2s(plane_buf + off, src + off, bytes, &sh.full[ch], pol);
                }
            }
        }
        __syncwarp();                                          // lanes 1..31 wait here for the elected thread
    } else {
        // =============================== consumer warps ===============================
        const int warp = t >> 5, lane = t & 31;
        const uint64_t pol_out = make_policy(a.io_policy);
        const bool mix = a.flags & 1, no_noise = a.flags & 2, need_std = (a.flags & 4) && a.gamma_std != nullptr;     // the reference fills the cache whatever no_noise says (:165-168)
        const float inv_m1 = 1.0f / (float)(a.M - 1);
        const int N = a.N, C = a.C;
        for (int p = 0;; ++p) {
            const uint32_t par = p & 1;
            if (!mbar_wait(&sh.full[0], par, a.error)) break;
            const long long id = sh.item_id[par];
            if (id < 0) break;
            const int c = (int)(id / N), n = (int)(id - (long long)c * N);
            const int64_t plane = (int64_t)n * C + c;
            const float K = to_f32<T>(*reinterpret_cast<const T*>(plane_buf));
            // ---------------- moments, chunk by chunk as the chunks land ----------------
            Moments acc{0.f, 0.f, 0.f};
            bool ok = true;
            for (int ch = 0; ch < a.chunks; ++ch) {
                if (ch > 0 && !mbar_wait(&sh.full[ch], par, a.error)) { ok = false; break; }
                const int off = ch * a.chunk_bytes;
                const int nv = min(a.chunk_bytes, a.plane_bytes - off) >> 4;
                const char* base = plane_buf + off;
                if (nv == TC * U) {                                  // full chunk: U vectors per thread, one batch
                    float val[U][VE];
#pragma unroll
                    for (int j = 0; j < U; ++j) ResVec<T>::load(base + (size_t)(t + j * TC) * 16, val[j]);
                    float s = 0.f;
#pragma unroll
                    for (int j = 0; j < U; ++j)
#pragma unroll
                        for (int k = 0; k < VE; ++k) { val[j][k] -= K; s += val[j][k]; }
                    Moments b;
                    b.n = (float)(U * VE);
                    b.mean = s * (1.0f / (float)(U * VE));
                    float q = 0.f;
#pragma unroll
                    for (int j = 0; j < U; ++j)
#pragma unroll
                        for (int k = 0; k < VE; ++k) { const float d = val[j][k] - b.mean; q = fmaf(d, d, q); }
                    b.m2 = q;
                    acc = merge_fast(acc, b);
                } else {                                             // ragged last chunk: each vector is its own batch
                    for (int v = t; v < nv; v += TC) {
                        float val[VE];
                        ResVec<T>::load(base + (size_t)v * 16, val);
                        float s = 0.f;
#pragma unroll
                        for (int k = 0; k < VE; ++k) { val[k] -= K; s += val[k]; }
                        Moments b;
                        b.n = (float)VE;
                        b.mean = s * (1.0f / (float)VE);
                        float q = 0.f;
#pragma unroll
                        for (int k = 0; k < VE; ++k) { const float d = val[k] - b.mean; q = fmaf(d, d, q); }
                        b.m2 = q;
                        acc = merge_fast(acc, b);
                    }
                }
            }
            if (!ok) break;
            // ---------------- merge over the consumer warps (fixed order) ----------------
            acc = warp_merge(acc);
            named_sync(1, TC);                                       // red_* / coef may still be read from the previous item
            if (lane == 0) { sh.red_n[warp] = acc.n; sh.red_mean[warp] = acc.mean; sh.red_m2[warp] = acc.m2; }
            named_sync(1, TC);
            if (warp == 0) {
                // ---------------- warp 0: publish, wait for the planes this one depends on, coefficients ----------------
                Moments tot{0.f, 0.f, 0.f};
#pragma unroll
                for (int w = 0; w < CW; ++w) tot = merge(tot, Moments{sh.red_n[w], sh.red_mean[w], sh.red_m2[w]});
                const float mean = K + tot.mean;
                const float sg = sqrtf(tot.m2 * inv_m1 + a.eps);
                if (lane == 0) {
                    a.mu[plane] = mean;
                    a.sig[plane] = sg;
                    __threadfence();
                    st_release_u32(&a.ready[plane], 1u);
                }
                const int pr = mix ? (int)a.perm[n] : n;
                float gs = 0.f, bs = 0.f, sg_p = sg, mu_p = mean;
                if (need_std) {
                    // first forward: the whole channel (maxstyle.py:165-168), two-pass unbiased std in fixed order
                    const long long t0 = clock64();
                    for (int r = lane; r < N; r += 32) {
                        if (r == n) { sh.fin_mu[r] = mean; sh.fin_sig[r] = sg; continue; }
                        const int64_t q = (int64_t)r * C + c;
                        while (ld_acquire_u32(&a.ready[q]) == 0u) {
                            __nanosleep(64);
                            if (clock64() - t0 > kFusedSpinLimit) wait_timed_out(a.error);
                        }
                        sh.fin_mu[r] = __ldcg(a.mu + q);
                        sh.fin_sig[r] = __ldcg(a.sig + q);
                    }
                    __syncwarp();
                    float s_sig = 0.f, s_mu = 0.f;
                    for (int r = lane; r < N; r += 32) { s_sig += sh.fin_sig[r]; s_mu += sh.fin_mu[r]; }
                    s_sig = warp_sum(s_sig);
                    s_mu = warp_sum(s_mu);
                    const float mean_sig = s_sig / (float)N, mean_mu = s_mu / (float)N;
                    float q_sig = 0.f, q_mu = 0.f;
                    for (int r = lane; r < N; r += 32) {
                        const float ds = sh.fin_sig[r] - mean_sig, dm = sh.fin_mu[r] - mean_mu;
                        q_sig = fmaf(ds, ds, q_sig);
                        q_mu = fmaf(dm, dm, q_mu);
                    }
                    q_sig = warp_sum(q_sig);
                    q_mu = warp_sum(q_mu);
                    gs = sqrtf(q_sig / (float)(N - 1));
                    bs = sqrtf(q_mu / (float)(N - 1));
                    if (lane == 0 && n == 0) { a.gamma_std[c] = gs; a.beta_std[c] = bs; }
                    sg_p = sh.fin_sig[pr];
                    mu_p = sh.fin_mu[pr];
                    __syncwarp();                                    // fin_* are rewritten by the next item
                } else {
                    if (!no_noise) { gs = a.gamma_std[c]; bs = a.beta_std[c]; }
                    if (mix && pr != n && lane == 0) {
                        const int64_t q = (int64_t)pr * C + c;
                        const long long t0 = clock64();
                        while (ld_acquire_u32(&a.ready[q]) == 0u) {
                            __nanosleep(32);
                            if (clock64() - t0 > kFusedSpinLimit) wait_timed_out(a.error);
                        }
                        sg_p = __ldcg(a.sig + q);
                        mu_p = __ldcg(a.mu + q);
                    }
                }
                if (lane == 0) {
                    float sc, shf;
                    style_coeffs(sg, mean, sg_p, mu_p, mix, no_noise, mix ? a.lmda[n] : 0.f, no_noise ? 0.f : a.gamma_noise[plane],
                                 no_noise ? 0.f : a.beta_noise[plane], gs, bs, sc, shf, !(a.flags & 8));
                    a.scale[plane] = sc;
                    a.shift[plane] = shf;
                    sh.coef[0] = mean; sh.coef[1] = sc; sh.coef[2] = shf;
                }
            }
            named_sync(1, TC);
            const float m = sh.coef[0], sc = sh.coef[1], shf = sh.coef[2];
            // ---------------- apply out of shared memory; release each chunk to the producer ----------------
            T* dst = y + plane * a.M;
            for (int ch = 0; ch < a.chunks; ++ch) {
                const int off = ch * a.chunk_bytes;
                const int nv = min(a.chunk_bytes, a.plane_bytes - off) >> 4;
                const char* base = plane_buf + off;
                T* out = dst + (size_t)(off >> 4) * VE;
                if (nv == TC * U) {
                    float val[U][VE];
#pragma unroll
                    for (int j = 0; j < U; ++j) ResVec<T>::load(base + (size_t)(t + j * TC) * 16, val[j]);
#pragma unroll
                    for (int j = 0; j < U; ++j) {
#pragma unroll
                        for (int k = 0; k < VE; ++k) val[j][k] = fmaf(val[j][k] - m, sc, shf);
                        Vec<T, VE>::store(out + (size_t)(t + j * TC) * VE, val[j], pol_out);
                    }
                } else {
                    for (int v = t; v < nv; v += TC) {
                        float val[VE];
                        ResVec<T>::load(base + (size_t)v * 16, val);
#pragma unroll
                        for (int k = 0; k < VE; ++k) val[k] = fmaf(val[k] - m, sc, shf);
                        Vec<T, VE>::store(out + (size_t)v * VE, val, pol_out);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&sh.empty[ch]);
            }
        }
    }
    // ---- leave the workspace zeroed: the last CTA out resets the ticket counter and the plane flags ----
    __syncthreads();
    if (t == 0) {
        __threadfence();
        sh.flag = atomicAdd(a.done, 1u) == gridDim.x - 1u;
    }
    __syncthreads();
    if (sh.flag) {
        __threadfence();
        for (int64_t i = t; i < a.total_items; i += THREADS) a.ready[i] = 0u;
        if (t == 0) { *a.queue = 0ull; *a.done = 0u; }
    }
}

}  // namespace ms
