// Fused forward: the whole MaxStyle forward (maxstyle.py:157-185) in ONE persistent kernel whose HBM
// traffic is the algorithmic minimum -- x read once, y written once.
//
// The layer needs every plane twice: once for its moments, once (after the style tables of its
// channel are known) to apply them.  The tensor (257 MB in the headline configuration) does not fit
// in the 126 MB L2, but a window of a few channels does.  So the work is cut into items -- a piece
// (<= ~64 KB) of one (n,c) plane, either "statistics" or "apply" -- laid out in one ORDERED queue,
// channel-major:   stats(ch 0) ... stats(ch D-1), then  stats(ch j), apply(ch j-D)  for j = D..C-1,
// then apply(ch C-D) ... apply(ch C-1).  CTAs take items with an atomic ticket, in order.
//   * statistics item: stream the piece (256-bit loads marked evict-last, so the lines stay in L2),
//     block-reduce its shifted moments, publish them, arrive on the channel's counter; the CTA whose
//     arrival completes a channel merges its planes in fixed order, computes mu/sig, the batch std
//     (first forward), the mixed + perturbed style A, B, writes the [N,C] tables and raises the
//     channel's ready flag (one warp does this while the CTA's other warps already stream the next item);
//   * apply item: wait for the channel's ready flag (raised ~D channels = tens of MB of traffic earlier),
//     read (mu, A/sig, B), stream the piece again -- out of L2 -- and write y (evict-first both ways).
// An item only ever waits for items that were handed out BEFORE it, i.e. that some running CTA already
// holds: no co-residency requirement, no cooperative launch, no deadlock.  D is sized so that the window
// (D channels of x) is ~32 MB.
#pragma once
#include "common.cuh"
#include "kernels_nchw.cuh"
#include "tables.cuh"

namespace ms {

constexpr long long kFusedSpinLimit = 2000000000LL;   // clock cycles (~1 s) before a wait gives up and flags an error

struct FusedArgs {
    int N, C;
    int64_t M;                 // elements per plane
    int nvec;                  // vectors per plane
    int pieces;                // pieces per plane (Kp)
    int piece_vecs;            // vectors per piece (the last piece of a plane may be shorter)
    int items_per_channel;     // N * Kp
    int window;                // D: channels between statistics and apply (<= C)
    int chunk;                 // consecutive items a CTA takes per ticket (1..kFusedMaxChunk)
    int keep_from;             // apply items of channels >= keep_from re-read x with evict-last: the tail of x stays in L2
                               // for the backward sweep that follows the forward (C: keep nothing)
    int64_t total_items;       // 2 * C * items_per_channel
    int flags;
    float eps;
    float *mu, *sig;                                 // [n_global, ld] statistics tables (single GPU: n_global == N, ld == C)
    float *scale, *shift;                            // [N, C] coefficients of the local rows, written by the channel finalisers
    int n_global, row_offset, ld;                    // this rank holds rows [row_offset, row_offset + N) of the global batch
    PeerTables pt;                                   // pt.world > 1: the finaliser exchanges the channel's rows over peer memory
    const int64_t* perm;
    const float *lmda, *gamma_noise, *beta_noise;
    float *gamma_std, *beta_std;                     // [C]: written when MAXSTYLE_COMPUTE_BATCH_STD, else read
    float4* partials;                                // [C * items_per_channel] (n, shifted mean, M2, shift K)
    unsigned int *arrived, *ready;                   // [C] each, zero between calls
    unsigned long long* queue;                       // ticket counter, zero between calls
    unsigned int* done;                              // CTAs that have left the loop, zero between calls
    int* error;                                      // set to 1 if a wait timed out (results are then invalid)
};

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned int* p, unsigned int v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int atom_add_release_u32(unsigned int* p, unsigned int v) {
    unsigned int old;
    asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}

struct FusedItem {
    int apply;       // 0: statistics, 1: apply
    int c, n, k;     // channel, sample, piece within the plane
};

// Ordered queue position -> item (see the file comment).  Host-callable: tests/host/queue_check.cu walks the whole queue.
MS_HD FusedItem fused_item(const FusedArgs& a, int64_t id) {
    const int64_t Ic = a.items_per_channel;
    const int64_t head = (int64_t)a.window * Ic;
    FusedItem it;
    int64_t i;
    if (id < head) {
        it.apply = 0;
        it.c = (int)(id / Ic);
        i = id - (int64_t)it.c * Ic;
    } else {
        const int64_t id2 = id - head;
        const int64_t mid = (int64_t)(a.C - a.window) * 2 * Ic;
        if (id2 < mid) {
            const int64_t step = id2 / (2 * Ic);
            const int64_t rem = id2 - step * 2 * Ic;
            if (rem < Ic) { it.apply = 0; it.c = a.window + (int)step; i = rem; }
            else { it.apply = 1; it.c = (int)step; i = rem - Ic; }
        } else {
            const int64_t id3 = id2 - mid;
            it.apply = 1;
            const int64_t cc = id3 / Ic;
            it.c = a.C - a.window + (int)cc;
            i = id3 - cc * Ic;
        }
    }
    it.n = (int)(i / a.pieces);
    it.k = (int)(i - (int64_t)it.n * a.pieces);
    return it;
}

// Runs in ONE WARP of the CTA whose arrival completed channel c: everything the reference does on the
// [N,C] tables for that channel (maxstyle.py:157-159 final merge, :165-168, :172-185).  Lanes own rows
// n = lane, lane+32, ...; fin_mu / fin_sig are that warp's shared scratch ([kFusedMaxN]).
__device__ __forceinline__ void fused_finalize_channel(const FusedArgs& a, int c, float* fin_mu, float* fin_sig) {
    const int lane = threadIdx.x & 31;
    const int N = a.N, C = a.C, Kp = a.pieces, NG = a.n_global, off = a.row_offset, ld = a.ld;
    const bool mix = a.flags & 1, no_noise = a.flags & 2, compute_std = a.flags & 4;
    const float inv_m1 = 1.0f / (float)(a.M - 1);
    const float4* part = a.partials + (int64_t)c * a.items_per_channel;
    float gs = 0.f, bs = 0.f;
    if (!compute_std && !no_noise) { gs = a.gamma_std[c]; bs = a.beta_std[c]; }
    // Multi-GPU: this rank's rows go to every peer's inbox as {value, epoch} words the moment they exist (tables.cuh).
    const bool multi = a.pt.world > 1;
    unsigned int epoch = 0;
    size_t words = 0;
    int par = 0;
    if (multi) {
        epoch = *(volatile unsigned int*)a.pt.epoch + 1u;
        par = (int)(epoch & 1u);
        words = (size_t)NG * 2 * C;
    }
    for (int n = lane; n < N; n += 32) {
        Moments m{0.f, 0.f, 0.f};
        float K = 0.f;
        for (int k0 = 0; k0 < Kp; k0 += 4) {                   // 4 independent loads, then the merges in piece order
            float4 p[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (k0 + j < Kp) p[j] = __ldcg(&part[n * Kp + k0 + j]);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (k0 + j < Kp) { m = merge(m, Moments{p[j].x, p[j].y, p[j].z}); K = p[j].w; }
        }
        const float mean = K + m.mean;
        const float sg = sqrtf(m.m2 * inv_m1 + a.eps);
        fin_mu[off + n] = mean;
        fin_sig[off + n] = sg;
        a.mu[(int64_t)(off + n) * ld + c] = mean;
        a.sig[(int64_t)(off + n) * ld + c] = sg;
        if (multi) {
            for (int r = 0; r < a.pt.world; ++r) {
                if (r == a.pt.rank) continue;
                uint2* dst = reinterpret_cast<uint2*>(a.pt.peers[r]) + par * words + (size_t)(off + n) * 2 * C + c;
                st_ll(dst, mean, epoch);
                st_ll(dst + C, sg, epoch);
            }
        }
    }
    if (multi) {
        // the other ranks' rows of this channel: spin on the tagged words in the local inbox
        const uint2* inbox = reinterpret_cast<const uint2*>(a.pt.peers[a.pt.rank]) + par * words;
        const int others = (a.pt.world - 1) * N;
        for (int i = lane; i < others; i += 32) {
            int r = i / N;
            const int n = i - r * N;
            if (r >= a.pt.rank) ++r;
            const int row = r * N + n;
            const uint2* src = inbox + (size_t)row * 2 * C + c;
            float m = 0.f, sg = 0.f;
            const long long t0 = clock64();
            while (!(ld_ll(src, epoch, m) & ld_ll(src + C, epoch, sg))) {
                if (clock64() - t0 > 20 * kFusedSpinLimit) wait_timed_out(a.error);      // another rank: ~20 s
            }
            fin_mu[row] = m;
            fin_sig[row] = sg;
            a.mu[(int64_t)row * ld + c] = m;
            a.sig[(int64_t)row * ld + c] = sg;
        }
    }
    __syncwarp();
    if (compute_std) {                                          // two-pass unbiased std over the GLOBAL batch (:165-168)
        float s_sig = 0.f, s_mu = 0.f;
        for (int n = lane; n < NG; n += 32) { s_sig += fin_sig[n]; s_mu += fin_mu[n]; }
        s_sig = warp_sum(s_sig);
        s_mu = warp_sum(s_mu);
        const float mean_sig = s_sig / (float)NG, mean_mu = s_mu / (float)NG;
        float q_sig = 0.f, q_mu = 0.f;
        for (int n = lane; n < NG; n += 32) {
            const float ds = fin_sig[n] - mean_sig, dm = fin_mu[n] - mean_mu;
            q_sig = fmaf(ds, ds, q_sig);
            q_mu = fmaf(dm, dm, q_mu);
        }
        q_sig = warp_sum(q_sig);
        q_mu = warp_sum(q_mu);
        gs = sqrtf(q_sig / (float)(NG - 1));
        bs = sqrtf(q_mu / (float)(NG - 1));
        if (lane == 0) { a.gamma_std[c] = gs; a.beta_std[c] = bs; }
    }
    for (int n = lane; n < N; n += 32) {
        const int row = off + n;
        const int64_t pr = mix ? a.perm[row] : row;
        float sc, sh;
        style_coeffs(fin_sig[row], fin_mu[row], fin_sig[pr], fin_mu[pr], mix, no_noise, mix ? a.lmda[n] : 0.f,
                     no_noise ? 0.f : a.gamma_noise[(int64_t)n * C + c], no_noise ? 0.f : a.beta_noise[(int64_t)n * C + c], gs, bs,
                     sc, sh, !(a.flags & 8));
        a.scale[(int64_t)n * C + c] = sc;
        a.shift[(int64_t)n * C + c] = sh;
    }
    __threadfence();
    __syncwarp();
    if (lane == 0) st_release_u32(&a.ready[c], 1u);
}

constexpr int kFusedStreamWarps = kThreads / 32 - 1;          // 7 warps stream, the 8th is the control warp
constexpr int kFusedStream = kFusedStreamWarps * 32;          // 224 streaming threads

// named barriers (id 0 is __syncthreads): GO[p] control -> streamers "item with parity p may start",
// TOT[p] streamers -> control "item with parity p is done (its moments are in sh.tot[p])", RED: streamers only.
enum { kBarGo0 = 1, kBarGo1 = 2, kBarTot0 = 3, kBarTot1 = 4, kBarRed = 5 };

__device__ __forceinline__ void named_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

constexpr int kFusedMaxChunk = 4;

struct FusedShared {
    long long item_id[2];            // first item of the unit with parity p (>= total_items: stop)
    float4 tot[2][kFusedMaxChunk];   // (n, shifted mean, M2, K) of the statistics items of the unit with parity p
    float red_n[kFusedStreamWarps], red_mean[kFusedStreamWarps], red_m2[kFusedStreamWarps];
    int flag;
    float fin_mu[kFusedMaxN], fin_sig[kFusedMaxN];
};

// Moments of the 7 streaming warps -> every streaming thread (fixed order).
__device__ __forceinline__ Moments stream_merge(Moments m, FusedShared& sh) {
    m = warp_merge(m);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    named_sync(kBarRed, kFusedStream);                          // red_* may still be read from the previous item
    if (lane == 0) { sh.red_n[warp] = m.n; sh.red_mean[warp] = m.mean; sh.red_m2[warp] = m.m2; }
    named_sync(kBarRed, kFusedStream);
    Moments r{0.f, 0.f, 0.f};
#pragma unroll
    for (int w = 0; w < kFusedStreamWarps; ++w) r = merge(r, Moments{sh.red_n[w], sh.red_mean[w], sh.red_m2[w]});
    return r;
}

// Warp specialisation: 7 warps only ever stream (no fences, no atomics, no polling); the control warp
// takes tickets, checks the ready flag of apply items ahead of time, publishes the moments of finished
// statistics items and finalises channels -- all of it off the streaming warps' critical path.
template <typename T, int VEC, int VPT>
__global__ void __launch_bounds__(kThreads, kBlocksPerSM)
fwd_fused_kernel(const T* __restrict__ x, T* __restrict__ y, FusedArgs a) {
    __shared__ FusedShared sh;
    constexpr int G = kFusedStream;
    constexpr int kAll = kThreads;
    const int t = threadIdx.x;

    if (t >= kFusedStream) {
        // =============================== control warp ===============================
        // A unit = `chunk` consecutive items taken with one ticket; lane b of this warp looks after item b of it.
        const int lane = t - kFusedStream;
        const int B = a.chunk;
        auto take = [&](int u) -> long long {                   // ticket of unit u -> its first item, published in sh.item_id
            long long id = 0;
            if (lane == 0) { id = (long long)atomicAdd(a.queue, 1ull) * B; sh.item_id[u & 1] = id; }
            return __shfl_sync(0xffffffffu, id, 0);
        };
        // may the streamers start this unit?  (statistics items: always; apply items: their channel's tables are ready)
        auto unit_ready = [&](long long first, bool block) -> bool {
            int ok = 1;
            const long long id = first + lane;
            if (lane < B && id < a.total_items) {
                const FusedItem it = fused_item(a, id);
                if (it.apply) {
                    ok = ld_acquire_u32(&a.ready[it.c]) != 0u;
                    if (!ok && block) {
                        const long long t0 = clock64();
                        while (!(ok = ld_acquire_u32(&a.ready[it.c]) != 0u)) {
                            __nanosleep(64);
                            if (clock64() - t0 > (a.pt.world > 1 ? 20 : 1) * kFusedSpinLimit) wait_timed_out(a.error);
                        }
                    }
                }
            }
            return __all_sync(0xffffffffu, ok) != 0;
        };
        auto go = [&](int u) { __syncwarp(); named_arrive((u & 1) ? kBarGo1 : kBarGo0, kAll); };

        long long cur = take(0);
        unit_ready(cur, true);                                  // nothing of this CTA is pending yet: blocking is safe
        go(0);
        for (int u = 0; cur < a.total_items; ++u) {
            // while the streamers work on unit u: ticket (and, if possible, clearance) for unit u+1
            const long long nxt = take(u + 1);
            const bool early = unit_ready(nxt, false);
            if (early) go(u + 1);
            // unit u is done: publish the moments of its statistics items, finalise the channels they complete
            named_sync((u & 1) ? kBarTot1 : kBarTot0, kAll);
            int last = 0, ch = 0;
            const long long id = cur + lane;
            if (lane < B && id < a.total_items) {
                const FusedItem it = fused_item(a, id);
                if (!it.apply) {
                    a.partials[(int64_t)it.c * a.items_per_channel + it.n * a.pieces + it.k] = sh.tot[u & 1][lane];
                    __threadfence();
                    last = atomicAdd(&a.arrived[it.c], 1u) == (unsigned int)a.items_per_channel - 1u;
                    if (last) __threadfence();
                    ch = it.c;
                }
            }
            __syncwarp();
            for (int b = 0; b < B; ++b) {
                const int lb = __shfl_sync(0xffffffffu, last, b), cb = __shfl_sync(0xffffffffu, ch, b);
                if (lb) fused_finalize_channel(a, cb, sh.fin_mu, sh.fin_sig);
            }
            if (!early) {                                       // the next unit waits for something that may have been ours
                unit_ready(nxt, true);
                go(u + 1);
            }
            cur = nxt;
        }
        // ---- leave the workspace zeroed: the last CTA out resets the queue and the channel flags ----
        int last_cta = 0;
        if (lane == 0) {
            __threadfence();
            last_cta = atomicAdd(a.done, 1u) == gridDim.x - 1u;
        }
        last_cta = __shfl_sync(0xffffffffu, last_cta, 0);
        if (last_cta) {
            for (int c = lane; c < a.C; c += 32) { a.arrived[c] = 0u; a.ready[c] = 0u; }
            if (lane == 0) {
                *a.queue = 0ull; *a.done = 0u;
                if (a.pt.world > 1) *a.pt.epoch = *(volatile unsigned int*)a.pt.epoch + 1u;     // close the exchange epoch
            }
        }
        return;
    }

    // =============================== streaming warps ===============================
    const uint64_t pol_keep = make_policy(kPolicyKeep), pol_stream = make_policy(kPolicyStream);
    for (int u = 0;; ++u) {
        named_sync((u & 1) ? kBarGo1 : kBarGo0, kAll);
        const long long first = sh.item_id[u & 1];
        if (first >= a.total_items) break;
      for (int ib = 0; ib < a.chunk && first + ib < a.total_items; ++ib) {
        const FusedItem it = fused_item(a, first + ib);
        const int64_t plane = (int64_t)it.n * a.C + it.c;
        Piece pc;
        pc.plane = plane;
        pc.v0 = it.k * a.piece_vecs;
        pc.v1 = min(a.nvec, pc.v0 + a.piece_vecs);
        const Batches<G, VPT> bt(pc, false);
        if (!it.apply) {
            // ---------------- statistics of the piece (same arithmetic as stats_nchw_kernel) ----------------
            const T* base = x + plane * a.M;
            const float K = to_f32<T>(__ldg(base));
            Moments acc{0.f, 0.f, 0.f};
            for (int b = 0; b < bt.full; ++b) {
                const T* p = base + (int64_t)(bt.begin(b) + t) * VEC;
                float val[VPT][VEC];
#pragma unroll
                for (int j = 0; j < VPT; ++j) Vec<T, VEC>::load(p + (int64_t)j * G * VEC, val[j], pol_keep);
                float s = 0.f;
#pragma unroll
                for (int j = 0; j < VPT; ++j)
#pragma unroll
                    for (int k = 0; k < VEC; ++k) { val[j][k] -= K; s += val[j][k]; }
                Moments bm;
                bm.n = (float)(VPT * VEC);
                bm.mean = s * (1.0f / (float)(VPT * VEC));
                float q = 0.f;
#pragma unroll
                for (int j = 0; j < VPT; ++j)
#pragma unroll
                    for (int k = 0; k < VEC; ++k) { const float d = val[j][k] - bm.mean; q = fmaf(d, d, q); }
                bm.m2 = q;
                acc = merge_fast(acc, bm);
            }
            if (bt.rem) {
                const int hi = bt.ragged_hi();
                for (int lo = bt.ragged_lo() + t; lo < hi; lo += G) {
                    float v[VEC];
                    Vec<T, VEC>::load(base + (int64_t)lo * VEC, v, pol_keep);
                    float s = 0.f;
#pragma unroll
                    for (int k = 0; k < VEC; ++k) { v[k] -= K; s += v[k]; }
                    Moments bm;
                    bm.n = (float)VEC;
                    bm.mean = s * (1.0f / (float)VEC);
                    float q = 0.f;
#pragma unroll
                    for (int k = 0; k < VEC; ++k) { const float d = v[k] - bm.mean; q = fmaf(d, d, q); }
                    bm.m2 = q;
                    acc = merge_fast(acc, bm);
                }
            }
            const Moments tot = stream_merge(acc, sh);
            if (t == 0) sh.tot[u & 1][ib] = make_float4(tot.n, tot.mean, tot.m2, K);
        } else {
            // ---------------- apply: the control warp has seen the channel's ready flag ----------------
            const float m = __ldcg(a.mu + ((int64_t)a.row_offset + it.n) * a.ld + it.c), sc = __ldcg(a.scale + plane), shf = __ldcg(a.shift + plane);
            const T* src = x + plane * a.M;
            T* dst = y + plane * a.M;
            const uint64_t pol_x = it.c >= a.keep_from ? pol_keep : pol_stream;
            for (int b = 0; b < bt.full; ++b) {
                const int64_t o = (int64_t)(bt.begin(b) + t) * VEC;
                float val[VPT][VEC];
#pragma unroll
                for (int j = 0; j < VPT; ++j) Vec<T, VEC>::load(src + o + (int64_t)j * G * VEC, val[j], pol_x);
#pragma unroll
                for (int j = 0; j < VPT; ++j) {
#pragma unroll
                    for (int k = 0; k < VEC; ++k) val[j][k] = fmaf(val[j][k] - m, sc, shf);
                    Vec<T, VEC>::store(dst + o + (int64_t)j * G * VEC, val[j], pol_stream);
                }
            }
            if (bt.rem) {
                const int hi = bt.ragged_hi();
                for (int lo = bt.ragged_lo() + t; lo < hi; lo += G) {
                    float v[VEC];
                    Vec<T, VEC>::load(src + (int64_t)lo * VEC, v, pol_x);
#pragma unroll
                    for (int k = 0; k < VEC; ++k) v[k] = fmaf(v[k] - m, sc, shf);
                    Vec<T, VEC>::store(dst + (int64_t)lo * VEC, v, pol_stream);
                }
            }
        }
      }
        named_arrive((u & 1) ? kBarTot1 : kBarTot0, kAll);       // hand the unit back to the control warp
    }
}

}  // namespace ms
