// Streamed forward: the L2-window forward of fused_fwd.cuh (same ordered statistics/apply queue, same channel
// finaliser, same results) with the data path rebuilt around a TMA ring.
//
// The register-staged window kernel streams an item (a <= 56 KB piece of a plane) with 2-3 batches of vector loads per
// thread: every item pays a full memory latency to ramp up and another to drain, and a warp that is folding moments has
// nothing in flight.  Here one elected PRODUCER thread per CTA walks the ticket queue and keeps a ring of S chunks
// (14 KB each) full with cp.async.bulk copies -- across item boundaries, and for apply items before their channel's
// tables are even ready, because the load of x never depends on them.  The CONSUMER warps take the chunks in order out
// of shared memory: statistics chunks are folded into shifted moments, apply chunks go through one FMA and out with
// 16-byte stores; a chunk is handed back to the producer (mbarrier) the moment a warp has read it.  Whatever a consumer
// waits for (the ready flag of an apply item, the block reduction at the end of a statistics item, once per channel the
// finaliser) is covered by the chunks already in flight.
// Ordering and deadlock freedom are those of the window kernel: an item depends only on items with LOWER tickets and a
// CTA works through its tickets in order, so the lowest unfinished item can always proceed; nothing has to be co-resident.
#pragma once
#include "common.cuh"
#include "fused_fwd.cuh"
#include "resident_fwd.cuh"

namespace ms {

constexpr int kRingMaxStages = 8;
constexpr int kRingChunkBytes = kFusedStream * 16 * 4;       // 14336: 4 x 16 bytes per consumer thread
constexpr int kRingCtrlBytes = 12288;                          // control block in front of the ring

struct RingGeom {
    int stages;                // S <= kRingMaxStages
    int plane_bytes, piece_bytes;   // piece_bytes is a multiple of kRingChunkBytes
};

struct RingDesc {
    long long item;            // ticket (-1: stop)
    int chunk;                 // chunk index inside the item
    int bytes;                 // payload of this chunk (multiple of 16)
    int last;                  // last chunk of the item
    int pad;
};

struct RingCtrl {
    uint64_t full[kRingMaxStages];
    uint64_t empty[kRingMaxStages];
    RingDesc desc[kRingMaxStages];
    float red_n[kFusedStreamWarps], red_mean[kFusedStreamWarps], red_m2[kFusedStreamWarps];
    int flag;
    float fin_mu[kFusedMaxN], fin_sig[kFusedMaxN];
};
static_assert(sizeof(RingCtrl) <= kRingCtrlBytes, "ring control block too large");

template <typename T>
__global__ void __launch_bounds__(kThreads, 3)
fwd_ring_kernel(const T* __restrict__ x, T* __restrict__ y, FusedArgs a, RingGeom rg) {
    constexpr int TC = kFusedStream;              // 224 consumer threads, warp 7 is the producer
    constexpr int VE = ResVec<T>::kElems;
    constexpr int U = 4;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    auto& sh = *reinterpret_cast<RingCtrl*>(smem_raw);
    char* ring = reinterpret_cast<char*>(smem_raw) + kRingCtrlBytes;
    const int t = threadIdx.x;
    const int S = rg.stages;

    if (t == 0) {
        for (int i = 0; i < kRingMaxStages; ++i) { mbar_init(&sh.full[i], 1); mbar_init(&sh.empty[i], kFusedStreamWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (t >= TC) {
        // =============================== producer ===============================
        if (t == TC) {
            const uint64_t pol_keep = make_policy(kPolicyKeep), pol_stream = make_policy(kPolicyStream);
            int slot = 0;
            uint32_t round = 0;                               // how many times the ring has wrapped
            auto acquire = [&]() {                             // slot is free again (its previous use was released)
                if (round > 0) mbar_wait(&sh.empty[slot], (round - 1) & 1, a.error);
            };
            auto advance = [&]() { if (++slot == S) { slot = 0; ++round; } };
            for (;;) {
                const long long id = (long long)atomicAdd(a.queue, 1ull);
                if (id >= a.total_items) {
                    acquire();
                    sh.desc[slot].item = -1;
                    mbar_arrive(&sh.full[slot]);
                    break;
                }
                const FusedItem it = fused_item(a, id);
                const int64_t plane = (int64_t)it.n * a.C + it.c;
                const int off0 = it.k * rg.piece_bytes;
                const int len = min(rg.piece_bytes, rg.plane_bytes - off0);
                const char* src = reinterpret_cast<const char*>(x + plane * a.M) + off0;
                const uint64_t pol = (it.apply && it.c < a.keep_from) ? pol_stream : pol_keep;   // statistics pass leaves x in L2 for the apply pass; the tail stays for the backward
                for (int off = 0, ch = 0; off < len; off += kRingChunkBytes, ++ch) {
                    acquire();
                    const int bytes = min(kRingChunkBytes, len - off);
                    RingDesc& d = sh.desc[slot];
                    d.item = id; d.chunk = ch; d.bytes = bytes; d.last = off + kRingChunkBytes >= len;
                    mbar_arrive_expect_tx(&sh.full[slot], (uint32_t)bytes);
                    bulk_g2s(ring + (size_t)slot * kRingChunkBytes, src + off, (uint32_t)bytes, &sh.full[slot], pol);
                    advance();
                }
            }
        }
        __syncwarp();
    } else {
        // =============================== consumers ===============================
        const int warp = t >> 5, lane = t & 31;
        const uint64_t pol_out = make_policy(kPolicyStream);
        int slot = 0;
        uint32_t round = 0;
        // state of the item being consumed
        int apply = 0, ch_c = 0, ch_n = 0, ch_k = 0;
        int64_t plane = 0;
        float K = 0.f, m = 0.f, sc = 0.f, shf = 0.f;
        Moments acc{0.f, 0.f, 0.f};
        T* dst = nullptr;
        for (;;) {
            if (!mbar_wait(&sh.full[slot], round & 1, a.error)) break;
            const RingDesc d = sh.desc[slot];
            if (d.item < 0) break;
            if (d.chunk == 0) {
                const FusedItem it = fused_item(a, d.item);
                apply = it.apply; ch_c = it.c; ch_n = it.n; ch_k = it.k;
                plane = (int64_t)it.n * a.C + it.c;
                if (apply) {
                    if (t == 0) {                                // raised ~`window` channels of traffic ago
                        const long long t0 = clock64();
                        while (ld_acquire_u32(&a.ready[it.c]) == 0u) {
                            __nanosleep(64);
                            if (clock64() - t0 > kFusedSpinLimit) wait_timed_out(a.error);
                        }
                    }
                    named_sync(kBarRed, TC);
                    m = __ldcg(a.mu + ((int64_t)a.row_offset + it.n) * a.ld + it.c); sc = __ldcg(a.scale + plane); shf = __ldcg(a.shift + plane);
                    dst = y + plane * a.M + (size_t)(it.k * rg.piece_bytes) / sizeof(T);
                } else {
                    K = to_f32<T>(__ldg(x + plane * a.M));      // the plane's first element: common shift of all its pieces
                    acc = Moments{0.f, 0.f, 0.f};
                }
            }
            const int nv = d.bytes >> 4;
            const char* base = ring + (size_t)slot * kRingChunkBytes;
            if (!apply) {
                if (nv == TC * U) {
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
                        for (int k = 0; k < VE; ++k) { const float dd = val[j][k] - b.mean; q = fmaf(dd, dd, q); }
                    b.m2 = q;
                    acc = merge_fast(acc, b);
                } else {
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
                        for (int k = 0; k < VE; ++k) { const float dd = val[k] - b.mean; q = fmaf(dd, dd, q); }
                        b.m2 = q;
                        acc = merge_fast(acc, b);
                    }
                }
            } else {
                T* out = dst + (size_t)d.chunk * (kRingChunkBytes / sizeof(T));
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
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sh.empty[slot]);
            if (d.last && !apply) {
                // ---- end of a statistics item: merge the warps (fixed order), publish, maybe finalise the channel ----
                Moments w = warp_merge(acc);
                named_sync(kBarRed, TC);                         // red_* may still be read from the previous item
                if (lane == 0) { sh.red_n[warp] = w.n; sh.red_mean[warp] = w.mean; sh.red_m2[warp] = w.m2; }
                named_sync(kBarRed, TC);
                if (warp == 0) {
                    int last = 0;
                    if (lane == 0) {
                        Moments tot{0.f, 0.f, 0.f};
#pragma unroll
                        for (int i = 0; i < kFusedStreamWarps; ++i) tot = merge(tot, Moments{sh.red_n[i], sh.red_mean[i], sh.red_m2[i]});
                        a.partials[(int64_t)ch_c * a.items_per_channel + ch_n * a.pieces + ch_k] = make_float4(tot.n, tot.mean, tot.m2, K);
                        __threadfence();
                        last = atomicAdd(&a.arrived[ch_c], 1u) == (unsigned int)a.items_per_channel - 1u;
                        if (last) __threadfence();
                    }
                    last = __shfl_sync(0xffffffffu, last, 0);
                    if (last) fused_finalize_channel(a, ch_c, sh.fin_mu, sh.fin_sig);
                }
            }
            if (++slot == S) { slot = 0; ++round; }
        }
    }
    // ---- leave the workspace zeroed: the last CTA out resets the queue and the channel flags ----
    __syncthreads();
    if (t == 0) {
        __threadfence();
        sh.flag = atomicAdd(a.done, 1u) == gridDim.x - 1u;
    }
    __syncthreads();
    if (sh.flag) {
        for (int c = t; c < a.C; c += kThreads) { a.arrived[c] = 0u; a.ready[c] = 0u; }
        if (t == 0) { *a.queue = 0ull; *a.done = 0u; }
    }
}

}  // namespace ms
