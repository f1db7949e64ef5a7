// Cluster-resident forward: the whole MaxStyle forward (maxstyle.py:157-185) in ONE persistent kernel in which every
// (n,c) plane crosses the SM boundary exactly once in each direction AND every SM always has a plane loading, a plane
// waiting for its style coefficients and a plane being written out.
//
// resident_fwd.cuh keeps ONE plane per CTA in shared memory: load -> moments -> publish -> partner -> store is a serial
// chain per SM, and with 196 KB planes (224x224 fp32, one CTA per SM) nothing hides the middle of it (measured 119 us
// against 79 us of pure traffic).  Here a plane is split over a thread-block CLUSTER of CS CTAs (CS SMs), each holding
// 1/CS of it, so that one CTA's shared memory holds S >= 2 parts of DIFFERENT planes in a ring of stages:
//     stage s+2: TMA bulk copies of the next plane's part are landing           (producer thread)
//     stage s+1: moments taken, partials sent to the cluster's leader over DSMEM, leader publishes (mu, sig), fetches the
//                mixing partner's, computes the style coefficients and broadcasts them to the cluster   (coordinator lanes)
//     stage s  : y = (x - mu) * A/sig + B streamed out of shared memory         (apply warps)
// Roles are warp-specialised and run through the ring at their own pace, linked only by mbarriers:
//     warp 0       producer: one thread issues cp.async.bulk (TMA 1-D) per chunk as soon as the apply warps release it
//     warp 1       coordinator (leader CTA only): lane s owns stage s -- merges the CS partials (fixed order), publishes,
//                  polls the partner's {value, tag} words, style_coeffs, st.shared::cluster + remote mbarrier arrive
//     warps 2-5    moments: shifted two-pass batches folded with Chan merges as chunks land; one partial per CTA
//     warps 6-13   apply: LDS.128 -> FMA -> STG.128, releasing each chunk to the producer as soon as it is read
// Items (planes) are assigned statically: cluster g owns items g, g+G, g+2G, ... of the channel-major order, so the N planes
// of a channel -- the only ones that depend on each other -- are in flight on neighbouring clusters at the same time.
// Statistics travel as 8-byte {value, tag} words (tag = launch number): one store publishes, one load both tests and fetches,
// nothing has to be zeroed between launches; in the multi-GPU layer the same words are pushed into every peer's inbox over
// NVLink, so the kernel IS the all-gather of the (mu | sig) rows.
// Deadlock freedom: an item's statistics are published without waiting for anything once its part has landed; a part lands
// as soon as its stage is free; a stage is freed when its previous item has its coefficients.  With 2 <= N <= G (host-side
// condition) the planes of a channel sit in one round, or in two adjacent rounds on disjoint clusters, and the waits form no
// cycle.  With N > G the host passes `order`: samples in cycle order of perm, so an item waits only for
// the NEXT item (or an earlier one) -- any N works, at one idle part-time per trip round the clusters when S == 1.
#pragma once
#include "common.cuh"
#include "kernels_nchw.cuh"
#include "tables.cuh"
#include "fused_fwd.cuh"
#include "resident_fwd.cuh"

namespace ms {

constexpr int kClMaxStages = 6;
constexpr int kClMaxChunks = 16;
constexpr int kClMaxCluster = 8;
constexpr int kClMomentWarps = 4;
constexpr int kClApplyWarps = 8;
constexpr int kClCoordWarps = kClMaxStages;                               // one coordinator warp per stage
constexpr int kClFirstMomentWarp = 1 + kClCoordWarps, kClFirstApplyWarp = kClFirstMomentWarp + kClMomentWarps;
constexpr int kClThreads = 32 * (kClFirstApplyWarp + kClApplyWarps);      // 608
constexpr int kClMT = 32 * kClMomentWarps, kClAT = 32 * kClApplyWarps;
constexpr int kClMaxPieces = 32;         // pieces of a plane (one lane of a coordinator warp polls one piece)
constexpr int kClStdRows = 16;           // first forward: rows of the channel a lane keeps in registers (n_global <= 512)
constexpr int kClCtrlBytes = 16384;      // control block in front of the stage buffers
constexpr int kClMaxN = 1024;            // rows the coordinator stages for the batch std / the order table
constexpr long long kClSpinLocal = 4000000000LL;     // ~2 s: waits on this GPU
constexpr long long kClSpinPeer = 40000000000LL;     // ~20 s: waits on another rank

struct ClusterArgs {
    int N, C;
    int64_t M;
    int plane_bytes;
    int pieces;                // P: a plane is cut into P pieces, each an item of its own (moments merged through L2)
    int part_bytes;            // bytes of a piece one CTA holds (multiple of 16); a piece is cluster * part_bytes
    int part_stride;           // bytes between stage buffers
    int chunk_bytes, chunks;   // a part is moved and released in `chunks` pieces
    int stages;                // S
    int cluster;               // CS
    int num_clusters;          // G
    int use_order;             // samples are visited in cycle order of perm (N > G)
    int64_t total_items;       // N * C * P
    int flags;
    float eps;
    int in_policy, io_policy;
    float *mu, *sig;           // [n_global, ld]
    float *scale, *shift;      // [N, C]
    int n_global, row_offset, ld;
    const int64_t* perm;
    const float *lmda, *gamma_noise, *beta_noise;
    float *gamma_std, *beta_std;
    uint2* ll;                 // [n_global][2][C] {value bits, tag} words (single GPU: workspace; multi GPU: own inbox, both parities)
    uint2* piece_ll;           // [N*C][P][2] {shifted mean | M2, tag} words of the pieces (P > 1)
    unsigned int* epoch;       // launch counter the tag comes from (multi GPU: the exchange epoch)
    unsigned int* done;
    int* error;
    PeerTables pt;             // pt.world > 1: push the words into every peer's inbox too
};

// ---- cluster PTX ----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(const void* local, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(local)), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_f4(uint32_t raddr, float a, float b, float c, float d) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(raddr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t rbar) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(rbar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// A wait that gave up means a broken launch (a peer rank died, the grid is not co-resident): results would be
// wrong, so the kernel raises the error word and traps -- the CUDA error reaches the host at its next call.
__device__ __forceinline__ void cl_fail(int* error) { wait_timed_out(error); }
__device__ __forceinline__ void mbar_wait_cl(uint64_t* bar, uint32_t parity, int* error) {
    if (mbar_try_wait_cluster(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait_cluster(bar, parity)) {
        if (clock64() - t0 > kClSpinPeer) cl_fail(error);
    }
}
__device__ __forceinline__ void mbar_wait_t(uint64_t* bar, uint32_t parity, int* error) {      // CTA-local barrier
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > kClSpinPeer) cl_fail(error);
    }
}
__device__ __forceinline__ void st_ll_gpu(void* p, float v, unsigned int tag) {
    asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(v)), "r"(tag) : "memory");
}

struct ClusterCtrl {
    uint64_t full[kClMaxStages][kClMaxChunks];    // TMA -> moments / apply warps: chunk landed
    uint64_t empty[kClMaxStages][kClMaxChunks];   // apply warps -> producer: chunk may be overwritten
    uint64_t stats_bar[kClMaxStages];             // leader: the CS partials of the stage's item have arrived (remote arrives)
    uint64_t coef_bar[kClMaxStages];              // every CTA: the leader wrote coef[stage]
    float4 part_in[kClMaxStages][kClMaxCluster];  // leader: (n, shifted mean, M2, K) from each CTA of the cluster
    float4 coef[kClMaxStages];                    // (mu, scale, shift, -)
    float red_n[kClMomentWarps], red_mean[kClMomentWarps], red_m2[kClMomentWarps];
    unsigned short order[kClMaxN];                // cycle order of perm (use_order)
    int next_of[kClMaxN];                         // scratch while the order is built
};
static_assert(sizeof(ClusterCtrl) <= kClCtrlBytes, "cluster control block too large");

// item id -> (channel, sample, piece); items are channel-major, the pieces of a plane adjacent
struct ClItem { int c, n, p; };
__device__ __forceinline__ ClItem cl_item(const ClusterArgs& a, const ClusterCtrl& sh, long long id) {
    ClItem it;
    const int per_channel = a.N * a.pieces;
    it.c = (int)(id / per_channel);
    const int r = (int)(id - (long long)it.c * per_channel);
    const int k = r / a.pieces;
    it.p = r - k * a.pieces;
    it.n = a.use_order ? (int)sh.order[k] : k;
    return it;
}
// bytes of piece p that CTA q of the cluster holds: [off, off + bytes) of the plane
__device__ __forceinline__ void cl_range(const ClusterArgs& a, int p, int q, int& off, int& bytes) {
    off = (p * a.cluster + q) * a.part_bytes;
    bytes = max(0, min(a.part_bytes, a.plane_bytes - off));
}

template <typename T>
__device__ __forceinline__ void cl_moments_chunk(const char* base, int nv, int mt, float K, Moments& acc) {
    constexpr int VE = ResVec<T>::kElems;
    constexpr int U = 4;
    int v = mt;
    for (; v + (U - 1) * kClMT < nv; v += U * kClMT) {
        float val[U][VE];
#pragma unroll
        for (int j = 0; j < U; ++j) ResVec<T>::load(base + (size_t)(v + j * kClMT) * 16, val[j]);
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
    }
    for (; v < nv; v += kClMT) {
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

template <typename T>
__device__ __forceinline__ void cl_apply_chunk(const char* base, T* out, int nv, int at, float m, float sc, float shf, uint64_t pol) {
    constexpr int VE = ResVec<T>::kElems;
    constexpr int U = 2;                                   // 2 x 256 threads x 16 B = one 8 KB chunk unit
    int v = at;
    for (; v + (U - 1) * kClAT < nv; v += U * kClAT) {
        float val[U][VE];
#pragma unroll
        for (int j = 0; j < U; ++j) ResVec<T>::load(base + (size_t)(v + j * kClAT) * 16, val[j]);
#pragma unroll
        for (int j = 0; j < U; ++j) {
#pragma unroll
            for (int k = 0; k < VE; ++k) val[j][k] = fmaf(val[j][k] - m, sc, shf);
            Vec<T, VE>::store(out + (size_t)(v + j * kClAT) * VE, val[j], pol);
        }
    }
    for (; v < nv; v += kClAT) {
        float val[VE];
        ResVec<T>::load(base + (size_t)v * 16, val);
#pragma unroll
        for (int k = 0; k < VE; ++k) val[k] = fmaf(val[k] - m, sc, shf);
        Vec<T, VE>::store(out + (size_t)v * VE, val, pol);
    }
}

// Poll a pair of {value, tag} words until both carry `tag`.
__device__ __forceinline__ void cl_poll_pair(const uint2* w0, const uint2* w1, unsigned int tag, bool remote, int* error, float& v0, float& v1) {
    if (ld_ll(w0, tag, v0) & ld_ll(w1, tag, v1)) return;
    const long long t0 = clock64();
    const long long limit = remote ? kClSpinPeer : kClSpinLocal;
    while (!(ld_ll(w0, tag, v0) & ld_ll(w1, tag, v1))) {
        __nanosleep(20);
        if (clock64() - t0 > limit) cl_fail(error);
    }
}

template <typename T>
__global__ void __launch_bounds__(kClThreads, 1)
fwd_cluster_kernel(const T* __restrict__ x, T* __restrict__ y, ClusterArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    auto& sh = *reinterpret_cast<ClusterCtrl*>(smem_raw);
    char* stage_buf = reinterpret_cast<char*>(smem_raw) + kClCtrlBytes;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int S = a.stages, NCH = a.chunks, CS = a.cluster, G = a.num_clusters, P = a.pieces;
    const uint32_t q = cluster_ctarank();
    const long long g = (long long)cluster_id_x();
    const int my_items = g < a.total_items ? (int)((a.total_items - g + G - 1) / G) : 0;
    // tag of this launch: the counter is advanced by the last CTA out, i.e. after every CTA has read it
    const unsigned int tag = *(volatile unsigned int*)a.epoch + 1u;
    const bool multi = a.pt.world > 1;
    uint2* ll_mine = a.ll;                                                     // multi GPU: the parity half of the own inbox
    size_t ll_words = 0;
    if (multi) {
        ll_words = (size_t)a.n_global * 2 * a.C;
        ll_mine = reinterpret_cast<uint2*>(a.pt.peers[a.pt.rank]) + (tag & 1u) * ll_words;
    }

    if (t == 0) {
        for (int s = 0; s < S; ++s) {
            for (int ch = 0; ch < NCH; ++ch) { mbar_init(&sh.full[s][ch], 1); mbar_init(&sh.empty[s][ch], kClApplyWarps); }
            mbar_init(&sh.stats_bar[s], CS);
            mbar_init(&sh.coef_bar[s], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (a.use_order) {
        // samples in cycle order of perm: n0, perm[n0], perm[perm[n0]], ... -- the plane an item waits for is the next in order
        for (int i = t; i < a.N; i += kClThreads) sh.next_of[i] = (int)a.perm[a.row_offset + i] - a.row_offset;
        __syncthreads();
        if (t == 0) {
            int k = 0;
            unsigned int seen[kClMaxN / 32];                  // one bit per sample
#pragma unroll
            for (int w = 0; w < kClMaxN / 32; ++w) seen[w] = 0u;
            for (int s0 = 0; s0 < a.N; ++s0) {
                int n = s0;
                while (!((seen[n >> 5] >> (n & 31)) & 1u)) {
                    seen[n >> 5] |= 1u << (n & 31);
                    sh.order[k++] = (unsigned short)n;
                    n = sh.next_of[n];
                }
            }
        }
    }
    __syncthreads();
    cluster_sync_all();                                   // every CTA's barriers exist before anyone arrives remotely

    if (warp == 0) {
        // =============================== producer ===============================
        if (lane == 0) {
            const uint64_t pol = make_policy(a.in_policy);
            for (int j = 0; j < my_items; ++j) {
                const int s = j % S;
                const uint32_t prev = (uint32_t)((j / S - 1) & 1);
                const ClItem it = cl_item(a, sh, g + (long long)j * G);
                int my_off, my_bytes;
                cl_range(a, it.p, (int)q, my_off, my_bytes);
                const char* src = reinterpret_cast<const char*>(x + ((int64_t)it.n * a.C + it.c) * a.M) + my_off;
                char* dst = stage_buf + (size_t)s * a.part_stride;
                for (int ch = 0; ch < NCH; ++ch) {
                    if (j >= S) mbar_wait_t(&sh.empty[s][ch], prev, a.error);
                    const int off = ch * a.chunk_bytes;
                    const int bytes = max(0, min(a.chunk_bytes, my_bytes - off));
                    if (bytes > 0) {
                        mbar_arrive_expect_tx(&sh.full[s][ch], (uint32_t)bytes);
                        bulk_g2s(dst + off, src + off, (uint32_t)bytes, &sh.full[s][ch], pol);
                    } else {
                        mbar_arrive(&sh.full[s][ch]);
                    }
                }
            }
        }
    } else if (warp < kClFirstMomentWarp) {
        // =============================== coordinator warps (leader CTA): warp 1+s owns stage s ===============================
        const int s = warp - 1;
        if (q == 0 && s < S) {
            const bool mix = a.flags & 1, no_noise = a.flags & 2, compute_std = a.flags & 4;
            const float inv_m1 = 1.0f / (float)(a.M - 1);
            const int lo = a.row_offset, hi = a.row_offset + a.N;
            const int NG = a.n_global;
            constexpr int ES = (int)sizeof(T);
            if (multi && tag > 1u) {
                // Flow control for the two-parity inboxes: this launch writes the words of launch tag-2 over.  A peer that has
                // published anything in launch tag-1 has finished launch tag-2, so wait for one word of launch tag-1 from each
                // peer (its first row, last channel: every exchange kernel publishes it) before the first push.
                const uint2* prev_inbox = reinterpret_cast<const uint2*>(a.pt.peers[a.pt.rank]) + ((tag - 1u) & 1u) * ll_words;
                for (int r = lane; r < a.pt.world; r += 32) {
                    if (r == a.pt.rank) continue;
                    const uint2* src = prev_inbox + ((size_t)(r * a.N) * 2) * a.C + (a.C - 1);
                    const long long t0 = clock64();
                    for (;;) {
                        unsigned int bits, got;
                        asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(bits), "=r"(got) : "l"(src) : "memory");
                        if ((int)(got - (tag - 1u)) >= 0) break;
                        __nanosleep(100);
                        if (clock64() - t0 > kClSpinPeer) cl_fail(a.error);
                    }
                }
                __syncwarp();
            }
            for (int j = s; j < my_items; j += S) {
                const ClItem it = cl_item(a, sh, g + (long long)j * G);
                const int64_t plane = (int64_t)it.n * a.C + it.c;
                const int row = lo + it.n;
                // everything that does not depend on the moments is requested before the wait
                const int prow = mix ? (int)a.perm[row] : row;
                const float lm = mix ? a.lmda[it.n] : 0.f;
                float gn = 0.f, bn = 0.f, gs = 0.f, bs = 0.f;
                if (!no_noise) {
                    gn = a.gamma_noise[plane]; bn = a.beta_noise[plane];
                    if (!compute_std) { gs = a.gamma_std[it.c]; bs = a.beta_std[it.c]; }
                }
                mbar_wait_cl(&sh.stats_bar[s], (uint32_t)((j / S) & 1), a.error);
                // ---- this piece: the CS partials in CTA order ----
                Moments m{0.f, 0.f, 0.f};
                float K = 0.f;
                for (int r = 0; r < CS; ++r) {
                    const float4 pp = sh.part_in[s][r];
                    m = merge(m, Moments{pp.x, pp.y, pp.z});
                    K = pp.w;
                }
                if (P > 1) {
                    // ---- the plane: publish this piece, fetch the others (lane l: piece l), merge in piece order ----
                    uint2* mine = a.piece_ll + ((size_t)plane * P + it.p) * 2;
                    if (lane == 0) { st_ll_gpu(mine, m.mean, tag); st_ll_gpu(mine + 1, m.m2, tag); }
                    float pm = m.mean, pq = m.m2;
                    if (lane < P && lane != it.p) {
                        const uint2* w = a.piece_ll + ((size_t)plane * P + lane) * 2;
                        cl_poll_pair(w, w + 1, tag, false, a.error, pm, pq);
                    }
                    __syncwarp();
                    Moments tot{0.f, 0.f, 0.f};
                    const int piece_bytes = CS * a.part_bytes;
                    for (int l = 0; l < P; ++l) {
                        const float lmean = __shfl_sync(0xffffffffu, pm, l), lm2 = __shfl_sync(0xffffffffu, pq, l);
                        const int pb = min(piece_bytes, a.plane_bytes - l * piece_bytes);
                        tot = merge(tot, Moments{(float)(pb / ES), lmean, lm2});
                    }
                    m = tot;
                }
                const float mean = K + m.mean;
                const float sg = sqrtf(m.m2 * inv_m1 + a.eps);
                if (it.p == 0) {
                    // ---- publish the plane's statistics (and push them to the peers) ----
                    uint2* w = ll_mine + ((size_t)row * 2) * a.C + it.c;
                    if (multi) {
                        if (lane == 0) { st_ll(w, mean, tag); st_ll(w + a.C, sg, tag); }
                        for (int r = lane; r < a.pt.world; r += 32) {
                            if (r == a.pt.rank) continue;
                            uint2* dst = reinterpret_cast<uint2*>(a.pt.peers[r]) + (tag & 1u) * ll_words + ((size_t)row * 2) * a.C + it.c;
                            st_ll(dst, mean, tag);
                            st_ll(dst + a.C, sg, tag);
                        }
                    } else if (lane == 0) {
                        st_ll_gpu(w, mean, tag);
                        st_ll_gpu(w + a.C, sg, tag);
                    }
                    if (lane == 0) {
                        a.mu[(int64_t)row * a.ld + it.c] = mean;
                        a.sig[(int64_t)row * a.ld + it.c] = sg;
                    }
                }
                float mu_p = mean, sg_p = sg;
                if (compute_std) {
                    // ---- first forward (maxstyle.py:165-168): the whole channel; lane l keeps rows l, l+32, ... in registers ----
                    float rm[kClStdRows], rs[kClStdRows];
#pragma unroll
                    for (int i = 0; i < kClStdRows; ++i) {
                        const int r = lane + 32 * i;
                        rm[i] = 0.f; rs[i] = 0.f;
                        if (r < NG) {
                            if (r == row) { rm[i] = mean; rs[i] = sg; }
                            else {
                                const bool remote = r < lo || r >= hi;
                                const uint2* w = ll_mine + ((size_t)r * 2) * a.C + it.c;
                                cl_poll_pair(w, w + a.C, tag, remote, a.error, rm[i], rs[i]);
                                if (remote && it.p == 0) { a.mu[(int64_t)r * a.ld + it.c] = rm[i]; a.sig[(int64_t)r * a.ld + it.c] = rs[i]; }
                            }
                        }
                    }
                    float s_sig = 0.f, s_mu = 0.f;
#pragma unroll
                    for (int i = 0; i < kClStdRows; ++i) { s_sig += rs[i]; s_mu += rm[i]; }      // rows >= NG hold zeros
                    s_sig = warp_sum(s_sig);
                    s_mu = warp_sum(s_mu);
                    const float mean_sig = s_sig / (float)NG, mean_mu = s_mu / (float)NG;
                    float q_sig = 0.f, q_mu = 0.f;
#pragma unroll
                    for (int i = 0; i < kClStdRows; ++i) {
                        if (lane + 32 * i < NG) {
                            const float ds = rs[i] - mean_sig, dm = rm[i] - mean_mu;
                            q_sig = fmaf(ds, ds, q_sig);
                            q_mu = fmaf(dm, dm, q_mu);
                        }
                    }
                    q_sig = warp_sum(q_sig);
                    q_mu = warp_sum(q_mu);
                    gs = sqrtf(q_sig / (float)(NG - 1));
                    bs = sqrtf(q_mu / (float)(NG - 1));
                    if (lane == 0 && it.n == 0 && it.p == 0 && a.gamma_std != nullptr) { a.gamma_std[it.c] = gs; a.beta_std[it.c] = bs; }
                }
                if (lane == 0) {
                    if (prow != row) {
                        const bool remote = prow < lo || prow >= hi;
                        const uint2* w = ll_mine + ((size_t)prow * 2) * a.C + it.c;
                        cl_poll_pair(w, w + a.C, tag, remote, a.error, mu_p, sg_p);
                        if (remote && it.p == 0) {               // the backward reads the partner's row from the local table
                            a.mu[(int64_t)prow * a.ld + it.c] = mu_p;
                            a.sig[(int64_t)prow * a.ld + it.c] = sg_p;
                        }
                    }
                    float sc, shf;
                    style_coeffs(sg, mean, sg_p, mu_p, mix, no_noise, lm, gn, bn, gs, bs, sc, shf, !(a.flags & 8));
                    if (it.p == 0) { a.scale[plane] = sc; a.shift[plane] = shf; }
                    for (int r = 0; r < CS; ++r) {
                        st_cluster_f4(map_to_cta(&sh.coef[s], (uint32_t)r), mean, sc, shf, 0.f);
                        mbar_arrive_remote(map_to_cta(&sh.coef_bar[s], (uint32_t)r));
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp < kClFirstApplyWarp) {
        // =============================== moments warps ===============================
        const int mt = t - 32 * kClFirstMomentWarp, mw = warp - kClFirstMomentWarp;
        for (int j = 0; j < my_items; ++j) {
            const int s = j % S;
            const uint32_t par = (uint32_t)((j / S) & 1);
            const ClItem it = cl_item(a, sh, g + (long long)j * G);
            int my_off, my_bytes;
            cl_range(a, it.p, (int)q, my_off, my_bytes);
            const float K = to_f32<T>(__ldg(x + ((int64_t)it.n * a.C + it.c) * a.M));      // the plane's first element: the same shift everywhere
            const char* buf = stage_buf + (size_t)s * a.part_stride;
            Moments acc{0.f, 0.f, 0.f};
            for (int ch = 0; ch < NCH; ++ch) {
                mbar_wait_t(&sh.full[s][ch], par, a.error);
                const int off = ch * a.chunk_bytes;
                const int nv = max(0, min(a.chunk_bytes, my_bytes - off)) >> 4;
                cl_moments_chunk<T>(buf + off, nv, mt, K, acc);
            }
            acc = warp_merge(acc);
            named_sync(1, kClMT);                                    // red_* may still be read from the previous item
            if (lane == 0) { sh.red_n[mw] = acc.n; sh.red_mean[mw] = acc.mean; sh.red_m2[mw] = acc.m2; }
            named_sync(1, kClMT);
            if (mt == 0) {
                Moments tot{0.f, 0.f, 0.f};
#pragma unroll
                for (int w = 0; w < kClMomentWarps; ++w) tot = merge(tot, Moments{sh.red_n[w], sh.red_mean[w], sh.red_m2[w]});
                st_cluster_f4(map_to_cta(&sh.part_in[s][q], 0u), tot.n, tot.mean, tot.m2, K);
                mbar_arrive_remote(map_to_cta(&sh.stats_bar[s], 0u));
            }
        }
    } else {
        // =============================== apply warps ===============================
        const int at = t - 32 * kClFirstApplyWarp;
        const uint64_t pol_out = make_policy(a.io_policy);
        constexpr int VE = ResVec<T>::kElems;
        for (int j = 0; j < my_items; ++j) {
            const int s = j % S;
            const uint32_t par = (uint32_t)((j / S) & 1);
            const ClItem it = cl_item(a, sh, g + (long long)j * G);
            int my_off, my_bytes;
            cl_range(a, it.p, (int)q, my_off, my_bytes);
            mbar_wait_cl(&sh.coef_bar[s], par, a.error);
            const float4 cf = sh.coef[s];
            const char* buf = stage_buf + (size_t)s * a.part_stride;
            T* dst = y + ((int64_t)it.n * a.C + it.c) * a.M + (size_t)(my_off >> 4) * VE;
            for (int ch = 0; ch < NCH; ++ch) {
                const int off = ch * a.chunk_bytes;
                const int nv = max(0, min(a.chunk_bytes, my_bytes - off)) >> 4;
                mbar_wait_t(&sh.full[s][ch], par, a.error);           // long complete: orders the TMA writes before these reads
                cl_apply_chunk<T>(buf + off, dst + (size_t)(off >> 4) * VE, nv, at, cf.x, cf.y, cf.z, pol_out);
                __syncwarp();
                if (lane == 0) mbar_arrive(&sh.empty[s][ch]);
            }
        }
    }
    // ---- nobody leaves while a peer CTA may still write into its shared memory; the last CTA out closes the launch ----
    __syncthreads();
    cluster_sync_all();
    if (t == 0) {
        __threadfence();
        if (atomicAdd(a.done, 1u) == gridDim.x - 1u) {
            *a.done = 0u;
            *a.epoch = tag;
            __threadfence();
        }
    }
}

}  // namespace ms
