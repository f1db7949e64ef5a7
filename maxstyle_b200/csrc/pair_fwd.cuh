// Paired forward: the whole MaxStyle forward (maxstyle.py:157-185) in ONE persistent kernel that reads x from HBM once.
//
// In the steady state of the layer (gamma_std / beta_std cached, maxstyle.py:165-168) a plane depends on exactly ONE other
// plane: its mixing partner (perm[n], same channel).  So the unit of work is a PIECE of a plane (<= ~100 KB) that one CTA
// owns for both of its uses, back to back:
//     pass 1   stream the piece with 256-bit loads, shifted moments (the arithmetic of stats_nchw_kernel)       [HBM -> L2 -> SM]
//     publish  the piece's (mean, M2) as two 8-byte {value, tag} words; fetch the plane's other pieces and merge them in piece
//              order (every owner of a piece computes the same plane statistics); the owner of piece 0 publishes (mu, sig):
//              table words, the fp32 tables of the backward and -- when the batch is sharded over GPUs -- every peer's inbox
//     partner  fetch (mu, sig) of the partner plane (first forward: of the whole channel, and take the batch std); style_coeffs
//     pass 2   stream the SAME piece again -- it was read microseconds ago and is still in L2 -- and write
//              y = (x - mu) * A/sig + B                                                                           [L2 -> SM -> HBM]
// 4 CTAs x 256 threads per SM run this loop; while one CTA sits in its publish / partner gap (2-3 us of L2 round trips, warp 0
// only) the other three stream.  The k-th CTA of an SM starts k * 3 us late: started together, all CTAs read for ~9 us and then
// all sit in pass 2 (L2 hits in, L2 write-allocates out) with DRAM idle for ~10 us, and the rhythm persists (DRAM time series in
// profiles/r02_pm_series.txt).  The pieces live in L2 only between their two passes (592 CTAs x <= 100 KB, about what the L2 holds
// of such a stream: measured 15-24 % of the second reads miss at 59 MB, 68 % at 89 MB); the last batch of pass 1 does not even
// leave the registers.  Compared with fused_fwd.cuh (ordered statistics / apply queue with a 32 MB window): no control warp, no
// named-barrier hand-off per item, no channel-wide finaliser on the steady-state path.
// Items are taken with an atomic ticket, in order (channel-major; the pieces of a plane adjacent).  An item publishes before
// it waits, and waits only on publishes: its plane's pieces, and the plane words of its partner plane, which that plane's
// piece-0 owner (on this rank or another) publishes after waiting for nothing but publishes.  In natural sample order all of that
// lies within W = N*P positions (one channel); when a channel does not fit the grid the samples are taken in cycle order of the
// GLOBAL perm, the same walk on every rank, so the partner is the NEXT plane of one order shared by all ranks (or an earlier
// one, for the sample that closes a cycle): W = 2P.  The ticket of the next item is taken only after the wait, so a CTA never
// parks a ticket behind a wait; with more CTAs than W the lowest unpublished plane of the whole job can always be taken by a
// free CTA of its rank (host-side condition grid > W; the first forward awaits whole channels and needs N*P < grid).
// Variants measured and dropped (profiles/r02_fwd_experiments.txt): two items open per CTA, a control warp
// resolving item k while seven warps stream item k+1, partner statistics from piece words, bulk L2 prefetch, a read window.
#pragma once
#include "common.cuh"
#include "kernels_nchw.cuh"
#include "tables.cuh"
#include "fused_fwd.cuh"

namespace ms {

constexpr int kPairMaxN = 1024;          // rows of the order table (samples of this rank)
constexpr int kPairMaxNG = 2048;         // samples of the global batch the cycle walk can hold
constexpr int kPairStdRows = 16;         // first forward: rows of the channel a lane keeps in registers (n_global <= 512)
constexpr long long kPairSpinLocal = 4000000000LL;      // ~2 s
constexpr long long kPairSpinPeer = 40000000000LL;      // ~20 s: another rank may be late

struct PairArgs {
    int N, C;
    int64_t M;
    int nvec;                  // vectors per plane
    int pieces;                // P
    int piece_vecs;            // vectors per piece (the last piece of a plane may be shorter)
    int use_order;             // samples visited in cycle order of perm
    int stagger_cycles, slot_div;   // CTA b starts (b / slot_div) * stagger_cycles late, so the CTAs of an SM run out of phase
    int64_t total_items;       // N * C * P
    int flags;
    float eps;
    int pol_first, pol_second, pol_out;
    int pre_op;                // activation applied to x as it is loaded (common.cuh: kPre*), x = act(z)
    float pre_param;
    unsigned int *ymin, *ymax; // optional [N*C] ordered-integer min / max of y per plane (atomicMin / atomicMax)
    float *mu, *sig;           // [n_global, ld]
    float *scale, *shift;      // [N, C]
    int n_global, row_offset, ld;
    const int64_t* perm;
    const float *lmda, *gamma_noise, *beta_noise;
    float *gamma_std, *beta_std;
    uint2* ll;                 // [n_global][2][C] {value, tag} words (single GPU: workspace; multi GPU: the own inbox is used instead)
    uint2* piece_ll;           // [N*C][P][2]
    unsigned int* wepoch;      // launches on this workspace: tags the words that live in it
    unsigned int* epoch;       // multi GPU: the exchange epoch shared with the peers: tags the plane words in the inboxes
    unsigned long long* queue;
    unsigned int* done;
    int* error;
    PeerTables pt;
};

__device__ __forceinline__ void pair_fail(int* error) { wait_timed_out(error); }
__device__ __forceinline__ void st_ll_dev(void* p, float v, unsigned int tag) {
    asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(v)), "r"(tag) : "memory");
}
// Poll a pair of {value, tag} words until both carry `tag`.
__device__ __forceinline__ void pair_poll(const uint2* w0, const uint2* w1, unsigned int tag, bool remote, int* error, float& v0, float& v1) {
    if (ld_ll(w0, tag, v0) & ld_ll(w1, tag, v1)) return;
    const long long t0 = clock64();
    const long long limit = remote ? kPairSpinPeer : kPairSpinLocal;
    while (!(ld_ll(w0, tag, v0) & ld_ll(w1, tag, v1))) {
        __nanosleep(32);
        if (clock64() - t0 > limit) pair_fail(error);
    }
}

struct PairShared {
    Scratch scratch;
    long long next_id;
    unsigned int tag, xtag;                  // this launch's tags (kept here, not in registers, across the streaming loops)
    float4 coef;                             // (mu, scale, shift) of the current item
    unsigned short order[kPairMaxN];         // this rank's samples in cycle order of the GLOBAL perm
    unsigned short next_g[kPairMaxNG];       // perm, staged for the walk
    unsigned int seen[kPairMaxNG / 32];
};

// item id -> (channel, sample, piece); recomputed where needed instead of being kept in registers across the streaming loops
struct PairItem { int c, n, p; };
MS_HD PairItem pair_item(const PairArgs& a, const unsigned short* order, long long id) {
    PairItem it;
    const int per_channel = a.N * a.pieces;
    it.c = (int)(id / per_channel);
    const int r0 = (int)(id - (long long)it.c * per_channel);
    const int k = r0 / a.pieces;
    it.p = r0 - k * a.pieces;
    it.n = a.use_order ? (int)order[k] : k;
    return it;
}

// This rank's samples [lo, lo + N) in the order the cycles of the GLOBAL permutation meet them (next_g = perm as 16-bit entries,
// seen = a zeroed bitmap of NG bits).  A sample's partner perm[g] is the next sample of the walk, or -- for the sample that closes
// a cycle, and for fixed points -- one met earlier; every rank walks the same cycles, so the positions agree across ranks.
// Host-checked by tests/host/queue_check.cu.
MS_HD int pair_cycle_order(const unsigned short* next_g, unsigned int* seen, int NG, int lo, int N, unsigned short* order) {
    int k = 0;
    for (int s0 = 0; s0 < NG; ++s0) {
        int g = s0;
        while (!((seen[g >> 5] >> (g & 31)) & 1u)) {
            seen[g >> 5] |= 1u << (g & 31);
            if (g >= lo && g < lo + N) order[k++] = (unsigned short)(g - lo);
            g = next_g[g];
        }
    }
    return k;
}

// Everything between the two passes of a piece, run by warp 0 and kept out of line (its register needs -- the rows of a channel
// on the first forward -- must not spill into the streaming loops).  `tag` numbers the launches on this workspace (piece
// words), `xtag` the exchanges on the plane-word table (single GPU: the same number; multi GPU: the peers' shared epoch).
template <int VEC>
__device__ __noinline__ void pair_resolve(const PairArgs& a, Moments m, float K, int c, int n, int p, unsigned int tag, unsigned int xtag,
                                          float4* coef) {
    const int lane = threadIdx.x & 31;
    const int P = a.pieces;
    const bool multi = a.pt.world > 1;
    const bool mix = a.flags & 1, no_noise = a.flags & 2, compute_std = a.flags & 4;
    const float inv_m1 = 1.0f / (float)(a.M - 1);
    const int lo = a.row_offset, hi = a.row_offset + a.N, NG = a.n_global;
    uint2* ll = a.ll;
    size_t ll_words = 0;
    if (multi) {
        ll_words = (size_t)NG * 2 * a.C;
        ll = reinterpret_cast<uint2*>(a.pt.peers[a.pt.rank]) + (xtag & 1u) * ll_words;
    }
    const int64_t plane = (int64_t)n * a.C + c;
    const int row = lo + n;
    // everything that does not depend on other CTAs is requested before the first wait
    const int prow = mix ? (int)a.perm[row] : row;
    const float lm = mix ? a.lmda[n] : 0.f;
    float gn = 0.f, bn = 0.f, gs = 0.f, bs = 0.f;
    if (!no_noise) {
        gn = a.gamma_noise[plane]; bn = a.beta_noise[plane];
        if (!compute_std) { gs = a.gamma_std[c]; bs = a.beta_std[c]; }
    }
    if (P > 1) {
        uint2* mine = a.piece_ll + ((size_t)plane * P + p) * 2;
        if (lane == 0) { st_ll_dev(mine, m.mean, tag); st_ll_dev(mine + 1, m.m2, tag); }
        float pm = m.mean, pq = m.m2;
        if (lane < P && lane != p) {
            const uint2* w = a.piece_ll + ((size_t)plane * P + lane) * 2;
            pair_poll(w, w + 1, tag, false, a.error, pm, pq);
        }
        __syncwarp();
        Moments tot{0.f, 0.f, 0.f};
        for (int l = 0; l < P; ++l) {                                // piece order: the same sum in every owner
            const float lmean = __shfl_sync(0xffffffffu, pm, l), lm2 = __shfl_sync(0xffffffffu, pq, l);
            const int cnt = (min(a.nvec, (l + 1) * a.piece_vecs) - l * a.piece_vecs) * VEC;
            tot = merge(tot, Moments{(float)cnt, lmean, lm2});
        }
        m = tot;
    }
    const float mean = K + m.mean;
    const float sg = sqrtf(m.m2 * inv_m1 + a.eps);
    if (p == 0) {
        uint2* w = ll + ((size_t)row * 2) * a.C + c;
        if (multi) {
            if (lane == 0) { st_ll(w, mean, xtag); st_ll(w + a.C, sg, xtag); }
            for (int r = lane; r < a.pt.world; r += 32) {
                if (r == a.pt.rank) continue;
                uint2* dst = reinterpret_cast<uint2*>(a.pt.peers[r]) + (xtag & 1u) * ll_words + ((size_t)row * 2) * a.C + c;
                st_ll(dst, mean, xtag);
                st_ll(dst + a.C, sg, xtag);
            }
        } else if (lane == 0) {
            st_ll_dev(w, mean, xtag);
            st_ll_dev(w + a.C, sg, xtag);
        }
        if (lane == 0) {
            a.mu[(int64_t)row * a.ld + c] = mean;
            a.sig[(int64_t)row * a.ld + c] = sg;
        }
    }
    if (compute_std) {
        // first forward (maxstyle.py:165-168): the whole channel; lane l keeps rows l, l+32, ... in registers
        float rm[kPairStdRows], rs[kPairStdRows];
#pragma unroll
        for (int i = 0; i < kPairStdRows; ++i) {
            const int r = lane + 32 * i;
            rm[i] = 0.f; rs[i] = 0.f;
            if (r < NG) {
                if (r == row) { rm[i] = mean; rs[i] = sg; }
                else {
                    const bool remote = r < lo || r >= hi;
                    const uint2* w = ll + ((size_t)r * 2) * a.C + c;
                    pair_poll(w, w + a.C, xtag, remote, a.error, rm[i], rs[i]);
                    if (remote && p == 0) { a.mu[(int64_t)r * a.ld + c] = rm[i]; a.sig[(int64_t)r * a.ld + c] = rs[i]; }
                }
            }
        }
        float s_sig = 0.f, s_mu = 0.f;
#pragma unroll
        for (int i = 0; i < kPairStdRows; ++i) { s_sig += rs[i]; s_mu += rm[i]; }      // rows >= NG hold zeros
        s_sig = warp_sum(s_sig);
        s_mu = warp_sum(s_mu);
        const float mean_sig = s_sig / (float)NG, mean_mu = s_mu / (float)NG;
        float q_sig = 0.f, q_mu = 0.f;
#pragma unroll
        for (int i = 0; i < kPairStdRows; ++i) {
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
        if (lane == 0 && n == 0 && p == 0 && a.gamma_std != nullptr) { a.gamma_std[c] = gs; a.beta_std[c] = bs; }
    }
    if (lane == 0) {
        float mu_p = mean, sg_p = sg;
        if (prow != row) {
            const bool remote = prow < lo || prow >= hi;
            const uint2* w = ll + ((size_t)prow * 2) * a.C + c;
            pair_poll(w, w + a.C, xtag, remote, a.error, mu_p, sg_p);
            if (remote && p == 0) {                      // the backward reads the partner's row from the local table
                a.mu[(int64_t)prow * a.ld + c] = mu_p;
                a.sig[(int64_t)prow * a.ld + c] = sg_p;
            }
        }
        float sc, shf;
        style_coeffs(sg, mean, sg_p, mu_p, mix, no_noise, lm, gn, bn, gs, bs, sc, shf, !(a.flags & 8));
        if (p == 0) { a.scale[plane] = sc; a.shift[plane] = shf; }
        *coef = make_float4(mean, sc, shf, 0.f);
    }
}

// ---- the two streaming passes over a piece [v0, v1) of a plane (inlined: ptxas 12.9 crashes on these loops in a
// non-inlined function); G threads take part.  A pass is a chain of round trips (issue VPT loads per thread, wait, compute), and
// with pieces of ~3 batches every round trip saved is ~8 % of the item, so:
//   * the ragged end of the piece (fewer than G*VPT vectors) is loaded in the shadow of the first batch's loads, not after them;
//   * the LAST batch of pass 1 stays in registers (`held`) across the publish / partner gap and is the FIRST batch pass 2
//     writes -- pass 2 walks the piece backwards, so it re-reads one batch less from L2 and starts storing without a load.
template <typename T, int VEC, int VPT, int G>
__device__ __forceinline__ void pair_load_batch(const T* base, int first_vec, int t, uint64_t pol, float (&val)[VPT][VEC]) {
    const T* ptr = base + (int64_t)(first_vec + t) * VEC;
#pragma unroll
    for (int j = 0; j < VPT; ++j) Vec<T, VEC>::load(ptr + (int64_t)j * G * VEC, val[j], pol);
}

template <typename T, int VEC, int VPT, int G>
__device__ __forceinline__ Moments pair_pass1(const T* base, int v0, int v1, int pol_kind, int t, int pre_op, float pre_param, float& K_out,
                                              float (&held)[VPT][VEC]) {
    constexpr int kStep = G * VPT;
    const int full = (v1 - v0) / kStep, tail_lo = v0 + full * kStep;
    const uint64_t pol = make_policy(pol_kind);
    const float K = pre_apply(to_f32<T>(__ldg(base)), pre_op, pre_param);      // the plane's first element: the same shift in every piece
    Moments acc{0.f, 0.f, 0.f};
    if (full > 0) pair_load_batch<T, VEC, VPT, G>(base, v0, t, pol, held);
    for (int v = tail_lo + t; v < v1; v += G) {                  // the ragged end: one vector per thread and trip
        float val[VEC];
        Vec<T, VEC>::load(base + (int64_t)v * VEC, val, pol);
        pre_apply_vec(val, pre_op, pre_param);
        float s = 0.f;
#pragma unroll
        for (int e = 0; e < VEC; ++e) { val[e] -= K; s += val[e]; }
        Moments bm;
        bm.n = (float)VEC;
        bm.mean = s * (1.0f / (float)VEC);
        float qq = 0.f;
#pragma unroll
        for (int e = 0; e < VEC; ++e) { const float d = val[e] - bm.mean; qq = fmaf(d, d, qq); }
        bm.m2 = qq;
        acc = merge_fast(acc, bm);
    }
    for (int b = 0; b < full; ++b) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
            pre_apply_vec(held[j], pre_op, pre_param);           // held keeps act(z): pass 2 writes from it
#pragma unroll
            for (int e = 0; e < VEC; ++e) s += held[j][e] - K;
        }
        Moments bm;
        bm.n = (float)(VPT * VEC);
        bm.mean = s * (1.0f / (float)(VPT * VEC));
        float qq = 0.f;
#pragma unroll
        for (int j = 0; j < VPT; ++j)
#pragma unroll
            for (int e = 0; e < VEC; ++e) { const float d = (held[j][e] - K) - bm.mean; qq = fmaf(d, d, qq); }
        bm.m2 = qq;
        acc = merge_fast(acc, bm);
        if (b + 1 < full) pair_load_batch<T, VEC, VPT, G>(base, v0 + (b + 1) * kStep, t, pol, held);
    }
    K_out = K;
    return acc;
}

template <typename T, int VEC, int VPT, int G>
__device__ __forceinline__ void pair_pass2(const T* base, T* dst, int v0, int v1, float mu0, float sc, float shf, int pol_in_kind,
                                           int pol_out_kind, int t, int pre_op, float pre_param, unsigned int* ymin, unsigned int* ymax,
                                           float (&held)[VPT][VEC]) {
    constexpr int kStep = G * VPT;
    float lo_y = INFINITY, hi_y = -INFINITY;
    const int full = (v1 - v0) / kStep, tail_lo = v0 + full * kStep;
    const uint64_t pol_in = make_policy(pol_in_kind), pol_out = make_policy(pol_out_kind);
    for (int b = full - 1; b >= 0; --b) {                        // batch full-1 is still in registers from pass 1
        const int64_t o = (int64_t)(v0 + b * kStep + t) * VEC;
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
            if (b != full - 1) pre_apply_vec(held[j], pre_op, pre_param);
#pragma unroll
            for (int e = 0; e < VEC; ++e) held[j][e] = fmaf(held[j][e] - mu0, sc, shf);
            if (ymin != nullptr) {
#pragma unroll
                for (int e = 0; e < VEC; ++e) { lo_y = fminf(lo_y, held[j][e]); hi_y = fmaxf(hi_y, held[j][e]); }
            }
            Vec<T, VEC>::store(dst + o + (int64_t)j * G * VEC, held[j], pol_out);
        }
        if (b > 0) pair_load_batch<T, VEC, VPT, G>(base, v0 + (b - 1) * kStep, t, pol_in, held);
        if (b == full - 1) {
            // the ragged end, in the shadow of the loads just issued
            for (int v = tail_lo + t; v < v1; v += G) {
                float val[VEC];
                Vec<T, VEC>::load(base + (int64_t)v * VEC, val, pol_in);
                pre_apply_vec(val, pre_op, pre_param);
#pragma unroll
                for (int e = 0; e < VEC; ++e) { val[e] = fmaf(val[e] - mu0, sc, shf); lo_y = fminf(lo_y, val[e]); hi_y = fmaxf(hi_y, val[e]); }
                Vec<T, VEC>::store(dst + (int64_t)v * VEC, val, pol_out);
            }
        }
    }
    if (full == 0) {
        for (int v = tail_lo + t; v < v1; v += G) {
            float val[VEC];
            Vec<T, VEC>::load(base + (int64_t)v * VEC, val, pol_in);
            pre_apply_vec(val, pre_op, pre_param);
#pragma unroll
            for (int e = 0; e < VEC; ++e) { val[e] = fmaf(val[e] - mu0, sc, shf); lo_y = fminf(lo_y, val[e]); hi_y = fmaxf(hi_y, val[e]); }
            Vec<T, VEC>::store(dst + (int64_t)v * VEC, val, pol_out);
        }
    }
    if (ymin != nullptr) warp_minmax_publish(lo_y, hi_y, ymin, ymax);
}

template <typename T, int VEC, int VPT, int MINB>
__global__ void __launch_bounds__(kThreads, MINB)
fwd_pair_kernel(const T* __restrict__ x, T* __restrict__ y, const __grid_constant__ PairArgs a) {
    __shared__ PairShared sh;
    constexpr int G = kThreads;
    const int t = threadIdx.x;
    {
        const unsigned int wtag = *(volatile unsigned int*)a.wepoch + 1u;      // advanced by the last CTA out, after every CTA has read them
        const unsigned int xtag = a.pt.world > 1 ? *(volatile unsigned int*)a.epoch + 1u : wtag;
        const int lo = a.row_offset;
        if (a.use_order) {
            // Walk the cycles of the GLOBAL permutation and keep this rank's samples in the order they are met: a sample's partner
            // is then the next sample of the walk (on this rank or on another one -- every rank walks the same cycles), or, for the
            // sample that closes a cycle, one that came earlier.  Dependencies only point one step ahead in ONE order shared by
            // all ranks, so the lowest unpublished plane of the whole job can always be taken by a free CTA of its rank.
            const int NG = a.n_global;
            for (int i = t; i < NG; i += G) sh.next_g[i] = (unsigned short)a.perm[i];
            for (int i = t; i < kPairMaxNG / 32; i += G) sh.seen[i] = 0u;
            __syncthreads();
            if (t == 0) pair_cycle_order(sh.next_g, sh.seen, NG, lo, a.N, sh.order);
        }
        if (t == 0) {
            if (a.stagger_cycles > 0) {                     // the ticket is taken after the delay: nothing waits on a sleeping CTA
                const long long wait = (long long)(blockIdx.x / a.slot_div) * a.stagger_cycles, t0 = clock64();
                while (clock64() - t0 < wait) __nanosleep(200);
            }
            sh.next_id = (long long)atomicAdd(a.queue, 1ull); sh.tag = wtag; sh.xtag = xtag;
        }
        if (a.pt.world > 1 && xtag > 1u && t < 32) {
            // Flow control for the two-parity inboxes: this launch overwrites the words of launch xtag-2.  A peer that has published
            // anything in launch xtag-1 has finished launch xtag-2: wait for one word of launch xtag-1 from each peer (its first row,
            // last channel -- every exchange kernel publishes it) before the first push.
            const size_t ll_words = (size_t)a.n_global * 2 * a.C;
            const uint2* prev_inbox = reinterpret_cast<const uint2*>(a.pt.peers[a.pt.rank]) + ((xtag - 1u) & 1u) * ll_words;
            for (int r = t; r < a.pt.world; r += 32) {
                if (r == a.pt.rank) continue;
                const uint2* src = prev_inbox + ((size_t)(r * a.N) * 2) * a.C + (a.C - 1);
                const long long t0 = clock64();
                for (;;) {
                    unsigned int bits, got;
                    asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(bits), "=r"(got) : "l"(src) : "memory");
                    if ((int)(got - (xtag - 1u)) >= 0) break;
                    __nanosleep(100);
                    if (clock64() - t0 > kPairSpinPeer) pair_fail(a.error);
                }
            }
        }
        __syncthreads();
    }
    long long id = sh.next_id;
    while (id < a.total_items) {
        const PairItem it = pair_item(a, sh.order, id);
        const int64_t plane = (int64_t)it.n * a.C + it.c;
        const int v0 = it.p * a.piece_vecs, v1 = min(a.nvec, v0 + a.piece_vecs);
        float K;
        float held[VPT][VEC];                                          // the piece's last batch: loaded in pass 1, written in pass 2
        const Moments acc = pair_pass1<T, VEC, VPT, G>(x + plane * a.M, v0, v1, a.pol_first, t, a.pre_op, a.pre_param, K, held);
        const Moments m = group_merge<G>(acc, sh.scratch);             // valid in every thread
        if (t < 32) {
            pair_resolve<VEC>(a, m, K, it.c, it.n, it.p, sh.tag, sh.xtag, &sh.coef);
            // the wait is over: only now may this CTA hold the ticket of another item (its latency hides under pass 2)
            if (t == 0) sh.next_id = (long long)atomicAdd(a.queue, 1ull);
        }
        __syncthreads();
        const float4 cf = sh.coef;
        id = sh.next_id;                       // coef / next_id are rewritten only after the next item's block reduction (two barriers)
        pair_pass2<T, VEC, VPT, G>(x + plane * a.M, y + plane * a.M, v0, v1, cf.x, cf.y, cf.z, a.pol_second, a.pol_out, t, a.pre_op, a.pre_param,
                                   a.ymin ? a.ymin + plane : nullptr, a.ymax ? a.ymax + plane : nullptr, held);
    }
    if (t == 0) {
        __threadfence();
        if (atomicAdd(a.done, 1u) == gridDim.x - 1u) {
            *a.done = 0u;
            *a.queue = 0ull;
            *a.wepoch = sh.tag;
            if (a.pt.world > 1) *a.epoch = sh.xtag;
            __threadfence();
        }
    }
}

}  // namespace ms
