// Host-side work decomposition shared by the C API and the workspace query.
//
// A tensor of `planes` = N*C planes of `nvec` vector accesses each is swept as ONE flat index
// space of planes*nvec vectors:
//   group == 256 (CTA mode):  CTA b owns vectors [b*per, (b+1)*per) -- a contiguous ~total/grid
//       slice that is cut into "pieces" at plane boundaries.  Work is balanced to one vector
//       whatever the plane size, a CTA reduces across its threads once per piece (not once per
//       32 KB), and a plane is shared by at most `slots` CTAs, each writing one partial.
//   group == 32 (warp mode, small planes):  warp w owns whole planes [w*per, (w+1)*per); no
//       block-wide synchronisation at all.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>

namespace ms {

constexpr int kBlocksPerSM = 4;                 // 4 x 256 threads resident per SM for the streaming kernels
constexpr int kThreadsPerBlock = 256;
constexpr int kMaxGrid = 2048;                  // workspace is sized for grids up to this (148 SMs x 4 = 592)
constexpr int64_t kWarpPlaneVecs = 512;         // planes up to this many vectors always go to single warps
constexpr int64_t kWarpPlaneVecsMany = 4096;    // ... and up to this many when there are >= 2 planes per warp

struct Plan {
    int vec;          // elements per vector access: 32 B or 16 B / sizeof(T), or 1 (scalar fallback)
    int group;        // threads cooperating on one piece: 32 (small planes) or 256
    int grid;         // CTAs to launch
    int slots;        // partial-result slots per plane (max number of CTAs sharing a plane)
    int64_t nvec;     // vectors per plane
    int64_t planes;   // N*C
    int64_t total;    // planes*nvec
    int64_t per;      // CTA mode: vectors per CTA;  warp mode: planes per warp
    int resident;     // CTA mode with oversub > 1: CTAs that fit on the device at once (grid slices are shared among them); else 0
};

inline int elem_size(int dtype) { return dtype == 0 ? 4 : 2; }

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// `align`: the largest power of two (<= 32) dividing every tensor base address involved.
// A plane is vector-accessible when the base is aligned and its byte size is a multiple of the
// vector width (then every plane start is aligned too).
inline Plan make_plan(int N, int C, int64_t M, int dtype, int align, int sms, int oversub = 1) {
    Plan p;
    p.resident = 0;
    const int es = elem_size(dtype);
    if (align >= 32 && (M * es) % 32 == 0) p.vec = 32 / es;
    else if (align >= 16 && (M * es) % 16 == 0) p.vec = 16 / es;
    else p.vec = 1;
    p.nvec = M / p.vec;
    p.planes = (int64_t)N * C;
    p.total = p.planes * p.nvec;
    int64_t max_ctas = (int64_t)sms * kBlocksPerSM;
    if (max_ctas > kMaxGrid) max_ctas = kMaxGrid;
    const int64_t max_warps = max_ctas * (kThreadsPerBlock / 32);
    const bool warp_mode = p.nvec <= kWarpPlaneVecs || (p.nvec <= kWarpPlaneVecsMany && p.planes >= 2 * max_warps);
    if (warp_mode) {
        p.group = 32;
        p.per = ceil_div(p.planes, max_warps);
        const int64_t warps = ceil_div(p.planes, p.per);
        p.grid = (int)ceil_div(warps, kThreadsPerBlock / 32);
        p.slots = 1;
    } else {
        p.group = 256;
        // oversub > 1 (backward): cut the tensor into `oversub` slices per resident CTA, never smaller than 16 rounds of the
        // CTA's loads (64 KB per tensor), never more than kMaxGrid (what the workspace's partial slots are sized for)
        if (oversub > 1) {
            p.resident = (int)max_ctas;
            int64_t v = max_ctas * oversub;
            if (v > kMaxGrid) v = kMaxGrid;
            const int64_t min_per = 16 * kThreadsPerBlock;
            if (p.total / v < min_per) v = p.total / min_per;
            if (v > max_ctas) max_ctas = v; else p.resident = 0;
        }
        int64_t per = ceil_div(p.total, max_ctas);
        per = ceil_div(per, kThreadsPerBlock) * kThreadsPerBlock;      // whole rows of threads
        p.per = per;
        p.grid = (int)ceil_div(p.total, per);
        p.slots = (int)(ceil_div(p.nvec, per) + 1);
    }
    return p;
}

// ---- NHWC (channels-last) feature maps: kernels_nhwc.cuh --------------------------------------------
// Memory is [N][H*W][C]: a pixel is a row of C contiguous elements = CV vector accesses, a sample is
// M*CV contiguous vectors.  The tensor is swept as ONE flat space of N*M*CV vectors cut into pieces at
// SAMPLE boundaries; `active` = floor(256/CV)*CV threads of a CTA stream it, so a thread meets the same
// VEC channels at every step ((v0 + t + i*active) mod CV is constant) and keeps their statistics in registers.
constexpr int kBlocksPerSMNhwc = 3;             // 3 x 256 threads per SM: up to 85 registers for the per-channel state

struct PlanNhwc {
    bool ok;
    int vec;          // elements per vector access (32 B or 16 B / sizeof(T), or 1)
    int cv;           // vectors per pixel = C / vec
    int active;       // streaming threads per CTA: floor(256/cv)*cv
    int grid;
    int slots;        // partial-result slots per (sample, channel) = max number of CTAs sharing a sample
    int64_t nvec;     // vectors per sample = M * cv
    int64_t total;    // N * nvec
    int64_t per;      // vectors per CTA (a multiple of `active`)
};

inline PlanNhwc make_plan_nhwc(int N, int C, int64_t M, int dtype, int align, int sms) {
    PlanNhwc p{};
    const int es = elem_size(dtype);
    // bf16 keeps to 16-byte vectors (8 channels): 16 channels of running statistics per thread do not fit the registers
    if (dtype == 0 && align >= 32 && ((int64_t)C * es) % 32 == 0) p.vec = 32 / es;
    else if (align >= 16 && ((int64_t)C * es) % 16 == 0) p.vec = 16 / es;
    else p.vec = 1;
    p.cv = C / p.vec;
    if (p.cv > kThreadsPerBlock) return p;                     // more than 256 vectors per pixel: no kernel
    p.active = kThreadsPerBlock / p.cv * p.cv;
    p.nvec = M * p.cv;
    p.total = (int64_t)N * p.nvec;
    int64_t max_ctas = (int64_t)sms * kBlocksPerSMNhwc;
    if (max_ctas > kMaxGrid) max_ctas = kMaxGrid;
    int64_t per = ceil_div(p.total, max_ctas);
    per = ceil_div(per, p.active) * p.active;
    p.per = per;
    p.grid = (int)ceil_div(p.total, per);
    p.slots = (int)(ceil_div(p.nvec, per) + 1);
    p.ok = true;
    return p;
}

// ---- fused forward (fused_fwd.cuh): ordered queue of statistics / apply items, channel-major --------
constexpr int kFusedStreamThreads = kThreadsPerBlock - 32;   // 7 streaming warps + 1 control warp per CTA
constexpr int kFusedMaxN = 1024;                             // rows a finalising warp keeps in shared memory
constexpr int64_t kFusedPieceBytes = 64 * 1024;              // target size of one item
constexpr int64_t kFusedWindowBytes = 32ll << 20;            // x kept in L2 between statistics and apply
constexpr int64_t kFusedProfitPlaneBytes = 64 * 1024;        // below this / above the next the two-pass path measured faster
constexpr int64_t kFusedProfitChannelBytes = 8ll << 20;      // (profiles/r01_sweep.jsonl: 25 KB planes, 12.8 MB channels)
constexpr int64_t kFusedMaxChannelBytes = 16ll << 20;        // the window must hold >= 2 channels: with one, every apply item
                                                             // waits for the finaliser of the channel streamed just before it
                                                             // (measured 2.3x slower than the two-pass path, profiles/r01_sweep.jsonl)

struct FusedPlan {
    bool ok;
    int vec, vpt, nvec, pieces, piece_vecs, items_per_channel, window, chunk;
    int64_t total_items;
    bool profitable;   // expected to beat the two-pass path (maxstyle_fwd only takes it then, unless forced)
};

// vector loads a thread keeps in flight for one tensor (same rule as the streaming kernels)
inline int vpt_one_tensor(int vec) { const int v = 32 / vec; return v < 1 ? 1 : (v > 4 ? 4 : v); }

// Development knobs.  Every environment variable the library looks at is read ONCE, here, the first time any planner asks
// (function-local static: initialised under the C++11 guard, immutable afterwards) -- no other getenv in the library, no
// per-call reads, nothing that can change between two launches of a captured graph.  Unset or non-positive = the default.
inline int64_t env_or(const char* name, int64_t dflt, int64_t unit) {
    const char* e = getenv(name);
    if (!e || !*e) return dflt;
    const long v = atol(e);
    return v > 0 ? (int64_t)v * unit : dflt;
}
inline int64_t env_flag(const char* name, int64_t dflt) {         // 0 is a meaningful value here
    const char* e = getenv(name);
    return (!e || !*e) ? dflt : (int64_t)atol(e);
}

constexpr int64_t kPairPieceBytesDefault = 112 * 1024;
constexpr int kClusterMaxStagesDefault = 6;

struct Tunables {
    int64_t fused_piece_bytes, fused_window_bytes, fused_chunk, fused_keep_bytes;      // MAXSTYLE_FUSED_{PIECE_KB,WINDOW_MB,CHUNK,KEEP_MB}
    int64_t ring_piece_chunks, ring_stages, ring_prefer;                               // MAXSTYLE_RING_{PIECE_CHUNKS,STAGES}, MAXSTYLE_RING
    int64_t cluster_enabled, cluster_cs, cluster_stages, cluster_pieces;               // MAXSTYLE_CLUSTER, _CS, _STAGES, _PIECES
    int64_t pair_enabled, pair_minb, pair_order, pair_piece_bytes;                     // MAXSTYLE_PAIR, _MINB, _ORDER, _PIECE_KB
    int64_t bwd_oversub;                                                               // MAXSTYLE_BWD_OVERSUB: slices per resident CTA in the backward
    int64_t pair_stagger_ns;                                                           // MAXSTYLE_PAIR_STAGGER_NS (-1: off)
};

inline const Tunables& tunables() {
    static const Tunables t = [] {
        Tunables v{};
        v.fused_piece_bytes = env_or("MAXSTYLE_FUSED_PIECE_KB", kFusedPieceBytes, 1024);
        v.fused_window_bytes = env_or("MAXSTYLE_FUSED_WINDOW_MB", kFusedWindowBytes, 1 << 20);
        v.fused_chunk = env_or("MAXSTYLE_FUSED_CHUNK", 1, 1);
        v.fused_keep_bytes = env_or("MAXSTYLE_FUSED_KEEP_MB", 0, 1 << 20);
        v.ring_piece_chunks = env_or("MAXSTYLE_RING_PIECE_CHUNKS", 8, 1);
        v.ring_stages = env_or("MAXSTYLE_RING_STAGES", 4, 1);
        v.ring_prefer = env_or("MAXSTYLE_RING", 0, 1);
        v.cluster_enabled = env_flag("MAXSTYLE_CLUSTER", 1);
        v.cluster_cs = env_or("MAXSTYLE_CLUSTER_CS", 0, 1);
        v.cluster_stages = env_or("MAXSTYLE_CLUSTER_STAGES", kClusterMaxStagesDefault, 1);
        v.cluster_pieces = env_or("MAXSTYLE_CLUSTER_PIECES", 0, 1);
        v.pair_enabled = env_flag("MAXSTYLE_PAIR", 1);
        v.pair_minb = env_or("MAXSTYLE_PAIR_MINB", 4, 1);          // CTAs per SM the register budget is set for (3: 85 registers, 4: 64)
        v.pair_order = env_or("MAXSTYLE_PAIR_ORDER", 0, 1);
        v.pair_piece_bytes = env_or("MAXSTYLE_PAIR_PIECE_KB", kPairPieceBytesDefault, 1024);
        v.pair_stagger_ns = env_flag("MAXSTYLE_PAIR_STAGGER_NS", 0);
        v.bwd_oversub = env_or("MAXSTYLE_BWD_OVERSUB", 4, 1);          // measured 1: 135.4, 2: 140.0, 3: 135.9, 4: 133.4, 6: 133.0 us (config 1)
        return v;
    }();
    return t;
}

inline FusedPlan make_fused_plan(int N, int C, int64_t M, int dtype, int align) {
    const int64_t piece_bytes = tunables().fused_piece_bytes, window_bytes = tunables().fused_window_bytes, chunk = tunables().fused_chunk;
    FusedPlan f{};
    const int es = elem_size(dtype);
    if (align >= 32 && (M * es) % 32 == 0) f.vec = 32 / es;
    else if (align >= 16 && (M * es) % 16 == 0) f.vec = 16 / es;
    else return f;                                            // scalar planes: two-pass path
    const int64_t nvec = M / f.vec;
    const int64_t channel_bytes = (int64_t)N * M * es;
    if (nvec < 512 || nvec > 0x7fffffff || N < 2 || N > kFusedMaxN || channel_bytes > kFusedMaxChannelBytes) return f;
    f.nvec = (int)nvec;
    f.vpt = vpt_one_tensor(f.vec);
    const int64_t step = (int64_t)kFusedStreamThreads * f.vpt;
    int64_t piece = piece_bytes / ((int64_t)f.vec * es) / step * step;
    if (piece < step) piece = step;
    f.piece_vecs = (int)piece;
    f.pieces = (int)ceil_div(nvec, piece);
    f.items_per_channel = N * f.pieces;
    int64_t d = window_bytes / channel_bytes;
    if (d < 1) d = 1;
    if (d > C) d = C;
    f.window = (int)d;
    f.total_items = 2ll * C * f.items_per_channel;
    f.chunk = (int)(chunk < 1 ? 1 : (chunk > 4 ? 4 : chunk));
    f.profitable = M * es >= kFusedProfitPlaneBytes && channel_bytes <= kFusedProfitChannelBytes;
    f.ok = true;
    return f;
}

// ---- streamed forward (ring_fwd.cuh): the window queue fed through a TMA ring -------------------------------
constexpr int kRingChunk = kFusedStreamThreads * 16 * 4;     // == kRingChunkBytes (14336)
constexpr int kRingCtrl = 12288;                             // == kRingCtrlBytes

struct RingPlan {
    bool ok;
    int stages, plane_bytes, piece_bytes, pieces, items_per_channel, window, smem;
    int64_t total_items;
    bool profitable;
};

inline RingPlan make_ring_plan(int N, int C, int64_t M, int dtype, int align) {
    const int64_t piece_chunks = tunables().ring_piece_chunks, stages = tunables().ring_stages, window_bytes = tunables().fused_window_bytes;
    RingPlan r{};
    const int64_t pb = M * elem_size(dtype);
    const int64_t channel_bytes = (int64_t)N * pb;
    if (align < 16 || pb % 16 != 0 || pb < 8192 || pb > 0x7fffffff) return r;
    if (N < 2 || N > kFusedMaxN || channel_bytes > kFusedMaxChannelBytes) return r;
    r.stages = (int)(stages < 2 ? 2 : (stages > 8 ? 8 : stages));
    r.plane_bytes = (int)pb;
    r.piece_bytes = (int)(piece_chunks < 1 ? 1 : piece_chunks) * kRingChunk;
    r.pieces = (int)ceil_div(pb, r.piece_bytes);
    r.items_per_channel = N * r.pieces;
    int64_t d = window_bytes / channel_bytes;
    if (d < 1) d = 1;
    if (d > C) d = C;
    r.window = (int)d;
    r.total_items = 2ll * C * r.items_per_channel;
    r.smem = kRingCtrl + r.stages * kRingChunk;
    // Measured equal to the register-staged window kernel within 2 % (profiles/r01_ring_knobs.txt: 110.7 vs 113.3 us on
    // the config-1 shape with 8-chunk pieces): the window kernel stays the default, MAXSTYLE_RING=1 or
    // MAXSTYLE_SWEEP_FORCE_RING select this one.
    const int64_t prefer = tunables().ring_prefer;
    r.profitable = prefer > 0 && pb >= kFusedProfitPlaneBytes && channel_bytes <= kFusedProfitChannelBytes;
    r.ok = true;
    return r;
}

// ---- resident forward (resident_fwd.cuh): a plane stays in shared memory between statistics and apply ----
constexpr int kResidentCtrlBytes = 4096;                     // == kResCtrlBytes
constexpr int kResidentMaxChunks = 16;
constexpr int kResidentMaxSmem = 232448;                     // 227 KB: the most one CTA may own on sm_100
constexpr int kResidentMinPlaneBytes = 16 * 1024;            // smaller planes: the flag round trip per item dominates
constexpr int kResidentMaxN = 304;                           // == kResMaxN

struct ResidentPlan {
    bool ok;
    int threads;       // 512 (one CTA per SM: planes > ~108 KB) or 256
    int plane_bytes, chunk_bytes, chunks;
    int smem;          // dynamic shared memory per CTA
    bool preferred;    // expected to beat the L2-window / two-pass paths (maxstyle_fwd only takes it then, unless forced)
};

inline ResidentPlan make_resident_plan(int N, int C, int64_t M, int dtype, int align) {
    ResidentPlan r{};
    const int64_t pb = M * elem_size(dtype);
    if (align < 16 || pb % 16 != 0 || pb < kResidentMinPlaneBytes) return r;
    if (N < 2 || N > kResidentMaxN) return r;
    const int64_t smem = kResidentCtrlBytes + (pb + 127) / 128 * 128;
    if (smem > kResidentMaxSmem) return r;
    r.threads = 2 * (smem + 1024) > 228 * 1024 ? 512 : 256;  // does a second CTA fit on the SM?
    const int consumers = r.threads - 32;
    r.chunk_bytes = consumers * 16 * (r.threads >= 512 ? 2 : 4);
    r.chunks = (int)ceil_div(pb, r.chunk_bytes);
    if (r.chunks > kResidentMaxChunks) return r;
    r.plane_bytes = (int)pb;
    r.smem = (int)smem;
    // Measured (profiles/r01_fwd_paths.txt): with >= 2 CTAs per SM the kernel beats the other paths by 10-15 % on tensors
    // that do not fit L2; with one 196 KB plane per SM the load -> moments -> partner -> store chain of a plane is not
    // hidden by anything else on the SM and the L2-window kernel wins; L2-sized tensors gain nothing from residency.
    r.preferred = r.threads == 256 && (int64_t)N * C * pb >= (64ll << 20);
    r.ok = true;
    return r;
}

// ---- cluster-resident forward (cluster_fwd.cuh): a plane split over a thread-block cluster, S stages per CTA ----
constexpr int kClusterCtrlBytes = 16384;                     // == kClCtrlBytes
constexpr int kClusterMaxStages = kClusterMaxStagesDefault;                        // == kClMaxStages
constexpr int kClusterMaxChunks = 16;                        // == kClMaxChunks
constexpr int kClusterChunkUnit = 8192;                      // apply warps: 256 threads x 16 B x 2 in flight
constexpr int kClusterMinPlaneBytes = 16 * 1024;
constexpr int kClusterMaxN = 1024;                           // == kClMaxN

constexpr int kClusterMaxPieces = 32;                        // == kClMaxPieces
constexpr int kClusterMinPartBytes = 16 * 1024;              // a CTA's share of a piece is at least this (P > 1 or CS > 1)
constexpr int kClusterStdRows = 512;                         // first forward: rows a coordinator warp keeps in registers (32 x kClStdRows)

struct ClusterPlan {
    bool ok;
    int cluster, pieces, stages, plane_bytes, part_bytes, part_stride, chunk_bytes, chunks, smem;
};

// Geometry for planes cut into `pieces` pieces, each held by a cluster of `cs` CTAs (1, 2, 4 or 8); `max_stages` caps S.
inline ClusterPlan make_cluster_plan(int64_t M, int dtype, int align, int cs, int pieces, int max_stages = kClusterMaxStages) {
    ClusterPlan c{};
    const int64_t pb = M * elem_size(dtype);
    if (align < 16 || pb % 16 != 0 || pb < kClusterMinPlaneBytes || pb > 0x7fffffff) return c;
    if (pieces < 1 || pieces > kClusterMaxPieces) return c;
    const int64_t shares = (int64_t)cs * pieces;
    const int64_t part = ceil_div(ceil_div(pb, shares), 16) * 16;
    if (shares > 1 && part < kClusterMinPartBytes) return c;
    if ((int64_t)(pieces - 1) * cs * part >= pb) return c;    // the last piece holds something
    const int64_t stride = ceil_div(part, 128) * 128;
    int64_t stages = (kResidentMaxSmem - kClusterCtrlBytes) / stride;
    if (stages < 1) return c;
    if (stages > kClusterMaxStages) stages = kClusterMaxStages;
    if (stages > max_stages) stages = max_stages;
    const int64_t chunk = kClusterChunkUnit * ceil_div(part, (int64_t)kClusterChunkUnit * kClusterMaxChunks);
    c.cluster = cs;
    c.pieces = pieces;
    c.stages = (int)stages;
    c.plane_bytes = (int)pb;
    c.part_bytes = (int)part;
    c.part_stride = (int)stride;
    c.chunk_bytes = (int)chunk;
    c.chunks = (int)ceil_div(part, chunk);
    c.smem = (int)(kClusterCtrlBytes + stages * stride);
    c.ok = true;
    return c;
}

// most pieces make_cluster_plan() accepts for this plane size (sizes the piece table of the workspace)
inline int cluster_max_pieces(int64_t M, int dtype) {
    const int64_t pb = M * elem_size(dtype);
    int64_t p = pb / kClusterMinPartBytes;
    if (p < 1) p = 1;
    return (int)(p > kClusterMaxPieces ? kClusterMaxPieces : p);
}

// ---- paired forward (pair_fwd.cuh): a CTA owns a piece of a plane for both of its passes ----------------------
constexpr int kPairMaxPieces = 32;
constexpr int64_t kPairPieceBytes = kPairPieceBytesDefault;             // largest piece: 592 CTAs x 112 KB live in L2 between the passes
constexpr int64_t kPairMinPlaneBytes = 8 * 1024;

struct PairPlan {
    bool ok;
    int vec, vpt, nvec, pieces, piece_vecs;
};

inline PairPlan make_pair_plan(int64_t M, int dtype, int align, int force_pieces = 0, int64_t planes = 0) {
    const int64_t piece_bytes = tunables().pair_piece_bytes;
    PairPlan p{};
    const int es = elem_size(dtype);
    if (align >= 32 && (M * es) % 32 == 0) p.vec = 32 / es;
    else if (align >= 16 && (M * es) % 16 == 0) p.vec = 16 / es;
    else return p;
    const int64_t nvec = M / p.vec, pb = M * es;
    if (pb < kPairMinPlaneBytes || nvec > 0x7fffffff) return p;
    p.nvec = (int)nvec;
    p.vpt = vpt_one_tensor(p.vec);
    int64_t P = force_pieces;
    // A tensor that all but fits in L2 anyway needs no pieces: whole planes (fewer items, no sibling exchange) measured 10-17 %
    // faster on 47-75 MB tensors of 147-196 KB planes (profiles/r02_fwd_paths.txt: 32x16x192x192, 20x16x224x224).
    if (P <= 0 && planes >= 296 && planes * pb <= (96ll << 20) && pb <= (256ll << 10)) P = 1;      // and enough planes to fill the SMs
    if (P <= 0) {
        // A piece of `len` vectors costs, per pass, len / round full rounds (256 threads x vpt loads in flight each) plus one
        // trip per 256 vectors of its ragged end, and ~1.5 rounds of publish / partner gap; 592 pieces must stay in L2 between
        // their passes (measured: 16 % of the second reads miss at 59 MB alive, 68 % at 89 MB): the fewest pieces of at most
        // `piece_bytes` (~112 KB), then the cheapest of the next few counts.
        const int64_t round = (int64_t)kThreadsPerBlock * p.vpt;
        const int64_t first = ceil_div(pb, piece_bytes);
        double best = 0.0;
        for (int64_t cand = first; cand <= first + 8 && cand <= kPairMaxPieces; ++cand) {
            const int64_t pv = ceil_div(nvec, cand);
            const double trips = (double)(pv / round) + (double)ceil_div(pv % round, kThreadsPerBlock);
            const double cost = (double)ceil_div(nvec, pv) * (trips + 1.5);
            if (P <= 0 || cost < best) { P = cand; best = cost; }
        }
    }
    const int64_t pv = ceil_div(nvec, P);
    P = ceil_div(nvec, pv);                                   // no empty pieces
    p.pieces = (int)P;
    p.piece_vecs = (int)pv;
    p.ok = true;
    return p;
}

// Workspace layout (bytes):  [plane tickets: planes x u64][sample tickets: N x u64]
//                            [done counter: 256 B][partials: planes x slots_bound x float4]
//                            [fused forward: error flag + queue + done 256 B][arrived | ready: 2 x C x u32][item partials]
//                            [resident forward: plane ready flags, planes x u32]
//                            [cluster forward: {value, tag} words, planes x 2 x 8 B][piece words, planes x max pieces x 2 x 8 B]
//                            [paired forward: {tag, pieces published} per plane, planes x u64]
//                            (their launch counter: u32 @24 of the error block)
// slots_bound covers every plan make_plan() can produce for this shape:
//   slots = ceil(nvec/per) + 1  with  per >= total/kMaxGrid  =>  slots <= kMaxGrid/planes + 2.
// NHWC (layout 1): the shared unit is the sample, each CTA publishes C partials for it (make_plan_nhwc).
struct Workspace {
    size_t plane_tickets, sample_tickets, done_counter, partials, res_error, res_flags, res_partials, plane_ready, cl_words, cl_pieces, pair_count, total;
    int slots_bound, max_pieces;
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

inline Workspace workspace_layout(int N, int C, int64_t M, int dtype, int layout = 0) {
    Workspace w;
    const int64_t planes = (int64_t)N * C;
    w.slots_bound = (int)(kMaxGrid / planes + 2);
    size_t off = 0;
    w.plane_tickets = off; off = align_up(off + planes * sizeof(uint64_t), 256);
    w.sample_tickets = off; off = align_up(off + (size_t)N * sizeof(uint64_t), 256);
    w.done_counter = off; off += 256;
    size_t partial_bytes = (size_t)planes * w.slots_bound * 16;
    if (layout == 1) {                                       // NHWC: a SAMPLE is shared by <= kMaxGrid/N + 2 CTAs, C partials each
        const size_t nhwc = (size_t)planes * (size_t)(kMaxGrid / N + 2) * 16;
        if (nhwc > partial_bytes) partial_bytes = nhwc;
    }
    w.partials = off; off = align_up(off + partial_bytes, 256);
    int64_t items = 0;                                       // the plan depends on the pointers' alignment: take the larger
    for (int align = 16; align <= 32; align *= 2) {
        const FusedPlan fp = make_fused_plan(N, C, M, dtype, align);
        if (fp.ok && (int64_t)C * fp.items_per_channel > items) items = (int64_t)C * fp.items_per_channel;
    }
    {
        const RingPlan rp = make_ring_plan(N, C, M, dtype, 16);
        if (rp.ok && (int64_t)C * rp.items_per_channel > items) items = (int64_t)C * rp.items_per_channel;
    }
    w.res_error = off; off += 256;                           // int error @0, u64 queue @8, u32 done @16, u32 launch counter @24
    w.res_flags = off; off = align_up(off + (size_t)2 * C * sizeof(uint32_t), 256);
    w.res_partials = off; off = align_up(off + (size_t)items * 16, 256);
    w.plane_ready = off; off = align_up(off + (size_t)planes * sizeof(uint32_t), 256);
    w.cl_words = off; off = align_up(off + (size_t)planes * 2 * 8, 256);
    {
        int mp = cluster_max_pieces(M, dtype);               // the piece table serves the cluster and the paired forward
        if (mp < kPairMaxPieces && M * elem_size(dtype) >= kPairMinPlaneBytes) {
            const int64_t want = ceil_div(M * elem_size(dtype), 8 * 1024);
            mp = (int)(want > kPairMaxPieces ? kPairMaxPieces : (want > mp ? want : mp));
        }
        w.cl_pieces = off; off = align_up(off + (size_t)planes * mp * 2 * 8, 256);
        w.max_pieces = mp;
    }
    w.pair_count = off; off = align_up(off + (size_t)planes * 8, 256);
    w.total = off;
    return w;
}

}  // namespace ms
