// Host-side work decomposition shared by the C API and the workspace query.
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace ms {

constexpr int kBlocksPerSM = 4;                 // 4 x 256 threads resident per SM for the streaming kernels
constexpr int kVPT = 4;                         // vector accesses a thread keeps in flight per tensor
constexpr int64_t kWarpPlaneVecs = 512;         // planes up to this many vectors are handled by one warp
constexpr int kMaxSplits = 256;

struct Plan {
    int vec;          // elements per vector access: 16 B / sizeof(T), or 1 (scalar fallback)
    int group;        // threads cooperating on one item: 32 (small planes) or 256
    int splits;       // items per plane
    int64_t nvec;     // vectors per plane
    int64_t chunk;    // vectors per item (last item of a plane may be shorter)
    int64_t planes;   // N*C
    int64_t items;    // planes*splits
};

inline int elem_size(int dtype) { return dtype == 0 ? 4 : 2; }

// `align`: the largest power of two (<= 32) dividing every tensor base address involved.
// A plane is vector-accessible when the base is aligned and its byte size is a multiple of the
// vector width (then every plane start is aligned too).
inline Plan make_plan(int N, int C, int64_t M, int dtype, int align) {
    Plan p;
    const int es = elem_size(dtype);
    if (align >= 32 && (M * es) % 32 == 0) p.vec = 32 / es;
    else if (align >= 16 && (M * es) % 16 == 0) p.vec = 16 / es;
    else p.vec = 1;
    p.nvec = M / p.vec;
    p.planes = (int64_t)N * C;
    p.group = p.nvec <= kWarpPlaneVecs ? 32 : 256;
    // an item is one batch of group*kVPT vectors (32 KB with 256-bit vectors and a CTA group);
    // planes up to two batches stay whole.
    const int64_t batch = (int64_t)p.group * kVPT;
    int64_t chunk = p.nvec;
    if (p.group == 256 && p.nvec > 2 * batch) chunk = batch;
    int64_t s = (p.nvec + chunk - 1) / chunk;
    if (s > kMaxSplits) { chunk = ((p.nvec + kMaxSplits - 1) / kMaxSplits + batch - 1) / batch * batch; s = (p.nvec + chunk - 1) / chunk; }
    p.splits = (int)s;
    p.chunk = chunk;
    p.items = p.planes * s;
    return p;
}

// Workspace layout (bytes):  [plane counters: planes x int32][sample counters: N x int32]
//                            [done counter: 64 B][partials: planes x kMaxSplitsUsed x float4]
// Sized for the worst plan (scalar fallback has the most vectors per plane).
struct Workspace {
    size_t plane_counters, sample_counters, done_counter, partials, total;
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

inline Workspace workspace_layout(int N, int C, int64_t M, int dtype) {
    Workspace w;
    const int64_t planes = (int64_t)N * C;
    int smax = 1;
    for (int align = 1; align <= 32; align *= 2) {
        const int s = make_plan(N, C, M, dtype, align).splits;
        if (s > smax) smax = s;
    }
    size_t off = 0;
    w.plane_counters = off; off = align_up(off + planes * sizeof(int32_t), 256);
    w.sample_counters = off; off = align_up(off + (size_t)N * sizeof(int32_t), 256);
    w.done_counter = off; off += 256;
    w.partials = off; off = align_up(off + (size_t)planes * smax * 16, 256);
    w.total = off;
    return w;
}

}  // namespace ms
