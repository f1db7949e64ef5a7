// Host-side check of the single-kernel forwards' work plans (plan.h) and of the ordered statistics/apply queue
// (fused_item, fused_fwd.cuh -- shared by the L2-window and the TMA-ring kernels).  Built and run by tests/test_plan_host.py.
//  * every (statistics | apply, channel, sample, piece) item appears exactly once in the queue;
//  * every apply item comes after ALL statistics items of its channel -- the property the kernels' deadlock-freedom
//    argument rests on (an item only ever waits for items with lower tickets);
//  * apply(c) is handed out `window` channels of statistics later, never earlier;
//  * pieces / chunks tile a plane exactly, stay within the shared-memory and mbarrier limits, and the workspace has
//    room for every item's partial moments.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../maxstyle_b200/csrc/fused_fwd.cuh"
#include "../../maxstyle_b200/csrc/pair_fwd.cuh"
#include "../../maxstyle_b200/csrc/plan.h"

using namespace ms;

static int failures = 0;
#define CHECK(cond, ...) do { if (!(cond)) { if (failures < 20) { printf("FAIL %s:%d: ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); } ++failures; } } while (0)

static void check_queue(int N, int C, int pieces, int items_per_channel, int window, int64_t total_items, const char* what) {
    FusedArgs a{};
    a.N = N; a.C = C; a.pieces = pieces; a.items_per_channel = items_per_channel; a.window = window; a.total_items = total_items;
    CHECK(total_items == 2ll * C * items_per_channel && items_per_channel == N * pieces && window >= 1 && window <= C, "%s: plan fields", what);
    std::vector<int64_t> pos_stats((size_t)C * items_per_channel, -1), pos_apply((size_t)C * items_per_channel, -1);
    for (int64_t id = 0; id < total_items; ++id) {
        const FusedItem it = fused_item(a, id);
        CHECK(it.c >= 0 && it.c < C && it.n >= 0 && it.n < N && it.k >= 0 && it.k < pieces && (it.apply == 0 || it.apply == 1),
              "%s: item %lld out of range (c=%d n=%d k=%d)", what, (long long)id, it.c, it.n, it.k);
        if (it.c < 0 || it.c >= C || it.n < 0 || it.n >= N || it.k < 0 || it.k >= pieces) continue;
        auto& slot = (it.apply ? pos_apply : pos_stats)[(size_t)it.c * items_per_channel + (size_t)it.n * pieces + it.k];
        CHECK(slot < 0, "%s: item visited twice", what);
        slot = id;
    }
    for (int c = 0; c < C; ++c) {
        int64_t last_stats = -1, first_apply = total_items;
        for (int i = 0; i < items_per_channel; ++i) {
            const int64_t s = pos_stats[(size_t)c * items_per_channel + i], p = pos_apply[(size_t)c * items_per_channel + i];
            CHECK(s >= 0 && p >= 0, "%s: channel %d item %d missing", what, c, i);
            if (s > last_stats) last_stats = s;
            if (p < first_apply) first_apply = p;
        }
        CHECK(last_stats < first_apply, "%s: channel %d: an apply item (%lld) precedes a statistics item (%lld)", what, c,
              (long long)first_apply, (long long)last_stats);
        // the window: before apply(c) starts, the statistics of channels c+1 .. min(c+window, C)-1 have been handed out too
        const int ahead = c + window < C ? c + window - 1 : C - 1;
        int64_t last_ahead = -1;
        for (int i = 0; i < items_per_channel; ++i) {
            const int64_t s = pos_stats[(size_t)ahead * items_per_channel + i];
            if (s > last_ahead) last_ahead = s;
        }
        CHECK(last_ahead < first_apply, "%s: channel %d: window %d not respected", what, c, window);
    }
}

// The paired forward's sample order (pair_cycle_order) and item map (pair_item):
//  * every rank's order is a permutation of its samples and the sub-sequence of ONE global walk;
//  * a sample's partner perm[g] sits at the next position of that walk, or at an earlier one (cycle closing / fixed point) --
//    what bounds an item's wait to W = 2P tickets ahead on any rank;
//  * item ids map onto (channel, sample, piece) one-to-one, channel-major, the pieces of a plane adjacent.
static void check_pair_order(int world, int N, int P, int C, unsigned seed) {
    const int NG = world * N;
    std::vector<unsigned short> perm((size_t)NG);
    for (int i = 0; i < NG; ++i) perm[i] = (unsigned short)i;
    unsigned s = seed * 2654435761u + 12345u;
    for (int i = NG - 1; i > 0; --i) { s = s * 1664525u + 1013904223u; const int j = (int)((s >> 8) % (unsigned)(i + 1)); std::swap(perm[i], perm[j]); }
    std::vector<unsigned int> seen((size_t)(NG + 31) / 32, 0u);
    std::vector<unsigned short> walk((size_t)NG);
    CHECK(pair_cycle_order(perm.data(), seen.data(), NG, 0, NG, walk.data()) == NG, "global walk length");
    std::vector<int> pos((size_t)NG, -1);
    for (int k = 0; k < NG; ++k) { CHECK(pos[walk[k]] == -1, "sample twice in the walk"); pos[walk[k]] = k; }
    for (int g = 0; g < NG; ++g) {
        const int p = perm[g];
        CHECK(pos[p] == pos[g] + 1 || pos[p] <= pos[g], "partner of %d at %d, own position %d (world %d N %d seed %u)", g, pos[p], pos[g], world, N, seed);
    }
    for (int r = 0; r < world; ++r) {
        std::fill(seen.begin(), seen.end(), 0u);
        std::vector<unsigned short> order((size_t)N);
        CHECK(pair_cycle_order(perm.data(), seen.data(), NG, r * N, N, order.data()) == N, "rank order length");
        int last = -1;
        std::vector<int> hit((size_t)N, 0);
        for (int k = 0; k < N; ++k) {
            CHECK(order[k] < N, "order entry out of range");
            if (order[k] >= N) continue;
            hit[order[k]]++;
            const int gp = pos[r * N + order[k]];
            CHECK(gp > last, "rank %d order is not a sub-sequence of the global walk", r);
            last = gp;
        }
        for (int n = 0; n < N; ++n) CHECK(hit[n] == 1, "rank %d sample %d appears %d times", r, n, hit[n]);
        // item map with this order
        PairArgs a{};
        a.N = N; a.C = C; a.pieces = P; a.use_order = 1; a.total_items = (int64_t)N * C * P;
        std::vector<int> cover((size_t)N * C * P, 0);
        int64_t prev_plane = -1; int prev_p = -1;
        for (int64_t id = 0; id < a.total_items; ++id) {
            const PairItem it = pair_item(a, order.data(), id);
            CHECK(it.c >= 0 && it.c < C && it.n >= 0 && it.n < N && it.p >= 0 && it.p < P, "pair_item out of range");
            if (!(it.c >= 0 && it.c < C && it.n >= 0 && it.n < N && it.p >= 0 && it.p < P)) continue;
            cover[((size_t)it.c * N + it.n) * P + it.p]++;
            const int64_t plane = (int64_t)it.c * N + it.n;
            if (it.p > 0) CHECK(plane == prev_plane && it.p == prev_p + 1, "pieces of a plane are not adjacent");
            prev_plane = plane; prev_p = it.p;
            CHECK(it.c == (int)(id / ((int64_t)N * P)), "items are not channel-major");
        }
        for (size_t i = 0; i < cover.size(); ++i) CHECK(cover[i] == 1, "item coverage %d", cover[i]);
    }
}

int main() {
    const int shapes[][4] = {{20, 64, 224, 224}, {20, 16, 96, 96}, {20, 1, 224, 224}, {3, 2, 160, 160}, {2, 3, 130, 130}, {32, 16, 192, 192},
                             {6, 8, 128, 128}, {40, 5, 72, 72}, {64, 64, 112, 112}, {64, 256, 56, 56}, {512, 16, 112, 112}, {2, 2, 512, 512},
                             {7, 3, 100, 100}, {160, 8, 160, 160}, {1024, 2, 48, 48}, {5, 4, 64, 66}};
    int cases = 0;
    for (auto& s : shapes)
        for (int dtype = 0; dtype < 2; ++dtype)
            for (int align : {32, 16}) {
                const int N = s[0], C = s[1];
                const int64_t M = (int64_t)s[2] * s[3];
                const int64_t pb = M * elem_size(dtype);
                const Workspace w = workspace_layout(N, C, M, dtype);
                const FusedPlan fp = make_fused_plan(N, C, M, dtype, align);
                if (fp.ok) {
                    check_queue(N, C, fp.pieces, fp.items_per_channel, fp.window, fp.total_items, "window");
                    const int64_t piece_bytes = (int64_t)fp.piece_vecs * fp.vec * elem_size(dtype);
                    CHECK((int64_t)fp.pieces * piece_bytes >= pb && (int64_t)(fp.pieces - 1) * piece_bytes < pb, "window: pieces do not tile the plane");
                    CHECK((int64_t)fp.nvec * fp.vec == M, "window: vectors do not tile the plane");
                    CHECK(w.plane_ready - w.res_partials >= (size_t)C * fp.items_per_channel * 16, "window: workspace too small for the item partials");
                    ++cases;
                }
                const RingPlan rp = make_ring_plan(N, C, M, dtype, align);
                if (rp.ok) {
                    check_queue(N, C, rp.pieces, rp.items_per_channel, rp.window, rp.total_items, "ring");
                    CHECK(rp.piece_bytes % kRingChunk == 0 && rp.plane_bytes == pb && pb % 16 == 0, "ring: piece / plane bytes");
                    CHECK((int64_t)rp.pieces * rp.piece_bytes >= pb && (int64_t)(rp.pieces - 1) * rp.piece_bytes < pb, "ring: pieces do not tile the plane");
                    CHECK(rp.stages >= 2 && rp.stages <= 8 && rp.smem == kRingCtrl + rp.stages * kRingChunk && rp.smem <= kResidentMaxSmem, "ring: stages / smem");
                    CHECK(w.plane_ready - w.res_partials >= (size_t)C * rp.items_per_channel * 16, "ring: workspace too small for the item partials");
                    ++cases;
                }
                const ResidentPlan sp = make_resident_plan(N, C, M, dtype, align);
                if (sp.ok) {
                    CHECK(sp.threads == 256 || sp.threads == 512, "resident: threads");
                    CHECK(sp.chunk_bytes == (sp.threads - 32) * 16 * (sp.threads == 512 ? 2 : 4), "resident: chunk bytes");
                    CHECK(sp.chunks >= 1 && sp.chunks <= kResidentMaxChunks && (int64_t)sp.chunks * sp.chunk_bytes >= pb &&
                          (int64_t)(sp.chunks - 1) * sp.chunk_bytes < pb, "resident: chunks do not tile the plane");
                    CHECK(sp.smem <= kResidentMaxSmem && sp.smem >= kResidentCtrlBytes + pb && sp.plane_bytes == pb && pb % 16 == 0, "resident: smem");
                    CHECK(sp.threads == 512 || 2 * (sp.smem + 1024) <= 228 * 1024, "resident: 256-thread variant must fit two CTAs per SM");
                    CHECK(w.total - w.plane_ready >= (size_t)N * C * 4, "resident: workspace too small for the plane flags");
                    ++cases;
                }
            }
    const int orders[][4] = {{1, 20, 2, 3}, {1, 2, 1, 2}, {1, 3, 16, 2}, {2, 8, 2, 4}, {2, 40, 16, 2}, {4, 20, 2, 3}, {8, 20, 2, 2}, {8, 256, 16, 1}, {3, 7, 3, 2},
                             {1, 1024, 2, 1}, {2, 1024, 1, 1}, {5, 33, 4, 2}};      // world, N per rank, pieces, channels
    for (auto& o : orders)
        for (unsigned seed = 1; seed <= 12; ++seed) { check_pair_order(o[0], o[1], o[2], o[3], seed); ++cases; }
    printf("%d cases, %d failures\n", cases, failures);
    return failures ? 1 : 0;
}
