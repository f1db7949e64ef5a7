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
    printf("%d cases, %d failures\n", cases, failures);
    return failures ? 1 : 0;
}
