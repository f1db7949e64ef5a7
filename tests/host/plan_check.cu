// Host-side exhaustive check of the flat-sweep work decomposition (plan.h + the iterators of
// kernels_nchw.cuh, which are __host__ __device__).  Built and run by tests/test_plan_host.py.
// For every (shape, SM count, direction): the pieces of all groups tile every plane exactly once;
// batches tile every piece exactly once; the CTAs sharing a plane are exactly plane_share() and fit
// the plan's slots and the workspace's slot bound; per-sample vector tickets add up to C*nvec.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../maxstyle_b200/csrc/kernels_nchw.cuh"
#include "../../maxstyle_b200/csrc/plan.h"

using namespace ms;

static int failures = 0;
#define CHECK(cond, ...) do { if (!(cond)) { if (failures < 20) { printf("FAIL %s:%d: ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); } ++failures; } } while (0)

// oversub > 1: the backward's plans -- more slices than resident CTAs, handed out by a ticket counter (slice index = partial slot)
template <int G, int VPT>
void check_case(int N, int C, int64_t M, int dtype, int align, int sms, int reverse, int oversub = 1) {
    const Plan p = make_plan(N, C, M, dtype, align, sms, oversub);
    if (p.group != G) return;
    const Workspace w = workspace_layout(N, C, M, dtype);
    CHECK(p.slots <= w.slots_bound, "slots %d > bound %d (N=%d C=%d M=%lld oversub=%d)", p.slots, w.slots_bound, N, C, (long long)M, oversub);
    CHECK(p.grid >= 1 && p.grid <= (oversub > 1 ? kMaxGrid : sms * kBlocksPerSM), "grid %d", p.grid);
    if (oversub > 1 && p.resident > 0) CHECK(p.resident <= sms * kBlocksPerSM && p.grid > p.resident, "resident %d grid %d", p.resident, p.grid);
    Sweep g{};
    g.M = M; g.nvec = p.nvec; g.planes = p.planes; g.total = p.total; g.per = p.per; g.slots = p.slots; g.reverse = reverse;
    std::vector<int> cover((size_t)p.total, 0);
    std::vector<int> touched((size_t)p.planes, 0);
    std::vector<unsigned long long> sample((size_t)N, 0);
    const int64_t groups = (int64_t)p.grid * (kThreads / G);
    for (int64_t gi = 0; gi < groups; ++gi) {
        PieceIter<G> it(g, gi);
        Piece pc;
        int64_t last_plane = -1;
        while (it.next(pc)) {
            CHECK(pc.plane >= 0 && pc.plane < p.planes && pc.v0 >= 0 && pc.v1 <= p.nvec && pc.v0 < pc.v1, "bad piece");
            if (last_plane >= 0) CHECK(reverse ? pc.plane == last_plane - 1 : pc.plane == last_plane + 1, "piece order");
            last_plane = pc.plane;
            touched[pc.plane]++;
            sample[pc.plane / C] += (unsigned long long)(pc.v1 - pc.v0);
            if (G > 32) {
                const PlaneShare sh = plane_share(g, pc.plane);
                const int64_t cta = gi;
                CHECK(cta >= sh.first && cta < sh.first + sh.count && sh.count <= p.slots, "plane_share: cta %lld first %lld count %d slots %d",
                      (long long)cta, (long long)sh.first, sh.count, p.slots);
                if (pc.v1 - pc.v0 == p.nvec) CHECK(sh.count == 1, "whole plane but shared");
            } else {
                CHECK(pc.v0 == 0 && pc.v1 == p.nvec, "warp mode piece must be a whole plane");
            }
            // batches
            const Batches<G, VPT> bt(pc, reverse != 0);
            std::vector<int> c2((size_t)(pc.v1 - pc.v0), 0);
            for (int i = 0; i < bt.full; ++i)
                for (int t = 0; t < G; ++t)
                    for (int j = 0; j < VPT; ++j) {
                        const int idx = bt.begin(i) + t + j * G;
                        CHECK(idx >= pc.v0 && idx < pc.v1, "full batch index out of piece");
                        if (idx >= pc.v0 && idx < pc.v1) c2[idx - pc.v0]++;
                    }
            if (bt.rem)
                for (int idx = bt.ragged_lo(); idx < bt.ragged_hi(); ++idx) {
                    CHECK(idx >= pc.v0 && idx < pc.v1, "ragged index out of piece");
                    if (idx >= pc.v0 && idx < pc.v1) c2[idx - pc.v0]++;
                }
            for (size_t i = 0; i < c2.size(); ++i) CHECK(c2[i] == 1, "batch coverage %d at %zu", c2[i], i);
            for (int v = pc.v0; v < pc.v1; ++v) cover[(size_t)(pc.plane * p.nvec + v)]++;
        }
    }
    for (int64_t i = 0; i < p.total; ++i)
        if (cover[(size_t)i] != 1) { CHECK(false, "vector %lld covered %d times (N=%d C=%d M=%lld sms=%d rev=%d)", (long long)i, cover[(size_t)i], N, C, (long long)M, sms, reverse); break; }
    for (int64_t pl = 0; pl < p.planes; ++pl) {
        if (G > 32) CHECK(touched[pl] == plane_share(g, pl).count, "plane %lld touched %d, share %d", (long long)pl, touched[pl], plane_share(g, pl).count);
        else CHECK(touched[pl] == 1, "warp plane touched %d", touched[pl]);
    }
    for (int n = 0; n < N; ++n) CHECK(sample[n] == (unsigned long long)C * p.nvec, "sample ticket total");
}

// NHWC: PlanNhwc + the thread loops of kernels_nhwc.cuh (thread t of a piece starts at v0 + t and steps by
// `active`, VPT accesses per full step).  Every vector of every sample is visited exactly once, a thread
// meets one channel vector per piece, the CTAs sharing a sample are plane_share() and fit the slot bounds,
// and the per-sample tickets add up to the sample's vector count.
void check_nhwc(int N, int C, int64_t M, int dtype, int align, int sms, int vpt) {
    const PlanNhwc p = make_plan_nhwc(N, C, M, dtype, align, sms);
    if (!p.ok) { CHECK(C / p.vec > kThreadsPerBlock, "plan refused although cv = %d", C / p.vec); return; }
    CHECK(p.cv * p.vec == C && p.active % p.cv == 0 && p.active <= kThreadsPerBlock && p.active > kThreadsPerBlock - p.cv, "geometry cv=%d active=%d", p.cv, p.active);
    CHECK(p.grid >= 1 && p.grid <= sms * kBlocksPerSMNhwc, "grid %d", p.grid);
    const Workspace w = workspace_layout(N, C, M, dtype, 1);
    const Workspace w0 = workspace_layout(N, C, M, dtype, 0);
    CHECK(w.total >= w0.total, "NHWC workspace smaller than NCHW");
    CHECK((size_t)N * C * p.slots * 16 <= w.res_error - w.partials, "partials do not fit: slots %d", p.slots);
    Sweep g{};
    g.M = M; g.nvec = p.nvec; g.planes = N; g.total = p.total; g.per = p.per; g.slots = p.slots; g.reverse = 0;
    std::vector<int> cover((size_t)p.total, 0);
    std::vector<int> touched((size_t)N, 0);
    std::vector<unsigned long long> ticket((size_t)N, 0);
    for (int64_t b = 0; b < p.grid; ++b) {
        PieceIter<kThreads> it(g, b);
        Piece pc;
        while (it.next(pc)) {
            CHECK(pc.plane >= 0 && pc.plane < N && pc.v0 >= 0 && pc.v1 <= p.nvec && pc.v0 < pc.v1, "bad piece");
            touched[pc.plane]++;
            ticket[pc.plane] += (unsigned long long)(pc.v1 - pc.v0);
            const PlaneShare sh = plane_share(g, pc.plane);
            CHECK(b >= sh.first && b < sh.first + sh.count && sh.count <= p.slots, "sample share");
            for (int t = 0; t < p.active; ++t) {
                const int cvv = (int)((pc.v0 + (int64_t)t) % p.cv);
                int v = pc.v0 + t;
                for (; v + (vpt - 1) * p.active < pc.v1; v += vpt * p.active)
                    for (int j = 0; j < vpt; ++j) {
                        const int idx = v + j * p.active;
                        CHECK(idx % p.cv == cvv, "channel vector changed");
                        cover[(size_t)(pc.plane * p.nvec + idx)]++;
                    }
                for (; v < pc.v1; v += p.active) { CHECK(v % p.cv == cvv, "channel vector changed (tail)"); cover[(size_t)(pc.plane * p.nvec + v)]++; }
            }
        }
    }
    for (int64_t i = 0; i < p.total; ++i)
        if (cover[(size_t)i] != 1) { CHECK(false, "NHWC vector %lld covered %d times (N=%d C=%d M=%lld sms=%d)", (long long)i, cover[(size_t)i], N, C, (long long)M, sms); break; }
    for (int n = 0; n < N; ++n) {
        CHECK(touched[n] == plane_share(g, n).count, "sample %d touched %d, share %d", n, touched[n], plane_share(g, n).count);
        CHECK(ticket[n] == (unsigned long long)p.nvec, "sample ticket total");
    }
}

int main() {
    const int shapes[][4] = {{20, 64, 224, 224}, {20, 16, 96, 96}, {20, 1, 224, 224}, {3, 2, 160, 160}, {2, 3, 224, 224}, {4, 5, 56, 56},
                             {20, 1, 64, 64}, {5, 3, 30, 30}, {6, 2, 37, 41}, {2, 2, 512, 512}, {33, 7, 12, 12}, {4, 3, 5, 7}, {2, 1, 8, 8},
                             {4, 3, 1, 2}, {64, 16, 28, 28}, {32, 16, 192, 192}, {7, 5, 112, 112}, {1, 1, 1024, 1024}, {2, 1, 3, 3},
                             {600, 9, 64, 64}, {300, 40, 48, 48}};
    int cases = 0;
    for (auto& s : shapes)
        for (int dtype = 0; dtype < 2; ++dtype)
            for (int align : {32, 16, 4})
                for (int sms : {148, 1, 7, 132})
                    for (int rev = 0; rev < 2; ++rev) {
                        const int64_t M = (int64_t)s[2] * s[3];
                        if ((int64_t)s[0] * s[1] * M > (int64_t)40000000) { if (sms != 148 || align != 32) continue; }
                        const Plan p = make_plan(s[0], s[1], M, dtype, align, sms);
                        const int tensors_vpt1 = 32 / p.vec < 1 ? 1 : (32 / p.vec > 4 ? 4 : 32 / p.vec);       // vpt_for<VEC,1>
                        const int tensors_vpt2 = 32 / (p.vec * 2) < 1 ? 1 : (32 / (p.vec * 2) > 4 ? 4 : 32 / (p.vec * 2));
                        for (int vpt : {tensors_vpt1, tensors_vpt2}) {
                            if (vpt == 4) { check_case<256, 4>(s[0], s[1], M, dtype, align, sms, rev); check_case<32, 4>(s[0], s[1], M, dtype, align, sms, rev); }
                            if (vpt == 2) { check_case<256, 2>(s[0], s[1], M, dtype, align, sms, rev); check_case<32, 2>(s[0], s[1], M, dtype, align, sms, rev); }
                            if (vpt == 1) { check_case<256, 1>(s[0], s[1], M, dtype, align, sms, rev); check_case<32, 1>(s[0], s[1], M, dtype, align, sms, rev); }
                            ++cases;
                            if (vpt == tensors_vpt2 && rev == 0) {                    // the backward: 2 tensors per thread, 4 slices per resident CTA
                                if (vpt == 4) check_case<256, 4>(s[0], s[1], M, dtype, align, sms, rev, 4);
                                if (vpt == 2) check_case<256, 2>(s[0], s[1], M, dtype, align, sms, rev, 4);
                                if (vpt == 1) check_case<256, 1>(s[0], s[1], M, dtype, align, sms, rev, 4);
                                ++cases;
                            }
                        }
                    }
    const int nhwc_shapes[][4] = {{20, 64, 56, 56}, {4, 64, 48, 48}, {3, 16, 96, 96}, {2, 8, 64, 64}, {5, 256, 14, 14}, {3, 24, 40, 40},
                                  {6, 12, 20, 20}, {4, 5, 17, 19}, {2, 320, 9, 9}, {40, 32, 12, 12}, {700, 8, 4, 4}, {2, 2, 300, 300},
                                  {2, 2048, 3, 3}, {2, 4096, 2, 2}, {1, 3, 2, 1}, {32, 16, 96, 96}};
    for (auto& s : nhwc_shapes)
        for (int dtype = 0; dtype < 2; ++dtype)
            for (int align : {32, 16, 4})
                for (int sms : {148, 1, 7})
                    for (int vpt : {1, 2, 4}) {
                        check_nhwc(s[0], s[1], (int64_t)s[2] * s[3], dtype, align, sms, vpt);
                        ++cases;
                    }
    printf("%d cases, %d failures\n", cases, failures);
    return failures ? 1 : 0;
}
