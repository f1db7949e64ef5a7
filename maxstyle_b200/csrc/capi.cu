// extern "C" boundary of libmaxstyle_b200.so -- see include/maxstyle_b200.h for the contract.
#include "../../include/maxstyle_b200.h"
#include "kernels_nchw.cuh"
#include "tables.cuh"
#include "fused_fwd.cuh"
#include "resident_fwd.cuh"
#include "ring_fwd.cuh"
#include "cluster_fwd.cuh"
#include "pair_fwd.cuh"
#include "ce2d.cuh"
#include "kernels_nhwc.cuh"

#include <cuda_runtime.h>
#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>

namespace {

using namespace ms;

int sm_count() {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    return sms;
}

// largest power of two <= 32 dividing all the given addresses (nullptr entries are ignored)
inline int common_align(const void* a, const void* b = nullptr, const void* c = nullptr) {
    const uintptr_t bits = reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c);
    int al = 32;
    while (al > 1 && (bits & (uintptr_t)(al - 1))) al >>= 1;
    return al;
}

inline int check_launch() { return cudaGetLastError() == cudaSuccess ? MAXSTYLE_OK : MAXSTYLE_ERR_CUDA; }

int check_shape(int N, int C, int H, int W, int dtype, int layout) {
    if (N <= 0 || C <= 0 || H <= 0 || W <= 0) return MAXSTYLE_ERR_BAD_ARG;
    if ((int64_t)H * W < 2) return MAXSTYLE_ERR_BAD_ARG;          // unbiased variance needs M >= 2
    if (dtype != MAXSTYLE_F32 && dtype != MAXSTYLE_BF16) return MAXSTYLE_ERR_UNSUPPORTED;
    if (layout != MAXSTYLE_NCHW && layout != MAXSTYLE_NHWC) return MAXSTYLE_ERR_UNSUPPORTED;
    if (layout == MAXSTYLE_NHWC && C > 1 && !make_plan_nhwc(N, C, (int64_t)H * W, dtype, 32, 148).ok &&
        !make_plan_nhwc(N, C, (int64_t)H * W, dtype, 1, 148).ok)
        return MAXSTYLE_ERR_UNSUPPORTED;                        // more than 256 vectors per pixel
    return MAXSTYLE_OK;
}

// With one channel the two layouts are the same memory: the NCHW kernels take it.
inline bool is_nhwc(int layout, int C) { return layout == MAXSTYLE_NHWC && C > 1; }

int check_workspace(const void* ws, size_t bytes, const Workspace& w) {
    if (ws == nullptr || bytes < w.total || (reinterpret_cast<uintptr_t>(ws) & 255u)) return MAXSTYLE_ERR_WORKSPACE;
    return MAXSTYLE_OK;
}

// sweep flags (include/maxstyle_b200.h) -> device geometry
Sweep sweep_of(const Plan& p, int64_t M, int sweep) {
    Sweep g;
    g.M = M; g.nvec = p.nvec; g.planes = p.planes; g.total = p.total; g.per = p.per; g.slots = p.slots;
    g.reverse = (sweep & MAXSTYLE_SWEEP_REVERSE) ? 1 : 0;
    g.in_policy = (sweep & MAXSTYLE_SWEEP_X_KEEP) ? kPolicyKeep : ((sweep & MAXSTYLE_SWEEP_X_STREAM) ? kPolicyStream : kPolicyNormal);
    g.io_policy = (sweep & MAXSTYLE_SWEEP_IO_NORMAL) ? kPolicyNormal : kPolicyStream;
    g.pre_op = kPreNone; g.pre_param = 0.f; g.vslices = 0;
    return g;
}

// The activation fused in front of the layer (include/maxstyle_b200.h: MAXSTYLE_PRE_*) and where min / max of y go.
struct PreOp {
    int op = kPreNone;
    float param = 0.f;
    unsigned int* ymin = nullptr;
    unsigned int* ymax = nullptr;
};

// Vector accesses a thread keeps in flight per tensor: sized so the fp32 copies of the loaded
// values fit a 32-register budget (the kernels run 4 CTAs x 256 threads per SM, 64 regs/thread).
template <int VEC, int TENSORS> constexpr int vpt_for() {
    constexpr int v = 32 / (VEC * TENSORS);
    return v < 1 ? 1 : (v > 4 ? 4 : v);
}

// ---- dispatch on (dtype, vector width, group size) ------------------------------------------
template <typename T, int VEC, int G>
void launch_stats(const void* x, float* mu, float* sig, TableRef tr, char* ws, const Workspace& w, const Plan& p, int64_t M,
                  float eps, int sweep, cudaStream_t s, const PreOp& pre) {
    Sweep g = sweep_of(p, M, sweep);
    g.pre_op = pre.op; g.pre_param = pre.param;
    stats_nchw_kernel<T, VEC, G, vpt_for<VEC, 1>()><<<p.grid, kThreads, 0, s>>>(
        static_cast<const T*>(x), mu, sig, tr, reinterpret_cast<float4*>(ws + w.partials),
        reinterpret_cast<unsigned long long*>(ws + w.plane_tickets), g, eps);
}

template <typename T, int VEC, int G>
void launch_apply(const void* x, void* y, const float* mu, TableRef tr, const float* scale, const float* shift, const Plan& p,
                  int64_t M, int sweep, cudaStream_t s, const PreOp& pre) {
    Sweep g = sweep_of(p, M, sweep);
    g.pre_op = pre.op; g.pre_param = pre.param;
    apply_nchw_kernel<T, VEC, G, vpt_for<VEC, 1>()><<<p.grid, kThreads, 0, s>>>(static_cast<const T*>(x), static_cast<T*>(y), mu, tr,
                                                                                scale, shift, g, pre.ymin, pre.ymax);
}

template <typename T, int VEC, int G>
void launch_bwd(const void* dy, const void* x, void* dx, char* ws, const Workspace& w, const Plan& p, int64_t M,
                const BwdTables& tb, const StepArgs& st, int sweep, cudaStream_t s, const PreOp& pre) {
    float4* partials = reinterpret_cast<float4*>(ws + w.partials);
    unsigned long long* tickets = reinterpret_cast<unsigned long long*>(ws + w.sample_tickets);
    int* done = reinterpret_cast<int*>(ws + w.done_counter);
    Sweep g = sweep_of(p, M, sweep);
    g.pre_op = pre.op; g.pre_param = pre.param;
    // CTA mode: p.grid may be a multiple of what fits on the device (make_plan's `oversub`); the resident CTAs share the slices
    int grid = p.grid;
    if (G > 32 && p.resident > 0 && p.grid > p.resident) { g.vslices = p.grid; grid = p.resident; }
    if (dx)
        bwd_nchw_kernel<T, VEC, G, vpt_for<VEC, 2>(), true><<<grid, kThreads, 0, s>>>(
            static_cast<const T*>(dy), static_cast<const T*>(x), static_cast<T*>(dx), partials, tickets, done, g, tb, st);
    else
        bwd_nchw_kernel<T, VEC, G, vpt_for<VEC, 2>(), false><<<grid, kThreads, 0, s>>>(
            static_cast<const T*>(dy), static_cast<const T*>(x), nullptr, partials, tickets, done, g, tb, st);
}

// ---- NHWC dispatch on (dtype, vector width) -----------------------------------------------------------
Sweep sweep_of_nhwc(const PlanNhwc& p, int N, int64_t M, int sweep) {
    Sweep g;
    g.M = M; g.nvec = p.nvec; g.planes = N; g.total = p.total; g.per = p.per; g.slots = p.slots;
    g.reverse = 0;                                              // the NHWC kernels sweep forward only
    g.in_policy = (sweep & MAXSTYLE_SWEEP_X_KEEP) ? kPolicyKeep : ((sweep & MAXSTYLE_SWEEP_X_STREAM) ? kPolicyStream : kPolicyNormal);
    g.io_policy = (sweep & MAXSTYLE_SWEEP_IO_NORMAL) ? kPolicyNormal : kPolicyStream;
    g.pre_op = kPreNone; g.pre_param = 0.f; g.vslices = 0;
    return g;
}

RowGeom geom_of(const PlanNhwc& p, int C) {
    RowGeom r;
    r.C = C; r.cv = p.cv; r.active = p.active;
    r.shuffle = (p.cv < 32 && 32 % p.cv == 0) ? 1 : 0;
    return r;
}

// vectors in flight per thread and tensor: 32 fp32 values' worth of loads for one tensor, 16 per tensor for two
template <int VEC, int TENSORS> constexpr int vpt_nhwc() {
    constexpr int v = (TENSORS == 1 ? 32 : 16) / VEC;
    return v < 1 ? 1 : (v > 4 ? 4 : v);
}

template <typename T, int VEC>
void launch_stats_nhwc(const void* x, float* mu, float* sig, TableRef tr, char* ws, const Workspace& w, const PlanNhwc& p, int N,
                       int64_t M, float eps, int sweep, cudaStream_t s) {
    stats_nhwc_kernel<T, VEC, vpt_nhwc<VEC, 1>()><<<p.grid, kThreads, 0, s>>>(
        static_cast<const T*>(x), mu, sig, tr, reinterpret_cast<float4*>(ws + w.partials),
        reinterpret_cast<unsigned long long*>(ws + w.plane_tickets), sweep_of_nhwc(p, N, M, sweep), geom_of(p, tr.C), eps);
}

template <typename T, int VEC>
void launch_apply_nhwc(const void* x, void* y, const float* mu, TableRef tr, const float* scale, const float* shift,
                       const PlanNhwc& p, int N, int64_t M, int sweep, cudaStream_t s) {
    apply_nhwc_kernel<T, VEC, vpt_nhwc<VEC, 1>()><<<p.grid, kThreads, 0, s>>>(
        static_cast<const T*>(x), static_cast<T*>(y), mu, tr, scale, shift, sweep_of_nhwc(p, N, M, sweep), geom_of(p, tr.C));
}

template <typename T, int VEC>
void launch_bwd_nhwc(const void* dy, const void* x, void* dx, char* ws, const Workspace& w, const PlanNhwc& p, int N, int64_t M,
                     const BwdTables& tb, const StepArgs& st, int sweep, cudaStream_t s) {
    float4* partials = reinterpret_cast<float4*>(ws + w.partials);
    unsigned long long* tickets = reinterpret_cast<unsigned long long*>(ws + w.sample_tickets);
    int* done = reinterpret_cast<int*>(ws + w.done_counter);
    const Sweep g = sweep_of_nhwc(p, N, M, sweep);
    const RowGeom rg = geom_of(p, tb.C);
    if (dx)
        bwd_nhwc_kernel<T, VEC, vpt_nhwc<VEC, 2>(), true><<<p.grid, kThreads, 0, s>>>(
            static_cast<const T*>(dy), static_cast<const T*>(x), static_cast<T*>(dx), partials, tickets, done, g, rg, tb, st);
    else
        bwd_nhwc_kernel<T, VEC, vpt_nhwc<VEC, 2>(), false><<<p.grid, kThreads, 0, s>>>(
            static_cast<const T*>(dy), static_cast<const T*>(x), nullptr, partials, tickets, done, g, rg, tb, st);
}

#define MS_DISPATCH_NHWC(FN, dtype, plan, ...)                                                     \
    do {                                                                                           \
        if ((dtype) == MAXSTYLE_F32) {                                                             \
            if ((plan).vec == 8) FN<float, 8>(__VA_ARGS__);                                        \
            else if ((plan).vec == 4) FN<float, 4>(__VA_ARGS__);                                   \
            else FN<float, 1>(__VA_ARGS__);                                                        \
        } else {                                                                                   \
            if ((plan).vec == 8) FN<__nv_bfloat16, 8>(__VA_ARGS__);                                \
            else FN<__nv_bfloat16, 1>(__VA_ARGS__);                                                \
        }                                                                                          \
    } while (0)

#define MS_DISPATCH_G(FN, T, V, plan, ...)                                                         \
    do {                                                                                           \
        if ((plan).group == 32) FN<T, V, 32>(__VA_ARGS__); else FN<T, V, 256>(__VA_ARGS__);        \
    } while (0)

#define MS_DISPATCH(FN, dtype, plan, ...)                                                          \
    do {                                                                                           \
        if ((dtype) == MAXSTYLE_F32) {                                                             \
            if ((plan).vec == 8) MS_DISPATCH_G(FN, float, 8, plan, __VA_ARGS__);                   \
            else if ((plan).vec == 4) MS_DISPATCH_G(FN, float, 4, plan, __VA_ARGS__);              \
            else MS_DISPATCH_G(FN, float, 1, plan, __VA_ARGS__);                                   \
        } else {                                                                                   \
            if ((plan).vec == 16) MS_DISPATCH_G(FN, __nv_bfloat16, 16, plan, __VA_ARGS__);         \
            else if ((plan).vec == 8) MS_DISPATCH_G(FN, __nv_bfloat16, 8, plan, __VA_ARGS__);      \
            else MS_DISPATCH_G(FN, __nv_bfloat16, 1, plan, __VA_ARGS__);                           \
        }                                                                                          \
    } while (0)

StepArgs to_step_args(const maxstyle_step_t* s) {
    StepArgs a{};
    if (s == nullptr) return a;
    a.mode = s->mode; a.maximize = s->maximize; a.update_noise = s->update_noise; a.update_mix = s->update_mix;
    a.lr = s->lr; a.beta1 = s->beta1; a.beta2 = s->beta2; a.eps = s->eps; a.t = s->t; a.step_dev = s->step_dev;
    a.gamma_noise = s->gamma_noise; a.beta_noise = s->beta_noise; a.lmda = s->lmda;
    a.gamma_m = s->gamma_m; a.gamma_v = s->gamma_v; a.beta_m = s->beta_m; a.beta_v = s->beta_v;
    a.lmda_m = s->lmda_m; a.lmda_v = s->lmda_v;
    return a;
}

int check_step(const maxstyle_step_t* s) {
    if (s == nullptr || s->mode == MAXSTYLE_STEP_NONE) return MAXSTYLE_OK;
    if (s->mode != MAXSTYLE_STEP_ADAM && s->mode != MAXSTYLE_STEP_SIGN) return MAXSTYLE_ERR_BAD_ARG;
    const bool adam = s->mode == MAXSTYLE_STEP_ADAM;
    if (s->update_noise) {
        if (!s->gamma_noise || !s->beta_noise) return MAXSTYLE_ERR_BAD_ARG;
        if (adam && (!s->gamma_m || !s->gamma_v || !s->beta_m || !s->beta_v)) return MAXSTYLE_ERR_BAD_ARG;
    }
    if (s->update_mix) {
        if (!s->lmda) return MAXSTYLE_ERR_BAD_ARG;
        if (adam && (!s->lmda_m || !s->lmda_v)) return MAXSTYLE_ERR_BAD_ARG;
    }
    if (adam && s->step_dev == nullptr && s->t < 1) return MAXSTYLE_ERR_BAD_ARG;
    return MAXSTYLE_OK;
}

// ---- fused forward ------------------------------------------------------------------------------------
struct FwdCall {
    const void* x; void* y; float *mu, *sig; const int64_t* perm; const float *lmda, *gamma_noise, *beta_noise;
    float *gamma_std, *beta_std, *scale, *shift; int N, C; int64_t M; int dtype, flags; float eps;
    char* ws; cudaStream_t stream;
    // statistics tables: [n_global, ld] with this rank's rows at row_offset (single GPU: n_global == N, ld == C, no peers)
    int n_global, row_offset, ld;
    PeerTables pt;
    PreOp pre;
};

void fill_tables(FusedArgs& a, const FwdCall& f) {
    a.mu = f.mu; a.sig = f.sig; a.scale = f.scale; a.shift = f.shift;
    a.n_global = f.n_global; a.row_offset = f.row_offset; a.ld = f.ld; a.pt = f.pt;
}

template <typename T, int VEC>
void launch_fused(const FwdCall& f, const FusedArgs& a, int grid) {
    fwd_fused_kernel<T, VEC, vpt_for<VEC, 1>()><<<grid, kThreads, 0, f.stream>>>(static_cast<const T*>(f.x), static_cast<T*>(f.y), a);
}

// Returns MAXSTYLE_OK after launching, -1 when this problem does not qualify (the caller then takes the
// two-pass path), or an error code.
// First channel whose x the forward leaves in L2 (evict-last on the apply pass's re-read) for the backward sweep that follows:
// the last MAXSTYLE_FUSED_KEEP_MB of the tensor.  Off by default: measured on the config-1 step it moves the backward from
// 133.6 to 132.4 us at 64 MB and costs the forward as much (profiles/r01_keep.txt) -- the backward is not limited by its
// DRAM reads of x.
int keep_from_channel(int N, int C, int64_t M, int dtype, bool keep_x) {
    const int64_t keep_bytes = tunables().fused_keep_bytes;
    if (!keep_x || keep_bytes <= 0) return C;
    const int64_t channel_bytes = (int64_t)N * M * elem_size(dtype);
    const int64_t k = keep_bytes / channel_bytes;
    return k >= C ? 0 : (int)(C - k);
}

int try_fused_fwd(const FwdCall& f, const Workspace& w, int sms, bool force, bool keep_x) {
    const FusedPlan fp = make_fused_plan(f.N, f.C, f.M, f.dtype, common_align(f.x, f.y));
    if (!fp.ok || (!fp.profitable && !force)) return -1;
    FusedArgs a;
    a.N = f.N; a.C = f.C; a.M = f.M;
    a.nvec = fp.nvec; a.pieces = fp.pieces; a.piece_vecs = fp.piece_vecs; a.items_per_channel = fp.items_per_channel;
    a.window = fp.window; a.chunk = fp.chunk; a.total_items = fp.total_items;
    a.keep_from = keep_from_channel(f.N, f.C, f.M, f.dtype, keep_x);
    a.flags = f.flags; a.eps = f.eps;
    fill_tables(a, f);
    a.perm = f.perm; a.lmda = f.lmda; a.gamma_noise = f.gamma_noise; a.beta_noise = f.beta_noise;
    a.gamma_std = f.gamma_std; a.beta_std = f.beta_std;
    a.partials = reinterpret_cast<float4*>(f.ws + w.res_partials);
    unsigned int* fl = reinterpret_cast<unsigned int*>(f.ws + w.res_flags);
    a.arrived = fl; a.ready = fl + f.C;
    a.error = reinterpret_cast<int*>(f.ws + w.res_error);
    a.queue = reinterpret_cast<unsigned long long*>(f.ws + w.res_error + 8);
    a.done = reinterpret_cast<unsigned int*>(f.ws + w.res_error + 16);
    int64_t cap = (int64_t)sms * kBlocksPerSM;
    const int64_t units = ceil_div(fp.total_items, fp.chunk);
    const int grid = (int)(units < cap ? units : cap);
    if (f.dtype == MAXSTYLE_F32) {
        if (fp.vec == 8) launch_fused<float, 8>(f, a, grid); else launch_fused<float, 4>(f, a, grid);
    } else {
        if (fp.vec == 16) launch_fused<__nv_bfloat16, 16>(f, a, grid); else launch_fused<__nv_bfloat16, 8>(f, a, grid);
    }
    return check_launch();
}

// ---- streamed forward (TMA ring) ------------------------------------------------------------------------
template <typename T>
int launch_ring(const FwdCall& f, const FusedArgs& a, const RingPlan& rp, int sms) {
    auto kern = fwd_ring_kernel<T>;
    static std::mutex mu;
    static std::map<std::pair<int, int>, int> cache;          // (device, smem) -> CTAs per SM
    int dev = 0, k = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    {
        std::lock_guard<std::mutex> lock(mu);
        const auto key = std::make_pair(dev, rp.smem);
        const auto it = cache.find(key);
        if (it != cache.end()) k = it->second;
        else {
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kRingCtrl + kRingMaxStages * kRingChunk) != cudaSuccess ||
                cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) != cudaSuccess ||
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&k, kern, kThreads, (size_t)rp.smem) != cudaSuccess) {
                cudaGetLastError();
                k = 0;
            }
            cache[key] = k;
        }
    }
    if (k <= 0) return -1;
    const int64_t cap = (int64_t)sms * k;
    const int grid = (int)(rp.total_items < cap ? rp.total_items : cap);
    RingGeom rg{rp.stages, rp.plane_bytes, rp.piece_bytes};
    kern<<<grid, kThreads, rp.smem, f.stream>>>(static_cast<const T*>(f.x), static_cast<T*>(f.y), a, rg);
    return check_launch();
}

// Same contract as try_fused_fwd.
int try_ring_fwd(const FwdCall& f, const Workspace& w, int sms, bool force, bool keep_x) {
    const RingPlan rp = make_ring_plan(f.N, f.C, f.M, f.dtype, common_align(f.x, f.y));
    if (!rp.ok || (!rp.profitable && !force)) return -1;
    FusedArgs a{};
    a.N = f.N; a.C = f.C; a.M = f.M;
    a.pieces = rp.pieces; a.items_per_channel = rp.items_per_channel;
    a.window = rp.window; a.chunk = 1; a.total_items = rp.total_items;
    a.keep_from = keep_from_channel(f.N, f.C, f.M, f.dtype, keep_x);
    a.flags = f.flags; a.eps = f.eps;
    fill_tables(a, f);
    a.perm = f.perm; a.lmda = f.lmda; a.gamma_noise = f.gamma_noise; a.beta_noise = f.beta_noise;
    a.gamma_std = f.gamma_std; a.beta_std = f.beta_std;
    a.partials = reinterpret_cast<float4*>(f.ws + w.res_partials);
    unsigned int* fl = reinterpret_cast<unsigned int*>(f.ws + w.res_flags);
    a.arrived = fl; a.ready = fl + f.C;
    a.error = reinterpret_cast<int*>(f.ws + w.res_error);
    a.queue = reinterpret_cast<unsigned long long*>(f.ws + w.res_error + 8);
    a.done = reinterpret_cast<unsigned int*>(f.ws + w.res_error + 16);
    return f.dtype == MAXSTYLE_F32 ? launch_ring<float>(f, a, rp, sms) : launch_ring<__nv_bfloat16>(f, a, rp, sms);
}

// ---- resident forward ---------------------------------------------------------------------------------
// CTAs of `kern` that fit one SM with `smem` bytes of dynamic shared memory (cached per device / kernel / size;
// the first call also raises the kernel's dynamic shared memory limit).
template <typename K>
int resident_ctas_per_sm(K kern, int threads, int smem) {
    static std::mutex mu;
    static std::map<std::tuple<int, const void*, int>, int> cache;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    const auto key = std::make_tuple(dev, reinterpret_cast<const void*>(kern), smem);
    std::lock_guard<std::mutex> lock(mu);
    const auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    int k = 0;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kResidentMaxSmem) != cudaSuccess ||
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&k, kern, threads, (size_t)smem) != cudaSuccess) {
        cudaGetLastError();
        k = 0;
    }
    cache[key] = k;
    return k;
}

template <typename T, int THREADS, int MINB>
int launch_resident(const FwdCall& f, ResidentArgs& a, const ResidentPlan& rp, int sms) {
    auto kern = fwd_resident_kernel<T, THREADS, MINB>;
    const int k = resident_ctas_per_sm(kern, THREADS, rp.smem);
    if (k <= 0) return -1;
    const int64_t cap = (int64_t)sms * k;
    const int grid = (int)(a.total_items < cap ? a.total_items : cap);
    if (f.N > grid) return -1;                                 // the co-residency argument needs N <= grid
    {   // same start-up stagger as the paired forward (try_pair_fwd): the CTAs of an SM load and store out of phase
        const int64_t env_ns = tunables().pair_stagger_ns;
        const int64_t ns = env_ns > 0 ? env_ns : (env_ns < 0 || k < 2 || a.total_items < 3ll * grid ? 0 : 3000);
        a.stagger_cycles = (int)(ns * 19 / 10);
        a.slot_div = sms;
    }
    kern<<<grid, THREADS, rp.smem, f.stream>>>(static_cast<const T*>(f.x), static_cast<T*>(f.y), a);
    return check_launch();
}

// Same contract as try_fused_fwd.
int try_resident_fwd(const FwdCall& f, const Workspace& w, int sms, int stats_sweep, int apply_sweep, bool force) {
    const ResidentPlan rp = make_resident_plan(f.N, f.C, f.M, f.dtype, common_align(f.x, f.y));
    if (!rp.ok || (!rp.preferred && !force)) return -1;
    ResidentArgs a;
    a.N = f.N; a.C = f.C; a.M = f.M;
    a.plane_bytes = rp.plane_bytes; a.chunk_bytes = rp.chunk_bytes; a.chunks = rp.chunks;
    a.total_items = (int64_t)f.N * f.C;
    a.flags = f.flags; a.eps = f.eps;
    // x passes through L2 once; "keep" leaves the most recent part of it there for the backward sweep
    a.in_policy = (stats_sweep & MAXSTYLE_SWEEP_X_KEEP) ? kPolicyKeep : kPolicyStream;
    a.io_policy = (apply_sweep & MAXSTYLE_SWEEP_IO_NORMAL) ? kPolicyNormal : kPolicyStream;
    a.mu = f.mu; a.sig = f.sig; a.scale = f.scale; a.shift = f.shift;
    a.perm = f.perm; a.lmda = f.lmda; a.gamma_noise = f.gamma_noise; a.beta_noise = f.beta_noise;
    a.gamma_std = f.gamma_std; a.beta_std = f.beta_std;
    a.ready = reinterpret_cast<unsigned int*>(f.ws + w.plane_ready);
    a.error = reinterpret_cast<int*>(f.ws + w.res_error);
    a.queue = reinterpret_cast<unsigned long long*>(f.ws + w.res_error + 8);
    a.done = reinterpret_cast<unsigned int*>(f.ws + w.res_error + 16);
    if (f.dtype == MAXSTYLE_F32)
        return rp.threads == 512 ? launch_resident<float, 512, 1>(f, a, rp, sms) : launch_resident<float, 256, 4>(f, a, rp, sms);
    return rp.threads == 512 ? launch_resident<__nv_bfloat16, 512, 1>(f, a, rp, sms)
                             : launch_resident<__nv_bfloat16, 256, 4>(f, a, rp, sms);
}

// ---- cluster-resident forward -------------------------------------------------------------------------
template <typename T>
cudaError_t cluster_config(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, int grid, int cs, int smem, cudaStream_t s) {
    cfg = cudaLaunchConfig_t{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kClThreads);
    cfg.dynamicSmemBytes = (size_t)smem;
    cfg.stream = s;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaSuccess;
}

// Clusters of `cs` CTAs with `smem` bytes each that the device keeps resident at once (cached per device / type / geometry).
template <typename T>
int cluster_capacity(int cs, int smem) {
    static std::mutex mu;
    static std::map<std::tuple<int, int, int>, int> cache;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    const auto key = std::make_tuple(dev, cs, smem);
    std::lock_guard<std::mutex> lock(mu);
    const auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    auto kern = fwd_cluster_kernel<T>;
    int n = 0;
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    cluster_config<T>(cfg, attr, cs, cs, smem, nullptr);
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kResidentMaxSmem) != cudaSuccess ||
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) != cudaSuccess ||
        cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) {
        cudaGetLastError();
        n = 0;
    }
    cache[key] = n;
    return n;
}

struct ClusterChoice {
    ClusterPlan plan;
    int clusters;      // G
    int use_order;     // visit the samples in cycle order of perm
};

// What the launch has to respect: `whole_channel` -- an item waits for planes anywhere in its channel in the order the ranks
// agree on (first forward: batch std; multi GPU: partners on other ranks) -- needs all N*P items of a channel inside one
// round of clusters; otherwise the samples can be visited in cycle order of perm, where an item only waits for its plane's
// other pieces and the next plane (2P items).
struct ClusterNeeds { int N; bool whole_channel; bool has_partner; };

// Pick cluster size and pieces: every SM busy, >= 3-4 stages per CTA, few idle slots in the last round, few pieces.
template <typename T>
ClusterChoice choose_cluster(int64_t planes, int64_t M, int dtype, int align, int sms, int sweep, ClusterNeeds need) {
    const int64_t env_cs = tunables().cluster_cs;
    const int64_t env_stages = tunables().cluster_stages;
    const int64_t env_pieces = tunables().cluster_pieces;
    const int sweep_cs = (sweep >> MAXSTYLE_SWEEP_CLUSTER_SIZE_SHIFT) & 15, sweep_stages = (sweep >> MAXSTYLE_SWEEP_CLUSTER_STAGES_SHIFT) & 7;
    const int sweep_pieces = (sweep >> MAXSTYLE_SWEEP_CLUSTER_PIECES_SHIFT) & 63;
    const int64_t force_cs = sweep_cs ? sweep_cs : env_cs;
    const int64_t max_stages = sweep_stages ? sweep_stages : env_stages;
    const int64_t force_pieces = sweep_pieces ? sweep_pieces : env_pieces;
    static const int piece_options[] = {1, 2, 3, 4, 6, 8, 12, 16, 24, 32};
    ClusterChoice best{};
    double best_score = 0.0;
    for (int cs = 1; cs <= kClMaxCluster; cs *= 2) {
        if (force_cs > 0 && cs != force_cs) continue;
        for (int P : piece_options) {
            if (force_pieces > 0 && P != force_pieces) continue;
            const ClusterPlan cp = make_cluster_plan(M, dtype, align, cs, P, (int)max_stages);
            if (!cp.ok) continue;
            const int64_t items = planes * P;
            int64_t G = cluster_capacity<T>(cs, cp.smem);
            if (G <= 0) continue;
            if (G > items) G = items;
            int use_order = 0;
            if ((int64_t)need.N * P > G && (need.has_partner || need.whole_channel)) {
                if (need.whole_channel || need.N > kClMaxN || 2 * P > G) continue;
                use_order = 1;
            }
            const double sm_use = (double)(G * cs) / sms > 1.0 ? 1.0 : (double)(G * cs) / sms;
            const double quant = (double)items / (double)(ceil_div(items, G) * G);
            const double stage = cp.stages >= 4 ? 1.0 : (cp.stages == 3 ? 0.97 : (cp.stages == 2 ? 0.88 : 0.72));
            const double piece = 1.0 - 0.004 * (P - 1) - (cs > 2 ? 0.02 : 0.0);
            const double score = sm_use * quant * stage * piece;
            if (score > best_score) { best_score = score; best.plan = cp; best.clusters = (int)G; best.use_order = use_order; }
        }
    }
    return best;
}

template <typename T>
int launch_cluster(const FwdCall& f, const Workspace& w, int sms, bool force, int sweep) {
    const bool multi = f.pt.world > 1, first = (f.flags & MAXSTYLE_COMPUTE_BATCH_STD) != 0;
    const bool mix = (f.flags & MAXSTYLE_MIX_STYLE) != 0;
    if (first && f.n_global > kClusterStdRows) return -1;
    const ClusterNeeds need{f.N, first || multi, mix};
    const ClusterChoice ch = choose_cluster<T>((int64_t)f.N * f.C, f.M, f.dtype, common_align(f.x, f.y), sms, sweep, need);
    if (!ch.plan.ok || ch.clusters <= 0) return -1;
    const ClusterPlan& cp = ch.plan;
    if (cp.pieces > cluster_max_pieces(f.M, f.dtype)) return -1;
    const int64_t enabled = tunables().cluster_enabled;
    if (!force && enabled == 2) return -1;                    // MAXSTYLE_CLUSTER=2: only when forced
    ClusterArgs a{};
    a.N = f.N; a.C = f.C; a.M = f.M;
    a.plane_bytes = cp.plane_bytes; a.pieces = cp.pieces; a.part_bytes = cp.part_bytes; a.part_stride = cp.part_stride;
    a.chunk_bytes = cp.chunk_bytes; a.chunks = cp.chunks; a.stages = cp.stages; a.cluster = cp.cluster;
    a.num_clusters = ch.clusters; a.use_order = ch.use_order; a.total_items = (int64_t)f.N * f.C * cp.pieces;
    a.flags = f.flags; a.eps = f.eps;
    a.in_policy = kPolicyStream; a.io_policy = kPolicyStream;
    a.mu = f.mu; a.sig = f.sig; a.scale = f.scale; a.shift = f.shift;
    a.n_global = f.n_global; a.row_offset = f.row_offset; a.ld = f.ld;
    a.perm = f.perm; a.lmda = f.lmda; a.gamma_noise = f.gamma_noise; a.beta_noise = f.beta_noise;
    a.gamma_std = f.gamma_std; a.beta_std = f.beta_std;
    a.ll = reinterpret_cast<uint2*>(f.ws + w.cl_words);
    a.piece_ll = reinterpret_cast<uint2*>(f.ws + w.cl_pieces);
    a.epoch = multi ? f.pt.epoch : reinterpret_cast<unsigned int*>(f.ws + w.res_error + 24);
    a.done = reinterpret_cast<unsigned int*>(f.ws + w.res_error + 16);
    a.error = reinterpret_cast<int*>(f.ws + w.res_error);
    a.pt = f.pt;
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    cluster_config<T>(cfg, attr, ch.clusters * cp.cluster, cp.cluster, cp.smem, f.stream);
    if (cudaLaunchKernelEx(&cfg, fwd_cluster_kernel<T>, static_cast<const T*>(f.x), static_cast<T*>(f.y), a) != cudaSuccess) {
        cudaGetLastError();
        return MAXSTYLE_ERR_CUDA;
    }
    return check_launch();
}

// Same contract as try_fused_fwd.  `preferred`: shapes where this kernel is the default.
int try_cluster_fwd(const FwdCall& f, const Workspace& w, int sms, int sweep) {
    const bool force = (sweep & MAXSTYLE_SWEEP_FORCE_CLUSTER) != 0;
    const int64_t pb = f.M * elem_size(f.dtype);
    if (f.N < 2 || pb < kClusterMinPlaneBytes) return -1;
    const int64_t enabled = tunables().cluster_enabled;
    if (!force && enabled == 0) return -1;
    return f.dtype == MAXSTYLE_F32 ? launch_cluster<float>(f, w, sms, force, sweep) : launch_cluster<__nv_bfloat16>(f, w, sms, force, sweep);
}

// ---- paired forward -----------------------------------------------------------------------------------
template <typename T, int VEC>
void launch_pair(const FwdCall& f, const PairArgs& a, int grid) {
    const int64_t minb = tunables().pair_minb;
    const T* xs = static_cast<const T*>(f.x);
    T* ys = static_cast<T*>(f.y);
    constexpr int VPT = vpt_for<VEC, 1>();
    if (minb == 3) fwd_pair_kernel<T, VEC, VPT, 3><<<grid, kThreads, 0, f.stream>>>(xs, ys, a);
    else fwd_pair_kernel<T, VEC, VPT, 4><<<grid, kThreads, 0, f.stream>>>(xs, ys, a);
}

struct PairChoice { PairPlan plan; int grid; int use_order; };

PairChoice choose_pair(int N, int C, int64_t M, int dtype, int align, int sms, int sweep, bool whole_channel, bool has_partner,
                       int n_global) {
    PairChoice ch{};
    const int force_pieces = (sweep >> MAXSTYLE_SWEEP_CLUSTER_PIECES_SHIFT) & 63;
    ch.plan = make_pair_plan(M, dtype, align, force_pieces, (int64_t)N * C);
    if (!ch.plan.ok) return ch;
    const int64_t minb = tunables().pair_minb;
    const int64_t items = (int64_t)N * C * ch.plan.pieces, cap = (int64_t)sms * (minb == 3 ? 3 : 4);
    ch.grid = (int)(items < cap ? items : cap);
    // An item waits for items within W positions: grid > W keeps a CTA free for the lowest missing one.  W = N * P when the
    // whole channel is awaited (first forward) or the samples are taken in natural order; 2 * P when they are taken in cycle
    // order of the global perm (pair_fwd.cuh) -- on one GPU and, in the steady state, across ranks.
    const bool can_order = !whole_channel && N <= kPairMaxN && n_global <= kPairMaxNG && 2 * ch.plan.pieces < ch.grid;
    if ((int64_t)N * ch.plan.pieces >= ch.grid && (has_partner || whole_channel)) {
        if (!can_order) { ch.plan.ok = false; return ch; }
        ch.use_order = 1;
    }
    // experiment: walk the samples in cycle order whenever it is allowed (partner planes are then taken back to back)
    const int64_t force_order = tunables().pair_order;
    if (force_order == 1 && has_partner && can_order) ch.use_order = 1;
    return ch;
}

// Where the paired kernel is the default: tensors that do not sit comfortably in L2 anyway (>= 16 MB); smaller ones are
// launch-latency-bound and the window / two-pass paths measured as fast or faster (profiles/r02_fwd_paths.txt).
bool pair_preferred(int N, int C, int64_t M, int dtype) {
    return (int64_t)N * C * M * elem_size(dtype) >= (16ll << 20);
}

// Same contract as try_fused_fwd.
int try_pair_fwd(const FwdCall& f, const Workspace& w, int sms, int sweep) {
    const bool force = (sweep & MAXSTYLE_SWEEP_FORCE_PAIR) != 0;
    const int64_t enabled = tunables().pair_enabled;
    if (!force && enabled != 1) return -1;
    if (f.N < 2) return -1;
    const bool multi = f.pt.world > 1, first = (f.flags & MAXSTYLE_COMPUTE_BATCH_STD) != 0;
    if (first && f.n_global > 32 * kPairStdRows) return -1;
    const PairChoice ch = choose_pair(f.N, f.C, f.M, f.dtype, common_align(f.x, f.y), sms, sweep, first, (f.flags & MAXSTYLE_MIX_STYLE) != 0,
                                      f.n_global);
    if (!ch.plan.ok || ch.plan.pieces > w.max_pieces) return -1;
    PairArgs a{};
    a.N = f.N; a.C = f.C; a.M = f.M;
    a.nvec = ch.plan.nvec; a.pieces = ch.plan.pieces; a.piece_vecs = ch.plan.piece_vecs; a.use_order = ch.use_order;
    a.total_items = (int64_t)f.N * f.C * ch.plan.pieces;
    // The k-th CTA of an SM starts k * 3 us late when every CTA has >= 3 items to go through: started together, the CTAs read
    // (pass 1) and write (pass 2) in a common rhythm, and HBM moves pure reads or pure writes ~20 % slower than a mix
    // (config 1: 110.1 -> 102.7 us, 64x32x512x512: 923 -> 868 us; profiles/r02_fwd_experiments.txt).  MAXSTYLE_PAIR_STAGGER_NS
    // overrides the delay (-1: none).
    {
        const int64_t env_ns = tunables().pair_stagger_ns;
        const int64_t ns = env_ns > 0 ? env_ns : (env_ns < 0 || a.total_items < 3ll * ch.grid ? 0 : 3000);
        a.stagger_cycles = (int)(ns * 19 / 10);
        a.slot_div = sms;
    }
    a.flags = f.flags; a.eps = f.eps;
    a.pol_first = (sweep & MAXSTYLE_SWEEP_X_STREAM) ? kPolicyNormal : kPolicyKeep;       // the piece is re-read microseconds later
    a.pol_second = kPolicyStream; a.pol_out = kPolicyStream;
    a.pre_op = f.pre.op; a.pre_param = f.pre.param; a.ymin = f.pre.ymin; a.ymax = f.pre.ymax;
    a.mu = f.mu; a.sig = f.sig; a.scale = f.scale; a.shift = f.shift;
    a.n_global = f.n_global; a.row_offset = f.row_offset; a.ld = f.ld;
    a.perm = f.perm; a.lmda = f.lmda; a.gamma_noise = f.gamma_noise; a.beta_noise = f.beta_noise;
    a.gamma_std = f.gamma_std; a.beta_std = f.beta_std;
    a.ll = reinterpret_cast<uint2*>(f.ws + w.cl_words);
    a.piece_ll = reinterpret_cast<uint2*>(f.ws + w.cl_pieces);
    a.wepoch = reinterpret_cast<unsigned int*>(f.ws + w.res_error + 24);
    a.epoch = multi ? f.pt.epoch : nullptr;
    a.queue = reinterpret_cast<unsigned long long*>(f.ws + w.res_error + 8);
    a.done = reinterpret_cast<unsigned int*>(f.ws + w.res_error + 16);
    a.error = reinterpret_cast<int*>(f.ws + w.res_error);
    a.pt = f.pt;
    if (f.dtype == MAXSTYLE_F32) {
        if (ch.plan.vec == 8) launch_pair<float, 8>(f, a, ch.grid); else launch_pair<float, 4>(f, a, ch.grid);
    } else {
        if (ch.plan.vec == 16) launch_pair<__nv_bfloat16, 16>(f, a, ch.grid); else launch_pair<__nv_bfloat16, 8>(f, a, ch.grid);
    }
    return check_launch();
}

}  // namespace

extern "C" {

const char* maxstyle_version(void) { return "maxstyle_b200 0.1 (sm_100a)"; }

const char* maxstyle_strerror(int code) {
    switch (code) {
        case MAXSTYLE_OK: return "ok";
        case MAXSTYLE_ERR_BAD_ARG: return "bad argument (null pointer, non-positive size, H*W < 2, or inconsistent rows)";
        case MAXSTYLE_ERR_UNSUPPORTED: return "unsupported dtype / layout / shape";
        case MAXSTYLE_ERR_WORKSPACE: return "workspace missing, too small or not 256-byte aligned";
        case MAXSTYLE_ERR_CUDA: return "CUDA launch error";
        case MAXSTYLE_ERR_NO_DEVICE: return "no usable CUDA device";
        case MAXSTYLE_ERR_TIMEOUT: return "a device-side wait of the fused forward timed out; its results are invalid";
        default: return "unknown error code";
    }
}

size_t maxstyle_workspace_bytes(int N, int C, int H, int W, int dtype, int layout) {
    if (check_shape(N, C, H, W, dtype, layout) != MAXSTYLE_OK) return 0;
    return workspace_layout(N, C, (int64_t)H * W, dtype, is_nhwc(layout, C)).total;
}

}  // extern "C"

namespace {
int stats_impl(const void* x, float* mu_all, float* sig_all, int table_ld, int row_offset, int N, int C, int H, int W,
               int dtype, int layout, float eps, int sweep, void* workspace, size_t workspace_bytes,
               maxstyle_stream_t stream, const PreOp& pre) {
    int rc = check_shape(N, C, H, W, dtype, layout);
    if (rc) return rc;
    if (!x || !mu_all || !sig_all || row_offset < 0 || table_ld < C) return MAXSTYLE_ERR_BAD_ARG;
    const int64_t M = (int64_t)H * W;
    const Workspace w = workspace_layout(N, C, M, dtype, is_nhwc(layout, C));
    if ((rc = check_workspace(workspace, workspace_bytes, w))) return rc;
    const int sms = sm_count();
    if (sms <= 0) return MAXSTYLE_ERR_NO_DEVICE;
    const TableRef tr{C, table_ld, row_offset};
    if (is_nhwc(layout, C)) {
        if (pre.op != kPreNone) return MAXSTYLE_ERR_UNSUPPORTED;       // the fused activation is built for NCHW
        const PlanNhwc pn = make_plan_nhwc(N, C, M, dtype, common_align(x), sms);
        if (!pn.ok) return MAXSTYLE_ERR_UNSUPPORTED;
        MS_DISPATCH_NHWC(launch_stats_nhwc, dtype, pn, x, mu_all, sig_all, tr, static_cast<char*>(workspace), w, pn, N, M, eps,
                         sweep, static_cast<cudaStream_t>(stream));
        return check_launch();
    }
    const Plan p = make_plan(N, C, M, dtype, common_align(x), sms);
    MS_DISPATCH(launch_stats, dtype, p, x, mu_all, sig_all, tr, static_cast<char*>(workspace), w, p, M, eps, sweep,
                static_cast<cudaStream_t>(stream), pre);
    return check_launch();
}
}  // namespace

extern "C" {

int maxstyle_stats(const void* x, float* mu_all, float* sig_all, int table_ld, int row_offset, int N, int C, int H, int W,
                   int dtype, int layout, float eps, int sweep, void* workspace, size_t workspace_bytes,
                   maxstyle_stream_t stream) {
    return stats_impl(x, mu_all, sig_all, table_ld, row_offset, N, C, H, W, dtype, layout, eps, sweep, workspace, workspace_bytes, stream,
                      PreOp{});
}

int maxstyle_tables(const float* mu_all, const float* sig_all, int table_ld, int N_global, int row_offset, int N, int C,
                    const int64_t* perm, const float* lmda, const float* gamma_noise, const float* beta_noise,
                    float* gamma_std, float* beta_std, int flags, float* scale, float* shift, maxstyle_stream_t stream) {
    if (!mu_all || !sig_all || !scale || !shift) return MAXSTYLE_ERR_BAD_ARG;
    if (N <= 0 || C <= 0 || N_global < N || row_offset < 0 || row_offset + N > N_global || table_ld < C)
        return MAXSTYLE_ERR_BAD_ARG;
    if ((flags & MAXSTYLE_MIX_STYLE) && (!perm || !lmda)) return MAXSTYLE_ERR_BAD_ARG;
    if (!(flags & MAXSTYLE_NO_NOISE) && (!gamma_noise || !beta_noise || !gamma_std || !beta_std)) return MAXSTYLE_ERR_BAD_ARG;
    tables_kernel<<<C, kTableThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        mu_all, sig_all, table_ld, N_global, row_offset, N, C, perm, lmda, gamma_noise, beta_noise, gamma_std, beta_std, flags,
        scale, shift);
    return check_launch();
}

size_t maxstyle_p2p_bytes(int N, int C, int world) {
    if (N <= 0 || C <= 0 || world <= 0) return 0;
    // two parities of {value, epoch} words for every (row, mu|sig, channel), then two parities of `world` barrier words
    return ((size_t)2 * N * world * 2 * C + (size_t)2 * world) * 8;
}

int maxstyle_rank_barrier(const uint64_t* peers, int rank, int world, int N, int C, uint32_t* bar_epoch, int* error,
                          maxstyle_stream_t stream) {
    if (!peers || !bar_epoch || !error || world < 1 || world > 32 || rank < 0 || rank >= world || N <= 0 || C <= 0)
        return MAXSTYLE_ERR_BAD_ARG;
    PeerTables pt;
    pt.peers = reinterpret_cast<const unsigned long long*>(peers);
    pt.rank = rank; pt.world = world; pt.epoch = nullptr; pt.done = nullptr; pt.error = error;
    rank_barrier_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(pt, (size_t)2 * N * world * 2 * C, bar_epoch);
    return check_launch();
}

int maxstyle_tables_p2p(const uint64_t* peers, int rank, int world, uint32_t* epoch, uint32_t* done, int* error,
                        float* mu_all, float* sig_all, int table_ld, int N_global, int row_offset, int N, int C,
                        const int64_t* perm, const float* lmda, const float* gamma_noise, const float* beta_noise,
                        float* gamma_std, float* beta_std, int flags, float* scale, float* shift, maxstyle_stream_t stream) {
    if (!peers || !epoch || !done || !error || !mu_all || !sig_all || !scale || !shift) return MAXSTYLE_ERR_BAD_ARG;
    if (world < 1 || world > kTableThreads || rank < 0 || rank >= world) return MAXSTYLE_ERR_BAD_ARG;
    if (N <= 0 || C <= 0 || N_global != N * world || row_offset != rank * N || table_ld < C) return MAXSTYLE_ERR_BAD_ARG;
    if ((flags & MAXSTYLE_MIX_STYLE) && (!perm || !lmda)) return MAXSTYLE_ERR_BAD_ARG;
    if (!(flags & MAXSTYLE_NO_NOISE) && (!gamma_noise || !beta_noise || !gamma_std || !beta_std)) return MAXSTYLE_ERR_BAD_ARG;
    PeerTables pt;
    pt.peers = reinterpret_cast<const unsigned long long*>(peers);
    pt.rank = rank; pt.world = world; pt.epoch = epoch; pt.done = done; pt.error = error;
    tables_p2p_kernel<<<C, kTableThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        pt, mu_all, sig_all, table_ld, N_global, row_offset, N, C, perm, lmda, gamma_noise, beta_noise, gamma_std, beta_std, flags,
        scale, shift);
    return check_launch();
}

}  // extern "C"

namespace {
int apply_impl(const void* x, void* y, const float* mu_all, int table_ld, int row_offset, const float* scale,
               const float* shift, int N, int C, int H, int W, int dtype, int layout, int sweep,
               maxstyle_stream_t stream, const PreOp& pre) {
    int rc = check_shape(N, C, H, W, dtype, layout);
    if (rc) return rc;
    if (!x || !y || !mu_all || !scale || !shift || row_offset < 0 || table_ld < C) return MAXSTYLE_ERR_BAD_ARG;
    const int sms = sm_count();
    if (sms <= 0) return MAXSTYLE_ERR_NO_DEVICE;
    const int64_t M = (int64_t)H * W;
    const TableRef tr{C, table_ld, row_offset};
    if (is_nhwc(layout, C)) {
        if (pre.op != kPreNone || pre.ymin != nullptr) return MAXSTYLE_ERR_UNSUPPORTED;
        const PlanNhwc pn = make_plan_nhwc(N, C, M, dtype, common_align(x, y), sms);
        if (!pn.ok) return MAXSTYLE_ERR_UNSUPPORTED;
        MS_DISPATCH_NHWC(launch_apply_nhwc, dtype, pn, x, y, mu_all, tr, scale, shift, pn, N, M, sweep,
                         static_cast<cudaStream_t>(stream));
        return check_launch();
    }
    const Plan p = make_plan(N, C, M, dtype, common_align(x, y), sms);
    MS_DISPATCH(launch_apply, dtype, p, x, y, mu_all, tr, scale, shift, p, M, sweep,
                static_cast<cudaStream_t>(stream), pre);
    return check_launch();
}
}  // namespace

extern "C" {

int maxstyle_apply(const void* x, void* y, const float* mu_all, int table_ld, int row_offset, const float* scale,
                   const float* shift, int N, int C, int H, int W, int dtype, int layout, int sweep,
                   maxstyle_stream_t stream) {
    return apply_impl(x, y, mu_all, table_ld, row_offset, scale, shift, N, C, H, W, dtype, layout, sweep, stream, PreOp{});
}

}  // extern "C"

namespace {
int fwd_impl(const void* x, void* y, float* mu, float* sig, const int64_t* perm, const float* lmda,
             const float* gamma_noise, const float* beta_noise, float* gamma_std, float* beta_std, float* scale,
             float* shift, int N, int C, int H, int W, int dtype, int layout, int flags, float eps, int stats_sweep,
             int apply_sweep, void* workspace, size_t workspace_bytes, maxstyle_stream_t stream, const PreOp& pre) {
    int rc = check_shape(N, C, H, W, dtype, layout);
    if (rc) return rc;
    if (!x || !y || !mu || !sig || !scale || !shift) return MAXSTYLE_ERR_BAD_ARG;
    if ((flags & MAXSTYLE_MIX_STYLE) && (!perm || !lmda)) return MAXSTYLE_ERR_BAD_ARG;
    if (!(flags & MAXSTYLE_NO_NOISE) && (!gamma_noise || !beta_noise || !gamma_std || !beta_std)) return MAXSTYLE_ERR_BAD_ARG;
    if (!(stats_sweep & MAXSTYLE_SWEEP_NO_FUSED) && (gamma_std && beta_std) && !is_nhwc(layout, C)) {
        const int64_t M = (int64_t)H * W;
        const Workspace w = workspace_layout(N, C, M, dtype);
        if ((rc = check_workspace(workspace, workspace_bytes, w))) return rc;
        const int sms = sm_count();
        if (sms <= 0) return MAXSTYLE_ERR_NO_DEVICE;
        FwdCall f{x, y, mu, sig, perm, lmda, gamma_noise, beta_noise, gamma_std, beta_std, scale, shift, N, C, M, dtype, flags, eps,
                  static_cast<char*>(workspace), static_cast<cudaStream_t>(stream), N, 0, C, PeerTables{}, pre};
        // Five single-kernel forwards, each the default where it measured fastest (profiles/r02_fwd_paths.txt):
        //   resident  planes of 16-108 KB, tensors >= 64 MB (two CTAs share an SM)               -- x crosses L2 -> SM once
        //   paired    steady state (cached batch std), tensors >= 16 MB, any plane >= 8 KB     -- piece re-read from L2 microseconds later
        //   window    the first forward of a module (whole-channel dependency), planes >= 64 KB -- ordered queue, 32 MB L2 window
        //   ring / cluster: only when forced (tests, experiments)
        const int force_other = stats_sweep & (MAXSTYLE_SWEEP_FORCE_RESIDENT | MAXSTYLE_SWEEP_FORCE_RING | MAXSTYLE_SWEEP_FORCE_WINDOW);
        const int force_any = force_other | (stats_sweep & (MAXSTYLE_SWEEP_FORCE_CLUSTER | MAXSTYLE_SWEEP_FORCE_PAIR));
        const bool first = (flags & MAXSTYLE_COMPUTE_BATCH_STD) != 0;
        const bool pair_allowed = !(stats_sweep & MAXSTYLE_SWEEP_NO_PAIR) && (!force_any || (stats_sweep & MAXSTYLE_SWEEP_FORCE_PAIR));
        const bool fused_neighbours = pre.op != kPreNone || pre.ymin != nullptr;       // only the paired and the two-pass kernels take them
        if ((stats_sweep & MAXSTYLE_SWEEP_FORCE_PAIR) || (fused_neighbours && pair_allowed)) {
            rc = try_pair_fwd(f, w, sms, stats_sweep);
            if (rc >= 0) return rc;
        }
        if (fused_neighbours) goto two_pass;
        if (!(stats_sweep & MAXSTYLE_SWEEP_NO_CLUSTER) && (stats_sweep & MAXSTYLE_SWEEP_FORCE_CLUSTER)) {
            rc = try_cluster_fwd(f, w, sms, stats_sweep);
            if (rc >= 0) return rc;
        }
        if (!(stats_sweep & MAXSTYLE_SWEEP_NO_RESIDENT)) {
            rc = try_resident_fwd(f, w, sms, stats_sweep, apply_sweep, (stats_sweep & MAXSTYLE_SWEEP_FORCE_RESIDENT) != 0);
            if (rc >= 0) return rc;
        }
        if (pair_allowed && !first && pair_preferred(N, C, M, dtype)) {
            rc = try_pair_fwd(f, w, sms, stats_sweep);
            if (rc >= 0) return rc;
        }
        if (!(stats_sweep & MAXSTYLE_SWEEP_NO_RING)) {
            rc = try_ring_fwd(f, w, sms, (stats_sweep & MAXSTYLE_SWEEP_FORCE_RING) != 0, (stats_sweep & MAXSTYLE_SWEEP_X_KEEP) != 0);
            if (rc >= 0) return rc;
        }
        rc = try_fused_fwd(f, w, sms, (stats_sweep & MAXSTYLE_SWEEP_FORCE_WINDOW) != 0, (stats_sweep & MAXSTYLE_SWEEP_X_KEEP) != 0);
        if (rc >= 0) return rc;
        if (pair_allowed && first && pair_preferred(N, C, M, dtype)) {          // first forward, no window: still one read of x
            rc = try_pair_fwd(f, w, sms, stats_sweep);
            if (rc >= 0) return rc;
        }
    }
two_pass:
    rc = stats_impl(x, mu, sig, C, 0, N, C, H, W, dtype, layout, eps, stats_sweep, workspace, workspace_bytes, stream, pre);
    if (rc) return rc;
    rc = maxstyle_tables(mu, sig, C, N, 0, N, C, perm, lmda, gamma_noise, beta_noise, gamma_std, beta_std, flags, scale, shift,
                         stream);
    if (rc) return rc;
    return apply_impl(x, y, mu, C, 0, scale, shift, N, C, H, W, dtype, layout, apply_sweep, stream, pre);
}
}  // namespace

extern "C" {

int maxstyle_fwd(const void* x, void* y, float* mu, float* sig, const int64_t* perm, const float* lmda,
                 const float* gamma_noise, const float* beta_noise, float* gamma_std, float* beta_std, float* scale,
                 float* shift, int N, int C, int H, int W, int dtype, int layout, int flags, float eps, int stats_sweep,
                 int apply_sweep, void* workspace, size_t workspace_bytes, maxstyle_stream_t stream) {
    return fwd_impl(x, y, mu, sig, perm, lmda, gamma_noise, beta_noise, gamma_std, beta_std, scale, shift, N, C, H, W, dtype, layout, flags,
                    eps, stats_sweep, apply_sweep, workspace, workspace_bytes, stream, PreOp{});
}

int maxstyle_fwd_act(const void* z, void* y, float* mu, float* sig, const int64_t* perm, const float* lmda,
                     const float* gamma_noise, const float* beta_noise, float* gamma_std, float* beta_std, float* scale,
                     float* shift, int N, int C, int H, int W, int dtype, int layout, int flags, float eps, int stats_sweep,
                     int apply_sweep, int pre_op, float pre_param, uint32_t* y_min, uint32_t* y_max,
                     void* workspace, size_t workspace_bytes, maxstyle_stream_t stream) {
    if (pre_op != MAXSTYLE_PRE_NONE && pre_op != MAXSTYLE_PRE_LEAKY_RELU && pre_op != MAXSTYLE_PRE_SIGMOID) return MAXSTYLE_ERR_BAD_ARG;
    if ((y_min == nullptr) != (y_max == nullptr)) return MAXSTYLE_ERR_BAD_ARG;
    PreOp pre;
    pre.op = pre_op; pre.param = pre_param; pre.ymin = y_min; pre.ymax = y_max;
    return fwd_impl(z, y, mu, sig, perm, lmda, gamma_noise, beta_noise, gamma_std, beta_std, scale, shift, N, C, H, W, dtype, layout, flags,
                    eps, stats_sweep, apply_sweep, workspace, workspace_bytes, stream, pre);
}

int maxstyle_fwd_p2p(const void* x, void* y, float* mu_all, float* sig_all, int table_ld, int N_global, int row_offset,
                     const int64_t* perm, const float* lmda, const float* gamma_noise, const float* beta_noise,
                     float* gamma_std, float* beta_std, float* scale, float* shift,
                     int N, int C, int H, int W, int dtype, int layout, int flags, float eps, int stats_sweep,
                     const uint64_t* peers, int rank, int world, uint32_t* epoch,
                     void* workspace, size_t workspace_bytes, maxstyle_stream_t stream) {
    int rc = check_shape(N, C, H, W, dtype, layout);
    if (rc) return rc;
    if (!x || !y || !mu_all || !sig_all || !scale || !shift || !peers || !epoch || !gamma_std || !beta_std) return MAXSTYLE_ERR_BAD_ARG;
    if (world < 2 || rank < 0 || rank >= world || N_global != N * world || row_offset != rank * N || table_ld < C)
        return MAXSTYLE_ERR_BAD_ARG;
    if ((flags & MAXSTYLE_MIX_STYLE) && (!perm || !lmda)) return MAXSTYLE_ERR_BAD_ARG;
    if (!(flags & MAXSTYLE_NO_NOISE) && (!gamma_noise || !beta_noise)) return MAXSTYLE_ERR_BAD_ARG;
    if (is_nhwc(layout, C)) return MAXSTYLE_ERR_UNSUPPORTED;
    const int64_t M = (int64_t)H * W;
    const Workspace w = workspace_layout(N, C, M, dtype);
    if ((rc = check_workspace(workspace, workspace_bytes, w))) return rc;
    const int sms = sm_count();
    if (sms <= 0) return MAXSTYLE_ERR_NO_DEVICE;
    PeerTables pt;
    pt.peers = reinterpret_cast<const unsigned long long*>(peers);
    pt.rank = rank; pt.world = world; pt.epoch = epoch; pt.done = nullptr;
    pt.error = reinterpret_cast<int*>(static_cast<char*>(workspace) + w.res_error);
    FwdCall f{x, y, mu_all, sig_all, perm, lmda, gamma_noise, beta_noise, gamma_std, beta_std, scale, shift, N, C, M, dtype, flags, eps,
              static_cast<char*>(workspace), static_cast<cudaStream_t>(stream), N_global, row_offset, table_ld, pt, PreOp{}};
    if (!(stats_sweep & MAXSTYLE_SWEEP_NO_PAIR) && !(stats_sweep & (MAXSTYLE_SWEEP_FORCE_WINDOW | MAXSTYLE_SWEEP_FORCE_CLUSTER))) {
        rc = try_pair_fwd(f, w, sms, stats_sweep);
        if (rc >= 0) return rc;
    }
    if (N_global > kFusedMaxN) return MAXSTYLE_ERR_UNSUPPORTED;        // the window / cluster kernels keep a channel's rows in shared memory
    if (!(stats_sweep & MAXSTYLE_SWEEP_NO_CLUSTER) && (stats_sweep & MAXSTYLE_SWEEP_FORCE_CLUSTER)) {
        rc = try_cluster_fwd(f, w, sms, stats_sweep);
        if (rc >= 0) return rc;
    }
    // the L2-window kernel hides the exchange behind a fixed 32 MB of streaming: beyond 2 ranks their skew outgrows it (measured
    // slower than statistics -> exchange + tables -> apply, profiles/r01_multi.txt), so it is only taken there when forced
    if (world > 2 && !(stats_sweep & MAXSTYLE_SWEEP_FORCE_WINDOW)) return MAXSTYLE_ERR_UNSUPPORTED;
    rc = try_fused_fwd(f, w, sms, (stats_sweep & MAXSTYLE_SWEEP_FORCE_WINDOW) != 0, (stats_sweep & MAXSTYLE_SWEEP_X_KEEP) != 0);
    return rc >= 0 ? rc : MAXSTYLE_ERR_UNSUPPORTED;
}

// ---- pixel-wise cross entropy ---------------------------------------------------------------
static int ce2d_grid(int64_t P, int sms) {
    const int64_t want = (P + kCeThreads - 1) / kCeThreads, cap = (int64_t)sms * 8;
    return (int)(want < cap ? (want < 1 ? 1 : want) : cap);
}

size_t maxstyle_ce2d_workspace_bytes(int N, int C, int H, int W) {
    if (N <= 0 || C <= 0 || H <= 0 || W <= 0) return 0;
    return 256 + (size_t)148 * 8 * 2 * sizeof(float);          // counter + per-CTA partials (grid <= 8 CTAs per SM)
}

int maxstyle_ce2d_fwd(const void* logits, const int64_t* target, const float* weight, const float* mask, float* loss,
                      int N, int C, int H, int W, int dtype, int size_average, void* workspace, size_t workspace_bytes,
                      maxstyle_stream_t stream) {
    if (!logits || !target || !loss || N <= 0 || C <= 0 || H <= 0 || W <= 0) return MAXSTYLE_ERR_BAD_ARG;
    if (dtype != MAXSTYLE_F32 && dtype != MAXSTYLE_BF16) return MAXSTYLE_ERR_UNSUPPORTED;
    if (!workspace || workspace_bytes < maxstyle_ce2d_workspace_bytes(N, C, H, W) || (reinterpret_cast<uintptr_t>(workspace) & 255u))
        return MAXSTYLE_ERR_WORKSPACE;
    const int sms = sm_count();
    if (sms <= 0) return MAXSTYLE_ERR_NO_DEVICE;
    const int64_t hw = (int64_t)H * W, P = (int64_t)N * hw;
    int grid = ce2d_grid(P, sms);
    if (grid > 148 * 8 * 2) grid = 148 * 8 * 2;
    const float inv = size_average ? 1.0f / (float)P : 1.0f;
    unsigned int* counter = static_cast<unsigned int*>(workspace);
    float* partials = reinterpret_cast<float*>(static_cast<char*>(workspace) + 256);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == MAXSTYLE_F32)
        ce2d_fwd_kernel<float><<<grid, kCeThreads, 0, s>>>(static_cast<const float*>(logits), reinterpret_cast<const long long*>(target),
                                                           weight, mask, loss, partials, counter, P, hw, C, inv);
    else
        ce2d_fwd_kernel<__nv_bfloat16><<<grid, kCeThreads, 0, s>>>(static_cast<const __nv_bfloat16*>(logits),
                                                                   reinterpret_cast<const long long*>(target), weight, mask, loss,
                                                                   partials, counter, P, hw, C, inv);
    return check_launch();
}

int maxstyle_ce2d_bwd(const void* logits, const int64_t* target, const float* weight, const float* mask, const float* dloss,
                      void* dlogits, int N, int C, int H, int W, int dtype, int size_average, maxstyle_stream_t stream) {
    if (!logits || !target || !dloss || !dlogits || N <= 0 || C <= 0 || H <= 0 || W <= 0) return MAXSTYLE_ERR_BAD_ARG;
    if (dtype != MAXSTYLE_F32 && dtype != MAXSTYLE_BF16) return MAXSTYLE_ERR_UNSUPPORTED;
    const int sms = sm_count();
    if (sms <= 0) return MAXSTYLE_ERR_NO_DEVICE;
    const int64_t hw = (int64_t)H * W, P = (int64_t)N * hw;
    const int grid = ce2d_grid(P, sms);
    const float inv = size_average ? 1.0f / (float)P : 1.0f;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == MAXSTYLE_F32)
        ce2d_bwd_kernel<float><<<grid, kCeThreads, 0, s>>>(static_cast<const float*>(logits), reinterpret_cast<const long long*>(target),
                                                           weight, mask, dloss, static_cast<float*>(dlogits), P, hw, C, inv);
    else
        ce2d_bwd_kernel<__nv_bfloat16><<<grid, kCeThreads, 0, s>>>(static_cast<const __nv_bfloat16*>(logits),
                                                                   reinterpret_cast<const long long*>(target), weight, mask, dloss,
                                                                   static_cast<__nv_bfloat16*>(dlogits), P, hw, C, inv);
    return check_launch();
}

int maxstyle_ce2d_fwd_grad(const void* logits, const int64_t* labels, const void* soft_target, int soft_is_probability,
                           const float* weight, const float* mask, float* loss, void* dlogits, void* dsoft_target,
                           int N, int C, int H, int W, int dtype, int size_average, void* workspace, size_t workspace_bytes,
                           maxstyle_stream_t stream) {
    if (!logits || !loss || N <= 0 || C <= 0 || H <= 0 || W <= 0) return MAXSTYLE_ERR_BAD_ARG;
    if ((labels == nullptr) == (soft_target == nullptr)) return MAXSTYLE_ERR_BAD_ARG;      // exactly one kind of target
    if (dsoft_target && !soft_target) return MAXSTYLE_ERR_BAD_ARG;
    if (dtype != MAXSTYLE_F32 && dtype != MAXSTYLE_BF16) return MAXSTYLE_ERR_UNSUPPORTED;
    if (C > kCeMaxC) return MAXSTYLE_ERR_UNSUPPORTED;
    if (!workspace || workspace_bytes < maxstyle_ce2d_workspace_bytes(N, C, H, W) || (reinterpret_cast<uintptr_t>(workspace) & 255u))
        return MAXSTYLE_ERR_WORKSPACE;
    const int sms = sm_count();
    if (sms <= 0) return MAXSTYLE_ERR_NO_DEVICE;
    const int64_t hw = (int64_t)H * W, P = (int64_t)N * hw;
    int grid = ce2d_grid(P, sms);
    if (grid > 148 * 8 * 2) grid = 148 * 8 * 2;
    const float inv = size_average ? 1.0f / (float)P : 1.0f;
    unsigned int* counter = static_cast<unsigned int*>(workspace);
    float* partials = reinterpret_cast<float*>(static_cast<char*>(workspace) + 256);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const long long* lab = reinterpret_cast<const long long*>(labels);
    if (dtype == MAXSTYLE_F32)
        ce2d_fused_kernel<float, kCeMaxC><<<grid, kCeThreads, 0, s>>>(static_cast<const float*>(logits), lab, static_cast<const float*>(soft_target),
                                                                     soft_is_probability, weight, mask, loss, static_cast<float*>(dlogits),
                                                                     static_cast<float*>(dsoft_target), partials, counter, P, hw, C, inv);
    else
        ce2d_fused_kernel<__nv_bfloat16, kCeMaxC><<<grid, kCeThreads, 0, s>>>(
            static_cast<const __nv_bfloat16*>(logits), lab, static_cast<const __nv_bfloat16*>(soft_target), soft_is_probability, weight, mask,
            loss, static_cast<__nv_bfloat16*>(dlogits), static_cast<__nv_bfloat16*>(dsoft_target), partials, counter, P, hw, C, inv);
    return check_launch();
}

int maxstyle_ce2d_scale(void* grad, const float* scale, int64_t count, int dtype, maxstyle_stream_t stream) {
    if (!grad || !scale || count <= 0) return MAXSTYLE_ERR_BAD_ARG;
    if (dtype != MAXSTYLE_F32 && dtype != MAXSTYLE_BF16) return MAXSTYLE_ERR_UNSUPPORTED;
    const int sms = sm_count();
    if (sms <= 0) return MAXSTYLE_ERR_NO_DEVICE;
    const int64_t want = (count + 256 * 4 - 1) / (256 * 4), cap = (int64_t)sms * 8;
    const int grid = (int)(want < cap ? want : cap);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == MAXSTYLE_F32) ce2d_scale_kernel<float><<<grid, 256, 0, s>>>(static_cast<float*>(grad), scale, count);
    else ce2d_scale_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(static_cast<__nv_bfloat16*>(grad), scale, count);
    return check_launch();
}

int maxstyle_fwd_kernels(int N, int C, int H, int W, int dtype, int layout, int stats_sweep) {
    if (check_shape(N, C, H, W, dtype, layout) != MAXSTYLE_OK) return 0;
    if ((stats_sweep & MAXSTYLE_SWEEP_NO_FUSED) || is_nhwc(layout, C)) return 3;
    const int64_t M = (int64_t)H * W;
    const int force_other = stats_sweep & (MAXSTYLE_SWEEP_FORCE_RESIDENT | MAXSTYLE_SWEEP_FORCE_RING | MAXSTYLE_SWEEP_FORCE_WINDOW);
    const int force_any = force_other | (stats_sweep & (MAXSTYLE_SWEEP_FORCE_CLUSTER | MAXSTYLE_SWEEP_FORCE_PAIR));
    auto pair_ok = [&]() {
        const int64_t pair_enabled = tunables().pair_enabled;
        if (N < 2 || (!(stats_sweep & MAXSTYLE_SWEEP_FORCE_PAIR) && pair_enabled != 1)) return false;
        const PairChoice ch = choose_pair(N, C, M, dtype, 32, sm_count(), stats_sweep, false, true, N);
        return ch.plan.ok && ch.plan.pieces <= workspace_layout(N, C, M, dtype).max_pieces;
    };
    if ((stats_sweep & MAXSTYLE_SWEEP_FORCE_PAIR) && pair_ok()) return 1;
    if (!(stats_sweep & MAXSTYLE_SWEEP_NO_CLUSTER) && (stats_sweep & MAXSTYLE_SWEEP_FORCE_CLUSTER) && N >= 2 &&
        M * elem_size(dtype) >= kClusterMinPlaneBytes) {
        const ClusterNeeds need{N, false, true};
        const ClusterChoice ch = dtype == MAXSTYLE_F32 ? choose_cluster<float>((int64_t)N * C, M, dtype, 32, sm_count(), stats_sweep, need)
                                                       : choose_cluster<__nv_bfloat16>((int64_t)N * C, M, dtype, 32, sm_count(), stats_sweep, need);
        if (ch.plan.ok && ch.clusters > 0) return 1;
    }
    if (!(stats_sweep & MAXSTYLE_SWEEP_NO_RESIDENT)) {
        const ResidentPlan rp = make_resident_plan(N, C, M, dtype, 32);
        if (rp.ok && (rp.preferred || (stats_sweep & MAXSTYLE_SWEEP_FORCE_RESIDENT))) {   // same grid rule as launch_resident
            int k = 0;
            if (dtype == MAXSTYLE_F32)
                k = rp.threads == 512 ? resident_ctas_per_sm(fwd_resident_kernel<float, 512, 1>, 512, rp.smem)
                                      : resident_ctas_per_sm(fwd_resident_kernel<float, 256, 4>, 256, rp.smem);
            else
                k = rp.threads == 512 ? resident_ctas_per_sm(fwd_resident_kernel<__nv_bfloat16, 512, 1>, 512, rp.smem)
                                      : resident_ctas_per_sm(fwd_resident_kernel<__nv_bfloat16, 256, 4>, 256, rp.smem);
            const int64_t cap = (int64_t)sm_count() * k, items = (int64_t)N * C;
            if (k > 0 && N <= (items < cap ? items : cap)) return 1;
        }
    }
    if (!(stats_sweep & MAXSTYLE_SWEEP_NO_PAIR) && !force_any && pair_preferred(N, C, M, dtype) && pair_ok()) return 1;
    if (!(stats_sweep & MAXSTYLE_SWEEP_NO_RING)) {
        const RingPlan gp = make_ring_plan(N, C, M, dtype, 32);
        if (gp.ok && (gp.profitable || (stats_sweep & MAXSTYLE_SWEEP_FORCE_RING))) return 1;
    }
    const FusedPlan fp = make_fused_plan(N, C, M, dtype, 32);
    return fp.ok && (fp.profitable || (stats_sweep & MAXSTYLE_SWEEP_FORCE_WINDOW)) ? 1 : 3;
}

int maxstyle_fwd_geometry(int N, int C, int H, int W, int dtype, int stats_sweep, int* out) {
    if (!out || check_shape(N, C, H, W, dtype, MAXSTYLE_NCHW) != MAXSTYLE_OK) return MAXSTYLE_ERR_BAD_ARG;
    const int sms = sm_count();
    if (sms <= 0) return MAXSTYLE_ERR_NO_DEVICE;
    const int64_t M = (int64_t)H * W;
    for (int i = 0; i < 12; ++i) out[i] = 0;
    if (!(stats_sweep & MAXSTYLE_SWEEP_FORCE_CLUSTER)) {
        const PairChoice pc = choose_pair(N, C, M, dtype, 32, sms, stats_sweep, false, true, N);
        if (!pc.plan.ok) return MAXSTYLE_ERR_UNSUPPORTED;
        out[2] = pc.grid; out[3] = pc.plan.piece_vecs * pc.plan.vec * elem_size(dtype); out[7] = sms; out[8] = pc.plan.pieces; out[9] = pc.use_order;
        out[10] = 1;
        return MAXSTYLE_OK;
    }
    const ClusterNeeds need{N, false, true};
    const ClusterChoice ch = dtype == MAXSTYLE_F32 ? choose_cluster<float>((int64_t)N * C, M, dtype, 32, sms, stats_sweep, need)
                                                   : choose_cluster<__nv_bfloat16>((int64_t)N * C, M, dtype, 32, sms, stats_sweep, need);
    if (!ch.plan.ok || ch.clusters <= 0) return MAXSTYLE_ERR_UNSUPPORTED;
    out[0] = ch.plan.cluster; out[1] = ch.plan.stages; out[2] = ch.clusters; out[3] = ch.plan.part_bytes;
    out[4] = ch.plan.chunk_bytes; out[5] = ch.plan.chunks; out[6] = ch.plan.smem; out[7] = sms; out[8] = ch.plan.pieces; out[9] = ch.use_order;
    return MAXSTYLE_OK;
}

int maxstyle_workspace_status(const void* workspace, size_t workspace_bytes, int N, int C, int H, int W, int dtype, int layout,
                              maxstyle_stream_t stream) {
    const int rc = check_shape(N, C, H, W, dtype, layout);
    if (rc) return rc;
    const Workspace w = workspace_layout(N, C, (int64_t)H * W, dtype, is_nhwc(layout, C));
    if (workspace == nullptr || workspace_bytes < w.total) return MAXSTYLE_ERR_WORKSPACE;
    int flag = 0;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (cudaMemcpyAsync(&flag, static_cast<const char*>(workspace) + w.res_error, sizeof(int), cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaStreamSynchronize(s) != cudaSuccess) {
        cudaGetLastError();
        return MAXSTYLE_ERR_CUDA;
    }
    return flag ? MAXSTYLE_ERR_TIMEOUT : MAXSTYLE_OK;
}

}  // extern "C"

namespace {
int bwd_impl(const void* dy, const void* x, void* dx, const float* mu_all, const float* sig_all, int table_ld, int N_global,
             int row_offset, const float* scale, const int64_t* perm, const float* lmda, const float* gamma_std,
             const float* beta_std, int flags, float* d_gamma, float* d_beta, float* d_lmda, const maxstyle_step_t* step,
             int N, int C, int H, int W, int dtype, int layout, int sweep, void* workspace, size_t workspace_bytes,
             maxstyle_stream_t stream, const PreOp& pre) {
    int rc = check_shape(N, C, H, W, dtype, layout);
    if (rc) return rc;
    if (!dy || !x || !mu_all || !sig_all || !scale) return MAXSTYLE_ERR_BAD_ARG;
    if (N_global < N || row_offset < 0 || row_offset + N > N_global || table_ld < C) return MAXSTYLE_ERR_BAD_ARG;
    if ((flags & MAXSTYLE_MIX_STYLE) && (!perm || !lmda)) return MAXSTYLE_ERR_BAD_ARG;
    if (!(flags & MAXSTYLE_NO_NOISE) && (!gamma_std || !beta_std)) return MAXSTYLE_ERR_BAD_ARG;
    if ((rc = check_step(step))) return rc;
    const int64_t M = (int64_t)H * W;
    const Workspace w = workspace_layout(N, C, M, dtype, is_nhwc(layout, C));
    if ((rc = check_workspace(workspace, workspace_bytes, w))) return rc;
    const int sms = sm_count();
    if (sms <= 0) return MAXSTYLE_ERR_NO_DEVICE;
    BwdTables tb;
    tb.mu_all = mu_all; tb.sig_all = sig_all; tb.scale = scale; tb.perm = perm; tb.lmda = lmda;
    tb.gamma_std = gamma_std; tb.beta_std = beta_std; tb.d_gamma = d_gamma; tb.d_beta = d_beta; tb.d_lmda = d_lmda;
    tb.row_offset = row_offset; tb.N = N; tb.C = C; tb.flags = flags; tb.ld = table_ld;
    const StepArgs st = to_step_args(step);
    if (is_nhwc(layout, C)) {
        if (pre.op != kPreNone) return MAXSTYLE_ERR_UNSUPPORTED;
        const PlanNhwc pn = make_plan_nhwc(N, C, M, dtype, common_align(dy, x, dx), sms);
        if (!pn.ok) return MAXSTYLE_ERR_UNSUPPORTED;
        MS_DISPATCH_NHWC(launch_bwd_nhwc, dtype, pn, dy, x, dx, static_cast<char*>(workspace), w, pn, N, M, tb, st, sweep,
                         static_cast<cudaStream_t>(stream));
        return check_launch();
    }
    const Plan p = make_plan(N, C, M, dtype, common_align(dy, x, dx), sms, (int)tunables().bwd_oversub);
    MS_DISPATCH(launch_bwd, dtype, p, dy, x, dx, static_cast<char*>(workspace), w, p, M, tb, st, sweep,
                static_cast<cudaStream_t>(stream), pre);
    return check_launch();
}
}  // namespace

extern "C" {

int maxstyle_bwd(const void* dy, const void* x, void* dx, const float* mu_all, const float* sig_all, int table_ld, int N_global,
                 int row_offset, const float* scale, const int64_t* perm, const float* lmda, const float* gamma_std,
                 const float* beta_std, int flags, float* d_gamma, float* d_beta, float* d_lmda, const maxstyle_step_t* step,
                 int N, int C, int H, int W, int dtype, int layout, int sweep, void* workspace, size_t workspace_bytes,
                 maxstyle_stream_t stream) {
    return bwd_impl(dy, x, dx, mu_all, sig_all, table_ld, N_global, row_offset, scale, perm, lmda, gamma_std, beta_std, flags, d_gamma,
                    d_beta, d_lmda, step, N, C, H, W, dtype, layout, sweep, workspace, workspace_bytes, stream, PreOp{});
}

int maxstyle_bwd_act(const void* dy, const void* z, void* dz, const float* mu_all, const float* sig_all, int table_ld, int N_global,
                     int row_offset, const float* scale, const int64_t* perm, const float* lmda, const float* gamma_std,
                     const float* beta_std, int flags, float* d_gamma, float* d_beta, float* d_lmda, const maxstyle_step_t* step,
                     int N, int C, int H, int W, int dtype, int layout, int sweep, int pre_op, float pre_param,
                     void* workspace, size_t workspace_bytes, maxstyle_stream_t stream) {
    if (pre_op != MAXSTYLE_PRE_NONE && pre_op != MAXSTYLE_PRE_LEAKY_RELU && pre_op != MAXSTYLE_PRE_SIGMOID) return MAXSTYLE_ERR_BAD_ARG;
    PreOp pre;
    pre.op = pre_op; pre.param = pre_param;
    return bwd_impl(dy, z, dz, mu_all, sig_all, table_ld, N_global, row_offset, scale, perm, lmda, gamma_std, beta_std, flags, d_gamma,
                    d_beta, d_lmda, step, N, C, H, W, dtype, layout, sweep, workspace, workspace_bytes, stream, pre);
}

int maxstyle_rescale(const void* y, const uint32_t* y_min, const uint32_t* y_max, void* out, float new_min, float new_max, float eps,
                     int N, int C, int H, int W, int dtype, maxstyle_stream_t stream) {
    if (!y || !y_min || !y_max || !out || N <= 0 || C <= 0 || H <= 0 || W <= 0) return MAXSTYLE_ERR_BAD_ARG;
    if (dtype != MAXSTYLE_F32 && dtype != MAXSTYLE_BF16) return MAXSTYLE_ERR_UNSUPPORTED;
    const int sms = sm_count();
    if (sms <= 0) return MAXSTYLE_ERR_NO_DEVICE;
    const int64_t M = (int64_t)H * W, planes = (int64_t)N * C;
    const int64_t want = planes * ((M + 1023) / 1024), cap = (int64_t)sms * 8;
    const int grid = (int)(want < cap ? want : cap);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == MAXSTYLE_F32)
        rescale_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(y), y_min, y_max, static_cast<float*>(out), new_min, new_max, eps, planes, M);
    else
        rescale_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(y), y_min, y_max, static_cast<__nv_bfloat16*>(out),
                                                           new_min, new_max, eps, planes, M);
    return check_launch();
}

int maxstyle_step(const float* d_gamma, const float* d_beta, const float* d_lmda, const maxstyle_step_t* step, int N, int C,
                  maxstyle_stream_t stream) {
    if (N <= 0 || C <= 0 || step == nullptr || step->mode == MAXSTYLE_STEP_NONE) return MAXSTYLE_ERR_BAD_ARG;
    int rc = check_step(step);
    if (rc) return rc;
    const StepArgs st = to_step_args(step);
    const int64_t total = (int64_t)N * C;
    const int blocks = (int)((total + 255) / 256);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    step_kernel<<<blocks, 256, 0, s>>>(d_gamma, d_beta, d_lmda, N, C, st);
    if (st.mode == MAXSTYLE_STEP_ADAM && st.step_dev) step_count_kernel<<<1, 1, 0, s>>>(st.step_dev);
    return check_launch();
}

}  // extern "C"
