// Device-side building blocks shared by the MaxStyle kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace ms {

constexpr int kThreads = 256;          // threads per CTA for every streaming kernel
constexpr int kWarps = kThreads / 32;

// ----------------------------------------------------------------------------------------
// Vectorised global access.  sm_100 adds 256-bit global loads/stores (SASS LDG.E.256 /
// STG.E.256).  Every access carries a run-time L2 cache policy (createpolicy + .L2::cache_hint):
// the host decides per call whether a tensor is streamed (evict-first) or will be re-read by
// the next kernel of the layer (evict-last), see "sweep flags" in include/maxstyle_b200.h.
// Vec<T,VEC> moves VEC elements of T per instruction:
//   32 bytes (8 x f32 / 16 x bf16)  -- the fast path, planes 32-byte aligned
//   16 bytes (4 x f32 /  8 x bf16)  -- planes only 16-byte aligned
//   1 element                       -- ragged / unaligned planes
// and converts to/from fp32 registers.
// ----------------------------------------------------------------------------------------
enum : int { kPolicyNormal = 0, kPolicyStream = 1, kPolicyKeep = 2 };

__device__ __forceinline__ uint64_t make_policy(int kind) {
    uint64_t p;
    if (kind == kPolicyStream) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    else if (kind == kPolicyKeep) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    return p;
}

struct Words8 { uint32_t w[8]; };
struct Words4 { uint32_t w[4]; };

__device__ __forceinline__ Words8 ld256(const void* p, uint64_t pol) {
    Words8 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
                 : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]),
                   "=r"(r.w[7]) : "l"(p), "l"(pol));
    return r;
}

__device__ __forceinline__ void st256(void* p, const Words8& r, uint64_t pol) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v8.b32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7}, %9;"
                 ::"r"(r.w[0]), "r"(r.w[1]), "r"(r.w[2]), "r"(r.w[3]), "r"(r.w[4]), "r"(r.w[5]), "r"(r.w[6]), "r"(r.w[7]),
                   "l"(p), "l"(pol) : "memory");
}

__device__ __forceinline__ Words4 ld128(const void* p, uint64_t pol) {
    Words4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.b32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]) : "l"(p), "l"(pol));
    return r;
}

__device__ __forceinline__ void st128(void* p, const Words4& r, uint64_t pol) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.b32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "r"(r.w[0]),
                 "r"(r.w[1]), "r"(r.w[2]), "r"(r.w[3]), "l"(pol) : "memory");
}

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// words <-> fp32 registers
template <typename T, int NW> struct Unpack;
template <int NW> struct Unpack<float, NW> {
    static __device__ __forceinline__ void to(const uint32_t (&w)[NW], float (&v)[NW]) {
#pragma unroll
        for (int i = 0; i < NW; ++i) v[i] = __uint_as_float(w[i]);
    }
    static __device__ __forceinline__ void from(const float (&v)[NW], uint32_t (&w)[NW]) {
#pragma unroll
        for (int i = 0; i < NW; ++i) w[i] = __float_as_uint(v[i]);
    }
};
template <int NW> struct Unpack<__nv_bfloat16, NW> {
    static __device__ __forceinline__ void to(const uint32_t (&w)[NW], float (&v)[2 * NW]) {
#pragma unroll
        for (int i = 0; i < NW; ++i) {           // bf16 -> fp32 is a 16-bit shift
            v[2 * i] = __uint_as_float(w[i] << 16);
            v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
    static __device__ __forceinline__ void from(const float (&v)[2 * NW], uint32_t (&w)[NW]) {
#pragma unroll
        for (int i = 0; i < NW; ++i) {
            __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
            w[i] = *reinterpret_cast<uint32_t*>(&h);
        }
    }
};

template <typename T, int VEC> struct Vec {
    static constexpr int kBytes = VEC * (int)sizeof(T);
    static_assert(kBytes == 32 || kBytes == 16 || VEC == 1, "unsupported vector width");
    static __device__ __forceinline__ void load(const T* p, float (&v)[VEC], uint64_t pol) {
        if constexpr (VEC == 1) {
            v[0] = to_f32<T>(__ldg(p));
        } else if constexpr (kBytes == 32) {
            const Words8 r = ld256(p, pol);
            Unpack<T, 8>::to(r.w, v);
        } else {
            const Words4 r = ld128(p, pol);
            Unpack<T, 4>::to(r.w, v);
        }
    }
    static __device__ __forceinline__ void store(T* p, const float (&v)[VEC], uint64_t pol) {
        if constexpr (VEC == 1) {
            *p = from_f32<T>(v[0]);
        } else if constexpr (kBytes == 32) {
            Words8 r;
            Unpack<T, 8>::from(v, r.w);
            st256(p, r, pol);
        } else {
            Words4 r;
            Unpack<T, 4>::from(v, r.w);
            st128(p, r, pol);
        }
    }
};

// ----------------------------------------------------------------------------------------
// Running moments (count, mean, M2 = sum of squared deviations) and their pairwise merge
// (Chan et al.), the parallel form of Welford's update.
// ----------------------------------------------------------------------------------------
struct Moments {
    float n, mean, m2;
};

__device__ __forceinline__ Moments merge(const Moments a, const Moments b) {
    const float n = a.n + b.n;
    if (n == 0.f) return a;
    const float w = b.n / n;                  // IEEE division: these merges are off the per-element path
    const float d = b.mean - a.mean;
    Moments r;
    r.n = n;
    r.mean = fmaf(d, w, a.mean);
    r.m2 = a.m2 + b.m2 + d * d * a.n * w;
    return r;
}

// Per-thread running merge on the per-element path: the weight b.n/n only has to be a consistent
// approximation (both uses see the same w), so the fast reciprocal is enough.
__device__ __forceinline__ Moments merge_fast(const Moments a, const Moments b) {
    const float n = a.n + b.n;
    if (b.n == 0.f) return a;
    const float w = __fdividef(b.n, n);
    const float d = b.mean - a.mean;
    Moments r;
    r.n = n;
    r.mean = fmaf(d, w, a.mean);
    r.m2 = a.m2 + b.m2 + d * d * a.n * w;
    return r;
}

__device__ __forceinline__ Moments shfl_xor(const Moments m, int lane_mask) {
    Moments r;
    r.n = __shfl_xor_sync(0xffffffffu, m.n, lane_mask);
    r.mean = __shfl_xor_sync(0xffffffffu, m.mean, lane_mask);
    r.m2 = __shfl_xor_sync(0xffffffffu, m.m2, lane_mask);
    return r;
}

__device__ __forceinline__ Moments warp_merge(Moments m) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = merge(m, shfl_xor(m, o));
    return m;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Group = the G threads that cooperate on one work item: a warp (G == 32) or the CTA (G == 256).
// The result is valid in every thread of the group.  `scratch` is per-CTA shared memory.
struct Scratch {
    float a[kWarps], b[kWarps], c[kWarps];
    int flag;
};

template <int G> __device__ __forceinline__ Moments group_merge(Moments m, Scratch& s) {
    m = warp_merge(m);
    if constexpr (G > 32) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        __syncthreads();                                  // scratch may still be read from the previous use
        if (lane == 0) { s.a[warp] = m.n; s.b[warp] = m.mean; s.c[warp] = m.m2; }
        __syncthreads();
        Moments t;
        t.n = lane < kWarps ? s.a[lane] : 0.f;
        t.mean = lane < kWarps ? s.b[lane] : 0.f;
        t.m2 = lane < kWarps ? s.c[lane] : 0.f;
#pragma unroll
        for (int o = kWarps / 2; o > 0; o >>= 1) t = merge(t, shfl_xor(t, o));
        m.n = __shfl_sync(0xffffffffu, t.n, 0);
        m.mean = __shfl_sync(0xffffffffu, t.mean, 0);
        m.m2 = __shfl_sync(0xffffffffu, t.m2, 0);
    }
    return m;
}

template <int G> __device__ __forceinline__ void group_sum2(float& u, float& v, Scratch& s) {
    u = warp_sum(u);
    v = warp_sum(v);
    if constexpr (G > 32) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        __syncthreads();
        if (lane == 0) { s.a[warp] = u; s.b[warp] = v; }
        __syncthreads();
        float tu = lane < kWarps ? s.a[lane] : 0.f;
        float tv = lane < kWarps ? s.b[lane] : 0.f;
#pragma unroll
        for (int o = kWarps / 2; o > 0; o >>= 1) {
            tu += __shfl_xor_sync(0xffffffffu, tu, o);
            tv += __shfl_xor_sync(0xffffffffu, tv, o);
        }
        u = __shfl_sync(0xffffffffu, tu, 0);
        v = __shfl_sync(0xffffffffu, tv, 0);
    }
}

template <int G> __device__ __forceinline__ void group_sync() {
    if constexpr (G > 32) __syncthreads(); else __syncwarp();
}

// Weighted "last arriver" ticket (called by ONE thread of a group): after publishing its partial
// result the group adds the amount of work it finished to `counter`; whoever brings the counter
// to `total` owns the merge.  The winner resets the counter, so the workspace is left zeroed
// for the next call (see include/maxstyle_b200.h).
__device__ __forceinline__ bool ticket_add(unsigned long long* counter, unsigned long long amount,
                                           unsigned long long total) {
    __threadfence();                                       // release: the partial is visible before the ticket
    const unsigned long long prev = atomicAdd(counter, amount);
    const bool last = (prev + amount == total);
    if (last) { *counter = 0ull; __threadfence(); }        // reset + acquire side
    return last;
}

// A bounded device-side wait that gave up means a broken launch (a peer rank died or is more than the bound late, the grid
// is not co-resident): carrying on would produce wrong statistics silently.  Raise the caller's error word (readable through
// maxstyle_workspace_status / PeerTableExchange.check) and trap -- the CUDA error surfaces at the host's next call, the way a
// hung NCCL collective ends in an error instead of a wrong answer.
__device__ __forceinline__ void wait_timed_out(int* error) {
    if (error != nullptr) *error = 1;
    __threadfence_system();
    __trap();
}

// ----------------------------------------------------------------------------------------
// Neighbour fusion.  The layer's input is, in the reference's decoders, the output of an activation that
// is a kernel of its own there: LeakyReLU(0.2) at the end of res_up_family (encoder_decoder.py:337-357) in front of layers
// 0-4, the sigmoid `last_act` (encoder_decoder.py:624-627) in front of layer 5.  With a pre-op the kernels take the
// PRE-activation tensor z and apply the activation to every value as it is loaded (x = act(z)); the backward returns dz.
// ----------------------------------------------------------------------------------------
enum : int { kPreNone = 0, kPreLeakyRelu = 1, kPreSigmoid = 2 };

__device__ __forceinline__ float pre_apply(float z, int op, float param) {
    if (op == kPreLeakyRelu) return z > 0.f ? z : z * param;
    if (op == kPreSigmoid) return 1.0f / (1.0f + expf(-z));                     // torch.sigmoid's formula
    return z;
}
// d act / d z expressed through x = act(z) (the value the kernels hold anyway): leaky relu keeps the sign, sigmoid' = x(1-x)
__device__ __forceinline__ float pre_grad(float x, int op, float param) {
    if (op == kPreLeakyRelu) return x > 0.f ? 1.f : param;
    if (op == kPreSigmoid) return x * (1.f - x);
    return 1.f;
}
template <int N> __device__ __forceinline__ void pre_apply_vec(float (&v)[N], int op, float param) {
    if (op != kPreNone) {
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = pre_apply(v[i], op, param);
    }
}

// Per-plane min / max of the output (the reference's rescale_intensity that follows layer 5, basic_operations.py:257-281, needs
// them): floats mapped to unsigned integers of the same order, so atomicMin / atomicMax do the reduction.  The caller
// initialises the arrays to 0xffffffff (min) and 0 (max).
__device__ __forceinline__ unsigned int ordered_bits(float f) {
    const unsigned int b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ordered_float(unsigned int u) {
    return __uint_as_float((u & 0x80000000u) ? (u ^ 0x80000000u) : ~u);
}
__device__ __forceinline__ void warp_minmax_publish(float mn, float mx, unsigned int* ymin, unsigned int* ymax) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0 && mn <= mx) {
        atomicMin(ymin, ordered_bits(mn));
        atomicMax(ymax, ordered_bits(mx));
    }
}

// Broadcast a flag computed by the group's first thread to the whole group.
template <int G> __device__ __forceinline__ bool group_bcast(bool flag, Scratch& s) {
    if constexpr (G > 32) {
        __syncthreads();
        if (threadIdx.x == 0) s.flag = flag;
        __syncthreads();
        return s.flag != 0;
    } else {
        return __shfl_sync(0xffffffffu, (int)flag, 0) != 0;
    }
}

}  // namespace ms
