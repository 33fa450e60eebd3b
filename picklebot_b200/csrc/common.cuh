// picklebot_b200 -- shared device/host helpers for the sm_100a kernels.
//
// Activation layout everywhere: NDHWC ("channels-last-3d"), i.e. a row-major matrix
// X[M][C] with M = B*T*H*W rows and C (a multiple of 8) contiguous channels.  A "sample" is a
// run of R = T*H*W consecutive rows.  Storage type T is __nv_bfloat16 (production) or float
// (the 1e-4 parity mode); accumulation is always fp32 (fp64 for cross-CTA statistics).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "../../include/picklebot_b200.h"

namespace pb {

// ---------------------------------------------------------------------------------------------
// error plumbing (no exceptions across the C ABI)
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int  cuda_fail(cudaError_t e, const char* what);
void count_launch(int n = 1);
void count_path(int path);

#define PB_REQUIRE(cond, ...)                                   \
    do {                                                        \
        if (!(cond)) {                                          \
            pb::set_error(__VA_ARGS__);                         \
            return PB_ERR_BAD_ARG;                              \
        }                                                       \
    } while (0)

#define PB_CHECK_LAUNCH(what)                                   \
    do {                                                        \
        cudaError_t e__ = cudaGetLastError();                   \
        if (e__ != cudaSuccess) return pb::cuda_fail(e__, what);\
        pb::count_launch();                                     \
    } while (0)

#define PB_CUDA(call)                                           \
    do {                                                        \
        cudaError_t e__ = (call);                               \
        if (e__ != cudaSuccess) return pb::cuda_fail(e__, #call);\
    } while (0)

// Dispatch on the storage dtype.  Body sees `T`.
#define PB_DISPATCH_DTYPE(dtype, ...)                                                    \
    do {                                                                                 \
        if ((dtype) == PB_F32) { using T = float; __VA_ARGS__; }                         \
        else if ((dtype) == PB_BF16) { using T = __nv_bfloat16; __VA_ARGS__; }           \
        else { pb::set_error("unsupported dtype %d", (int)(dtype)); return PB_ERR_BAD_ARG; } \
    } while (0)

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch: a kernel launched through launch_pdl may be scheduled while the previous
// kernel of the stream is still draining; everything it does before pdl_wait() (barrier init, TMEM
// allocation, descriptor prefetch, per-channel constants) overlaps that tail.  pdl_wait() returns once the
// previous kernel has completed and its writes are visible, so it must precede the first global-memory
// access that can alias another kernel's output -- and every thread of every such kernel must reach it,
// because the next kernel in the chain relies on this one not finishing before its predecessor.
// pdl_trigger() (first statement) lets the next kernel's CTAs be scheduled as soon as all of ours started.
// PB_PDL=0 in the environment launches everything fully serialised (the instructions become no-ops).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute: remember which devices of this process
// already have it (bit d of `done`), so a second GPU used by the same process is configured too.
template <typename KernelT>
static inline cudaError_t ensure_dyn_smem(KernelT kernel, int bytes, unsigned long long* done) {
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    if (__atomic_load_n(done, __ATOMIC_ACQUIRE) & bit) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) __atomic_fetch_or(done, bit, __ATOMIC_RELEASE);
    return e;
}

// ---------------------------------------------------------------------------------------------
// 8-channel vectors: one 16-byte access for bf16, two for fp32
// ---------------------------------------------------------------------------------------------
struct F8 { float v[8]; };

__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&p);
}

__device__ __forceinline__ F8 load8(const __nv_bfloat16* p) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    F8 r;
    r.v[0] = bf16_lo(u.x); r.v[1] = bf16_hi(u.x);
    r.v[2] = bf16_lo(u.y); r.v[3] = bf16_hi(u.y);
    r.v[4] = bf16_lo(u.z); r.v[5] = bf16_hi(u.z);
    r.v[6] = bf16_lo(u.w); r.v[7] = bf16_hi(u.w);
    return r;
}
__device__ __forceinline__ F8 load8(const float* p) {
    float4 a = *reinterpret_cast<const float4*>(p);
    float4 b = *reinterpret_cast<const float4*>(p + 4);
    F8 r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const F8& r) {
    uint4 u;
    u.x = pack_bf16x2(r.v[0], r.v[1]);
    u.y = pack_bf16x2(r.v[2], r.v[3]);
    u.z = pack_bf16x2(r.v[4], r.v[5]);
    u.w = pack_bf16x2(r.v[6], r.v[7]);
    *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ void store8(float* p, const F8& r) {
    *reinterpret_cast<float4*>(p)     = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(r.v[4], r.v[5], r.v[6], r.v[7]);
}
__device__ __forceinline__ F8 zero8() {
    F8 r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.v[i] = 0.f;
    return r;
}

__device__ __forceinline__ float to_float(float x) { return x; }
__device__ __forceinline__ float to_float(__nv_bfloat16 x) { return __bfloat162float(x); }
template <typename T> __device__ __forceinline__ T from_float(float x);
template <> __device__ __forceinline__ float from_float<float>(float x) { return x; }
template <> __device__ __forceinline__ __nv_bfloat16 from_float<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }

// Round to the storage type and back: mimics the reference's autocast casting weights/inputs to bf16.
template <typename T> __device__ __forceinline__ float round_to(float x) { return to_float(from_float<T>(x)); }

// ---------------------------------------------------------------------------------------------
// activations (mobilenet.py:56,148-157,228-230,20): forward and derivative w.r.t. the input
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float act_fwd(float u, int act, float slope) {
    switch (act) {
        case PB_ACT_RELU:     return u > 0.f ? u : 0.f;
        case PB_ACT_HSWISH:   return u * fminf(fmaxf(u + 3.f, 0.f), 6.f) * (1.f / 6.f);
        case PB_ACT_LRELU:    return u > 0.f ? u : u * slope;
        case PB_ACT_HSIGMOID: return fminf(fmaxf(u + 3.f, 0.f), 6.f) * (1.f / 6.f);
        default:              return u;
    }
}
// activation of a register vector with a RUN-TIME code: one switch around the loop, not one per element
template <int NV>
__device__ __forceinline__ void act_fwd_vec(float (&v)[NV], int act, float slope) {
    switch (act) {
        case PB_ACT_NONE: break;
        case PB_ACT_RELU:
#pragma unroll
            for (int j = 0; j < NV; ++j) v[j] = fmaxf(v[j], 0.f);
            break;
        case PB_ACT_HSWISH:
#pragma unroll
            for (int j = 0; j < NV; ++j) v[j] = v[j] * fminf(fmaxf(v[j] + 3.f, 0.f), 6.f) * (1.f / 6.f);
            break;
        default:
#pragma unroll
            for (int j = 0; j < NV; ++j) v[j] = act_fwd(v[j], act, slope);
    }
}
__device__ __forceinline__ float act_grad(float u, int act, float slope) {
    switch (act) {
        case PB_ACT_RELU:     return u > 0.f ? 1.f : 0.f;
        case PB_ACT_HSWISH:   return u < -3.f ? 0.f : (u <= 3.f ? (u * (1.f / 3.f) + 0.5f) : 1.f);
        case PB_ACT_LRELU:    return u > 0.f ? 1.f : slope;
        case PB_ACT_HSIGMOID: return (u > -3.f && u < 3.f) ? (1.f / 6.f) : 0.f;
        default:              return 1.f;
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace pb
