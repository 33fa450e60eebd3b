// BatchNorm (train/eval) + activation + Dropout3d mask, forward and backward, on NDHWC matrices.
// Replaces nn.BatchNorm3d / BatchNorm1d, nn.Hardswish / ReLU / LeakyReLU and nn.Dropout3d call sites
// (mobilenet.py:80-82,90-92,142-143,180-181,247-248; movinet.py:65,75-76,93,141-143,150-152).
//
// All row-streaming kernels share one thread layout: a thread owns 8 fixed channels (so the per-channel
// constants live in registers for the whole kernel) and walks over rows with a CTA-wide stride; the
// activation is a template parameter, index math is 32-bit.
#include <algorithm>

#include "reduce.cuh"

namespace pb {

__global__ void bn_finalize_kernel(const double* __restrict__ sums, double M, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ rmean,
                                   float* __restrict__ rvar, int training, float momentum, float eps,
                                   float* __restrict__ scale, float* __restrict__ shift,
                                   float* __restrict__ mean_o, float* __restrict__ invstd_o,
                                   long long* __restrict__ nbt, int C) {
    pdl_trigger();
    pdl_wait();
    // momentum < 0: nn.BatchNorm(momentum=None), the cumulative moving average with factor 1/num_batches_tracked
    // (after the increment).  That mode is launched as ONE block so that every thread reads the old counter
    // before thread 0 advances it.
    if (momentum < 0.f) {
        const long long seen = nbt ? *nbt : 0;
        momentum = 1.f / (float)(seen + 1);
        __syncthreads();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && nbt && training) *nbt += 1;           // nn.BatchNorm's num_batches_tracked
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gridDim.x * blockDim.x) {
    float mean, var;
    if (training) {
        double s0 = 0, s1 = 0;
        for (int r = 0; r < PB_STAT_REPLICAS; ++r) { s0 += sums[r * 2 * C + c]; s1 += sums[r * 2 * C + C + c]; }
        double mu = s0 / M;
        double v = s1 / M - mu * mu;
        if (v < 0) v = 0;
        mean = (float)mu; var = (float)v;
        if (rmean) {
            double unbiased = M > 1 ? v * (M / (M - 1.0)) : v;
            rmean[c] = (1.f - momentum) * rmean[c] + momentum * mean;
            rvar[c]  = (1.f - momentum) * rvar[c] + momentum * (float)unbiased;
        }
    } else {
        mean = rmean[c]; var = rvar[c];
    }
    float invstd = rsqrtf(var + eps);
    // rsqrtf is 2 ulp; one Newton step brings it to fp32 round-off (matters for the 1e-4 parity mode)
    invstd = invstd * (1.5f - 0.5f * (var + eps) * invstd * invstd);
    float g = gamma ? gamma[c] : 1.f;
    float bta = beta ? beta[c] : 0.f;
    float sc = g * invstd;
    scale[c] = sc;
    shift[c] = bta - mean * sc;
    if (mean_o) mean_o[c] = mean;
    if (invstd_o) invstd_o[c] = invstd;
    }
}

template <int ACT> __device__ __forceinline__ float actf(float u, float slope) { return act_fwd(u, ACT, slope); }
template <int ACT> __device__ __forceinline__ float actg(float u, float slope) { return act_grad(u, ACT, slope); }

struct RowLayout {
    int G, RPI, g, rr, c0;
    bool active;
    __device__ __forceinline__ RowLayout(int C) {
        G = C >> 3; RPI = blockDim.x / G;
        g = threadIdx.x % G; rr = threadIdx.x / G; c0 = g << 3;
        active = rr < RPI;
    }
};

__device__ __forceinline__ void load_vec8(const float* p, float (&v)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// out = act(z*scale + shift) * mask[b][c]
template <typename T, int ACT>
__global__ void __launch_bounds__(256)
bn_act_fwd_kernel(const T* __restrict__ z, const float* __restrict__ scale, const float* __restrict__ shift,
                  const float* __restrict__ mask, T* __restrict__ out, unsigned M, unsigned R, int C, float slope) {
    pdl_trigger();
    pdl_wait();
    const RowLayout L(C);
    if (!L.active) return;
    float sc[8], sh[8];
    load_vec8(scale + L.c0, sc);
    load_vec8(shift + L.c0, sh);
    const unsigned stride = gridDim.x * L.RPI;
    auto finish = [&](unsigned m, F8 v) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v.v[i] = actf<ACT>(fmaf(v.v[i], sc[i], sh[i]), slope);
        if (mask) {
            float mk[8];
            load_vec8(mask + (size_t)(m / R) * C + L.c0, mk);
#pragma unroll
            for (int i = 0; i < 8; ++i) v.v[i] *= mk[i];
        }
        store8(out + (size_t)m * C + L.c0, v);
    };
    unsigned m = blockIdx.x * L.RPI + L.rr;
    for (; m + stride < M; m += 2 * stride) {               // two rows in flight per thread
        const F8 v0 = load8(z + (size_t)m * C + L.c0);
        const F8 v1 = load8(z + (size_t)(m + stride) * C + L.c0);
        finish(m, v0);
        finish(m + stride, v1);
    }
    if (m < M) finish(m, load8(z + (size_t)m * C + L.c0));
}

// du = dout * mask * act'(u), u = z*scale + shift, xhat = z*a + bb with a = invstd, bb = -mean*invstd
template <typename T, int ACT, bool BCAST>
__device__ __forceinline__ void bn_bwd_elem(const void* dout, const T* z, const float* mask, unsigned m, unsigned R,
                                            int C, int c0, const float (&sc)[8], const float (&sh)[8],
                                            const float (&a)[8], const float (&bb)[8], float slope,
                                            float (&du)[8], float (&xh)[8]) {
    const size_t o = (size_t)m * C + c0;
    const F8 zv = load8(z + o);
    F8 dv;
    const unsigned b = (BCAST || mask) ? m / R : 0;
    if (BCAST) dv = load8(reinterpret_cast<const float*>(dout) + (size_t)b * C + c0);
    else       dv = load8(reinterpret_cast<const T*>(dout) + o);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float u = fmaf(zv.v[i], sc[i], sh[i]);
        du[i] = dv.v[i] * actg<ACT>(u, slope);
        xh[i] = fmaf(zv.v[i], a[i], bb[i]);
    }
    if (mask) {
        float mk[8];
        load_vec8(mask + (size_t)b * C + c0, mk);
#pragma unroll
        for (int i = 0; i < 8; ++i) du[i] *= mk[i];
    }
}

// pass 1: sums[0][c] = sum du, sums[1][c] = sum du*xhat
template <typename T, int ACT, bool BCAST>
__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(const void* __restrict__ dout, const T* __restrict__ z, const float* __restrict__ scale,
                     const float* __restrict__ shift, const float* __restrict__ mean,
                     const float* __restrict__ invstd, const float* __restrict__ mask, double* __restrict__ sums,
                     unsigned M, unsigned R, int C, float slope) {
    pdl_trigger();
    pdl_wait();
    __shared__ float sm_fold[16 * 256];
    const RowLayout L(C);
    float s[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) s[i] = 0.f;
    if (L.active) {
        float sc[8], sh[8], a[8], bb[8];
        load_vec8(scale + L.c0, sc);
        load_vec8(shift + L.c0, sh);
        load_vec8(invstd + L.c0, a);
        load_vec8(mean + L.c0, bb);
#pragma unroll
        for (int i = 0; i < 8; ++i) bb[i] = -bb[i] * a[i];
        const unsigned stride = gridDim.x * L.RPI;
        unsigned m = blockIdx.x * L.RPI + L.rr;
        for (; m + stride < M; m += 2 * stride) {           // two rows in flight
            float d0[8], x0[8], d1[8], x1[8];
            bn_bwd_elem<T, ACT, BCAST>(dout, z, mask, m, R, C, L.c0, sc, sh, a, bb, slope, d0, x0);
            bn_bwd_elem<T, ACT, BCAST>(dout, z, mask, m + stride, R, C, L.c0, sc, sh, a, bb, slope, d1, x1);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                s[i] += d0[i] + d1[i];
                s[8 + i] = fmaf(d0[i], x0[i], fmaf(d1[i], x1[i], s[8 + i]));
            }
        }
        if (m < M) {
            float d0[8], x0[8];
            bn_bwd_elem<T, ACT, BCAST>(dout, z, mask, m, R, C, L.c0, sc, sh, a, bb, slope, d0, x0);
#pragma unroll
            for (int i = 0; i < 8; ++i) { s[i] += d0[i]; s[8 + i] = fmaf(d0[i], x0[i], s[8 + i]); }
        }
    }
    cta_fold_rows<16>(s, L.G, L.RPI, L.rr, sm_fold);
    if (L.rr == 0) {
        double* dst = sums + (size_t)(blockIdx.x % PB_STAT_REPLICAS) * 2 * C;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            atomicAdd(&dst[L.c0 + i], (double)s[i]);
            atomicAdd(&dst[C + L.c0 + i], (double)s[8 + i]);
        }
    }
}

// batch statistics: sums[0][c] = sum z, sums[1][c] = sum z^2 (four rows in flight per thread)
template <typename T>
__global__ void __launch_bounds__(256)
colstats_kernel(const T* __restrict__ z, double* __restrict__ sums, unsigned M, int C) {
    pdl_trigger();
    pdl_wait();
    __shared__ float sm_fold[16 * 256];
    const RowLayout L(C);
    float s[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) s[i] = 0.f;
    if (L.active) {
        const unsigned stride = gridDim.x * L.RPI;
        unsigned m = blockIdx.x * L.RPI + L.rr;
        for (; m + 3 * stride < M; m += 4 * stride) {
            F8 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = load8(z + (size_t)(m + u * stride) * C + L.c0);
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i) { s[i] += v[u].v[i]; s[8 + i] = fmaf(v[u].v[i], v[u].v[i], s[8 + i]); }
        }
        for (; m < M; m += stride) {
            const F8 v = load8(z + (size_t)m * C + L.c0);
#pragma unroll
            for (int i = 0; i < 8; ++i) { s[i] += v.v[i]; s[8 + i] = fmaf(v.v[i], v.v[i], s[8 + i]); }
        }
    }
    cta_fold_rows<16>(s, L.G, L.RPI, L.rr, sm_fold);
    if (L.rr == 0) {
        double* dst = sums + (size_t)(blockIdx.x % PB_STAT_REPLICAS) * 2 * C;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            atomicAdd(&dst[L.c0 + i], (double)s[i]);
            atomicAdd(&dst[C + L.c0 + i], (double)s[8 + i]);
        }
    }
}

__global__ void bn_bwd_finalize_kernel(const double* __restrict__ sums, double M, int training,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta,
                                       float* __restrict__ coef, int C) {
    pdl_trigger();
    pdl_wait();
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double s0 = 0, s1 = 0;
    for (int r = 0; r < PB_STAT_REPLICAS; ++r) { s0 += sums[r * 2 * C + c]; s1 += sums[r * 2 * C + C + c]; }
    if (dbeta) dbeta[c] = (float)s0;
    if (dgamma) dgamma[c] = (float)s1;
    coef[c] = training ? (float)(s0 / M) : 0.f;
    coef[C + c] = training ? (float)(s1 / M) : 0.f;
}

// pass 2: dz = scale * (du - coef0 - xhat*coef1)
template <typename T, int ACT, bool BCAST>
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const void* __restrict__ dout, const T* __restrict__ z, const float* __restrict__ scale,
                    const float* __restrict__ shift, const float* __restrict__ mean,
                    const float* __restrict__ invstd, const float* __restrict__ mask,
                    const float* __restrict__ coef, T* __restrict__ dz, unsigned M, unsigned R, int C, float slope) {
    pdl_trigger();
    pdl_wait();
    const RowLayout L(C);
    if (!L.active) return;
    float sc[8], sh[8], a[8], bb[8], k0[8], k1[8];
    load_vec8(scale + L.c0, sc);
    load_vec8(shift + L.c0, sh);
    load_vec8(invstd + L.c0, a);
    load_vec8(mean + L.c0, bb);
    load_vec8(coef + L.c0, k0);
    load_vec8(coef + C + L.c0, k1);
#pragma unroll
    for (int i = 0; i < 8; ++i) bb[i] = -bb[i] * a[i];
    const unsigned stride = gridDim.x * L.RPI;
    for (unsigned m = blockIdx.x * L.RPI + L.rr; m < M; m += stride) {
        float du[8], xh[8];
        bn_bwd_elem<T, ACT, BCAST>(dout, z, mask, m, R, C, L.c0, sc, sh, a, bb, slope, du, xh);
        F8 o;
#pragma unroll
        for (int i = 0; i < 8; ++i) o.v[i] = sc[i] * (du[i] - k0[i] - xh[i] * k1[i]);
        store8(dz + (size_t)m * C + L.c0, o);
    }
}

// rows_per_thread: 8 for the reductions (few atomics per CTA), 2 for the pure streaming kernels (small tensors
// then still fill the machine)
static inline int row_grid(long long M, int C, int rows_per_thread = 8) {
    int G = C >> 3, RPI = 256 / G;
    long long per_cta = (long long)RPI * rows_per_thread;
    long long max_ctas = (M + per_cta - 1) / per_cta;
    long long want = 148LL * 8;
    return (int)std::max<long long>(1, std::min(max_ctas, want));
}

}  // namespace pb

using namespace pb;

// Dispatch on the activation code; body sees the compile-time constant ACT.
#define PB_DISPATCH_ACT(act, ...)                                                      \
    do {                                                                               \
        switch (act) {                                                                 \
            case PB_ACT_NONE:     { constexpr int ACT = PB_ACT_NONE; __VA_ARGS__; } break;     \
            case PB_ACT_RELU:     { constexpr int ACT = PB_ACT_RELU; __VA_ARGS__; } break;     \
            case PB_ACT_HSWISH:   { constexpr int ACT = PB_ACT_HSWISH; __VA_ARGS__; } break;   \
            case PB_ACT_LRELU:    { constexpr int ACT = PB_ACT_LRELU; __VA_ARGS__; } break;    \
            case PB_ACT_HSIGMOID: { constexpr int ACT = PB_ACT_HSIGMOID; __VA_ARGS__; } break; \
            default: pb::set_error("unknown activation %d", (int)(act)); return PB_ERR_BAD_ARG; \
        }                                                                              \
    } while (0)

extern "C" int pb_colstats(const void* x, int dtype, long long M, int C, double* sums, pb_stream_t stream) {
    PB_REQUIRE(x && sums && M > 0 && C > 0 && C % 8 == 0 && C <= 2048, "colstats: bad args (M=%lld C=%d)", M, C);
    cudaStream_t st = (cudaStream_t)stream;
    PB_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * PB_STAT_REPLICAS * 2 * C, st));
    PB_REQUIRE(M < (1LL << 31), "colstats: too many rows");
    const int grid = row_grid(M, C);
    PB_DISPATCH_DTYPE(dtype, {
        (void)launch_pdl(colstats_kernel<T>, dim3(grid), dim3(256), 0, st, (const T*)x, sums, (unsigned)M, C);
    });
    PB_CHECK_LAUNCH("colstats");
    return PB_OK;
}

extern "C" int pb_bn_finalize(const double* sums, long long M, const float* gamma, const float* beta,
                              float* running_mean, float* running_var, int training, float momentum, float eps,
                              float* scale, float* shift, float* mean, float* invstd,
                              long long* num_batches_tracked, int C, pb_stream_t stream) {
    PB_REQUIRE(scale && shift && C > 0, "bn_finalize: bad args");
    PB_REQUIRE(training ? (sums != nullptr && M > 0) : (running_mean && running_var), "bn_finalize: missing statistics");
    const bool cumulative = momentum < 0.f;
    PB_REQUIRE(!cumulative || !training || num_batches_tracked, "bn_finalize: momentum=None needs num_batches_tracked");
    (void)launch_pdl(bn_finalize_kernel, dim3(cumulative ? 1 : ceil_div(C, 128)), dim3(cumulative ? 1024 : 128), 0,
                     (cudaStream_t)stream, sums, (double)M,
                     gamma, beta, running_mean, running_var, training, momentum, eps, scale, shift, mean, invstd,
                     num_batches_tracked, C);
    PB_CHECK_LAUNCH("bn_finalize");
    return PB_OK;
}

extern "C" int pb_bn_act_fwd(const void* z, const float* scale, const float* shift, const float* mask, void* out,
                             int dtype, int B, long long R, int C, int act, float slope, pb_stream_t stream) {
    PB_REQUIRE(z && scale && shift && out && B > 0 && R > 0 && C > 0 && C % 8 == 0 && C <= 2048, "bn_act_fwd: bad args");
    const long long M = (long long)B * R;
    PB_REQUIRE(M < (1LL << 31) && R < (1LL << 31), "bn_act_fwd: too many rows");
    const int grid = row_grid(M, C, 2);
    PB_DISPATCH_DTYPE(dtype, PB_DISPATCH_ACT(act, {
        (void)launch_pdl(bn_act_fwd_kernel<T, ACT>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const T*)z,
                         scale, shift, mask, (T*)out, (unsigned)M, (unsigned)R, C, slope);
    }));
    PB_CHECK_LAUNCH("bn_act_fwd");
    return PB_OK;
}

extern "C" int pb_bn_act_bwd_reduce(const void* dout, int dout_bcast, const void* z, const float* scale,
                                    const float* shift, const float* mean, const float* invstd, const float* mask,
                                    double* sums, int dtype, int B, long long R, int C, int act, float slope,
                                    pb_stream_t stream) {
    PB_REQUIRE(dout && z && scale && shift && mean && invstd && sums, "bn_act_bwd_reduce: null pointer");
    PB_REQUIRE(B > 0 && R > 0 && C > 0 && C % 8 == 0 && C <= 2048, "bn_act_bwd_reduce: bad dims");
    const long long M = (long long)B * R;
    PB_REQUIRE(M < (1LL << 31) && R < (1LL << 31), "bn_act_bwd_reduce: too many rows");
    cudaStream_t st = (cudaStream_t)stream;
    PB_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * PB_STAT_REPLICAS * 2 * C, st));
    const int grid = row_grid(M, C);
    const size_t smem = 0;
    PB_DISPATCH_DTYPE(dtype, PB_DISPATCH_ACT(act, {
        if (dout_bcast)
            (void)launch_pdl(bn_bwd_reduce_kernel<T, ACT, true>, dim3(grid), dim3(256), smem, st, dout, (const T*)z,
                             scale, shift, mean, invstd, mask, sums, (unsigned)M, (unsigned)R, C, slope);
        else
            (void)launch_pdl(bn_bwd_reduce_kernel<T, ACT, false>, dim3(grid), dim3(256), smem, st, dout, (const T*)z,
                             scale, shift, mean, invstd, mask, sums, (unsigned)M, (unsigned)R, C, slope);
    }));
    PB_CHECK_LAUNCH("bn_act_bwd_reduce");
    return PB_OK;
}

extern "C" int pb_bn_bwd_finalize(const double* sums, long long M, int training, float* dgamma, float* dbeta,
                                  float* coef, int C, pb_stream_t stream) {
    PB_REQUIRE(sums && coef && M > 0 && C > 0, "bn_bwd_finalize: bad args");
    (void)launch_pdl(bn_bwd_finalize_kernel, dim3(ceil_div(C, 128)), dim3(128), 0, (cudaStream_t)stream, sums,
                     (double)M, training, dgamma, dbeta, coef, C);
    PB_CHECK_LAUNCH("bn_bwd_finalize");
    return PB_OK;
}

extern "C" int pb_bn_act_bwd_apply(const void* dout, int dout_bcast, const void* z, const float* scale,
                                   const float* shift, const float* mean, const float* invstd, const float* mask,
                                   const float* coef, void* dz, int dtype, int B, long long R, int C, int act,
                                   float slope, pb_stream_t stream) {
    PB_REQUIRE(dout && z && scale && shift && mean && invstd && coef && dz, "bn_act_bwd_apply: null pointer");
    PB_REQUIRE(B > 0 && R > 0 && C > 0 && C % 8 == 0 && C <= 2048, "bn_act_bwd_apply: bad dims");
    const long long M = (long long)B * R;
    PB_REQUIRE(M < (1LL << 31) && R < (1LL << 31), "bn_act_bwd_apply: too many rows");
    const int grid = row_grid(M, C, 2);
    cudaStream_t st = (cudaStream_t)stream;
    PB_DISPATCH_DTYPE(dtype, PB_DISPATCH_ACT(act, {
        if (dout_bcast)
            (void)launch_pdl(bn_bwd_apply_kernel<T, ACT, true>, dim3(grid), dim3(256), 0, st, dout, (const T*)z,
                             scale, shift, mean, invstd, mask, coef, (T*)dz, (unsigned)M, (unsigned)R, C, slope);
        else
            (void)launch_pdl(bn_bwd_apply_kernel<T, ACT, false>, dim3(grid), dim3(256), 0, st, dout, (const T*)z,
                             scale, shift, mean, invstd, mask, coef, (T*)dz, (unsigned)M, (unsigned)R, C, slope);
    }));
    PB_CHECK_LAUNCH("bn_act_bwd_apply");
    return PB_OK;
}
