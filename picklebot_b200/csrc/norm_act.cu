// BatchNorm (train/eval) + activation + Dropout3d mask, forward and backward, on NDHWC matrices.
// Replaces nn.BatchNorm3d / BatchNorm1d, nn.Hardswish / ReLU / LeakyReLU and nn.Dropout3d call sites
// (mobilenet.py:80-82,90-92,142-143,180-181,247-248; movinet.py:65,75-76,93,141-143,150-152).
#include "reduce.cuh"

namespace pb {

template <typename T>
struct StatsF {
    const T* x; long long R; int C;
    __device__ void operator()(int b, long long r, int c0, float (&out)[2][8]) const {
        F8 v = load8(x + ((long long)b * R + r) * C + c0);
#pragma unroll
        for (int i = 0; i < 8; ++i) { out[0][i] = v.v[i]; out[1][i] = v.v[i] * v.v[i]; }
    }
};

__global__ void bn_finalize_kernel(const double* __restrict__ sums, double M, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ rmean,
                                   float* __restrict__ rvar, int training, float momentum, float eps,
                                   float* __restrict__ scale, float* __restrict__ shift,
                                   float* __restrict__ mean_o, float* __restrict__ invstd_o, int C) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float mean, var;
    if (training) {
        double mu = sums[c] / M;
        double v = sums[C + c] / M - mu * mu;
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

template <typename T>
__global__ void __launch_bounds__(256)
bn_act_fwd_kernel(const T* __restrict__ z, const float* __restrict__ scale, const float* __restrict__ shift,
                  const float* __restrict__ mask, T* __restrict__ out, long long R, int C, int act, float slope,
                  long long total) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int G = C >> 3;
    int c0 = (int)(idx % G) << 3;
    long long m = idx / G;
    F8 v = load8(z + idx * 8);
    float4 s0 = __ldg(reinterpret_cast<const float4*>(scale + c0)), s1 = __ldg(reinterpret_cast<const float4*>(scale + c0 + 4));
    float4 h0 = __ldg(reinterpret_cast<const float4*>(shift + c0)), h1 = __ldg(reinterpret_cast<const float4*>(shift + c0 + 4));
    float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
    float sh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
    F8 o;
#pragma unroll
    for (int i = 0; i < 8; ++i) o.v[i] = act_fwd(fmaf(v.v[i], sc[i], sh[i]), act, slope);
    if (mask) {
        const float* mp = mask + (m / R) * C + c0;
#pragma unroll
        for (int i = 0; i < 8; ++i) o.v[i] *= __ldg(mp + i);
    }
    store8(out + idx * 8, o);
}

// du = dout * mask * act'(u), u = z*scale + shift, xhat = (z - mean) * invstd
template <typename T>
struct BnBwdCommon {
    const void* dout; int dout_bcast; const T* z;
    const float *scale, *shift, *mean, *invstd, *mask;
    long long R; int C; int act; float slope;
    __device__ __forceinline__ void eval(long long m, int c0, float (&du)[8], float (&xh)[8]) const {
        long long b = m / R;
        F8 zv = load8(z + m * C + c0);
        F8 dv;
        if (dout_bcast) dv = load8(reinterpret_cast<const float*>(dout) + b * C + c0);
        else            dv = load8(reinterpret_cast<const T*>(dout) + m * C + c0);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float u = fmaf(zv.v[i], scale[c0 + i], shift[c0 + i]);
            float g = dv.v[i] * act_grad(u, act, slope);
            if (mask) g *= mask[b * C + c0 + i];
            du[i] = g;
            xh[i] = (zv.v[i] - mean[c0 + i]) * invstd[c0 + i];
        }
    }
};

template <typename T>
struct BnBwdReduceF {
    BnBwdCommon<T> k;
    __device__ void operator()(int, long long r, int c0, float (&out)[2][8]) const {
        float du[8], xh[8];
        k.eval(r, c0, du, xh);
#pragma unroll
        for (int i = 0; i < 8; ++i) { out[0][i] = du[i]; out[1][i] = du[i] * xh[i]; }
    }
};

__global__ void bn_bwd_finalize_kernel(const double* __restrict__ sums, double M, int training,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta,
                                       float* __restrict__ coef, int C) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double s0 = sums[c], s1 = sums[C + c];
    if (dbeta) dbeta[c] = (float)s0;
    if (dgamma) dgamma[c] = (float)s1;
    coef[c] = training ? (float)(s0 / M) : 0.f;
    coef[C + c] = training ? (float)(s1 / M) : 0.f;
}

template <typename T>
__global__ void __launch_bounds__(256)
bn_act_bwd_apply_kernel(BnBwdCommon<T> k, const float* __restrict__ coef, T* __restrict__ dz, long long total) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int G = k.C >> 3;
    int c0 = (int)(idx % G) << 3;
    long long m = idx / G;
    float du[8], xh[8];
    k.eval(m, c0, du, xh);
    F8 o;
#pragma unroll
    for (int i = 0; i < 8; ++i)
        o.v[i] = k.scale[c0 + i] * (du[i] - coef[c0 + i] - xh[i] * coef[k.C + c0 + i]);
    store8(dz + idx * 8, o);
}

}  // namespace pb

using namespace pb;

extern "C" int pb_colstats(const void* x, int dtype, long long M, int C, double* sums, pb_stream_t stream) {
    PB_REQUIRE(x && sums && M > 0 && C > 0 && C % 8 == 0 && C <= 2048, "colstats: bad args (M=%lld C=%d)", M, C);
    cudaStream_t st = (cudaStream_t)stream;
    PB_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, st));
    dim3 grid = colreduce_grid(M, C, 1);
    PB_DISPATCH_DTYPE(dtype, {
        StatsF<T> f{(const T*)x, M, C};
        colreduce_kernel<StatsF<T>, 2, double><<<grid, 256, sizeof(float) * 2 * C, st>>>(f, M, C, sums, 1, 1.f);
    });
    PB_CHECK_LAUNCH("colstats");
    return PB_OK;
}

extern "C" int pb_bn_finalize(const double* sums, long long M, const float* gamma, const float* beta,
                              float* running_mean, float* running_var, int training, float momentum, float eps,
                              float* scale, float* shift, float* mean, float* invstd, int C, pb_stream_t stream) {
    PB_REQUIRE(scale && shift && C > 0, "bn_finalize: bad args");
    PB_REQUIRE(training ? (sums != nullptr && M > 0) : (running_mean && running_var), "bn_finalize: missing statistics");
    bn_finalize_kernel<<<ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(sums, (double)M, gamma, beta, running_mean,
                                                                        running_var, training, momentum, eps, scale,
                                                                        shift, mean, invstd, C);
    PB_CHECK_LAUNCH("bn_finalize");
    return PB_OK;
}

extern "C" int pb_bn_act_fwd(const void* z, const float* scale, const float* shift, const float* mask, void* out,
                             int dtype, int B, long long R, int C, int act, float slope, pb_stream_t stream) {
    PB_REQUIRE(z && scale && shift && out && B > 0 && R > 0 && C > 0 && C % 8 == 0, "bn_act_fwd: bad args");
    long long total = (long long)B * R * (C / 8);
    PB_DISPATCH_DTYPE(dtype, {
        bn_act_fwd_kernel<T><<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>((const T*)z, scale, shift, mask,
                                                                                  (T*)out, R, C, act, slope, total);
    });
    PB_CHECK_LAUNCH("bn_act_fwd");
    return PB_OK;
}

extern "C" int pb_bn_act_bwd_reduce(const void* dout, int dout_bcast, const void* z, const float* scale,
                                    const float* shift, const float* mean, const float* invstd, const float* mask,
                                    double* sums, int dtype, int B, long long R, int C, int act, float slope,
                                    pb_stream_t stream) {
    PB_REQUIRE(dout && z && scale && shift && mean && invstd && sums, "bn_act_bwd_reduce: null pointer");
    PB_REQUIRE(B > 0 && R > 0 && C > 0 && C % 8 == 0 && C <= 2048, "bn_act_bwd_reduce: bad dims");
    cudaStream_t st = (cudaStream_t)stream;
    PB_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, st));
    long long M = (long long)B * R;
    dim3 grid = colreduce_grid(M, C, 1);
    PB_DISPATCH_DTYPE(dtype, {
        BnBwdReduceF<T> f{{dout, dout_bcast, (const T*)z, scale, shift, mean, invstd, mask, R, C, act, slope}};
        colreduce_kernel<BnBwdReduceF<T>, 2, double><<<grid, 256, sizeof(float) * 2 * C, st>>>(f, M, C, sums, 1, 1.f);
    });
    PB_CHECK_LAUNCH("bn_act_bwd_reduce");
    return PB_OK;
}

extern "C" int pb_bn_bwd_finalize(const double* sums, long long M, int training, float* dgamma, float* dbeta,
                                  float* coef, int C, pb_stream_t stream) {
    PB_REQUIRE(sums && coef && M > 0 && C > 0, "bn_bwd_finalize: bad args");
    bn_bwd_finalize_kernel<<<ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(sums, (double)M, training, dgamma, dbeta, coef, C);
    PB_CHECK_LAUNCH("bn_bwd_finalize");
    return PB_OK;
}

extern "C" int pb_bn_act_bwd_apply(const void* dout, int dout_bcast, const void* z, const float* scale,
                                   const float* shift, const float* mean, const float* invstd, const float* mask,
                                   const float* coef, void* dz, int dtype, int B, long long R, int C, int act,
                                   float slope, pb_stream_t stream) {
    PB_REQUIRE(dout && z && scale && shift && mean && invstd && coef && dz, "bn_act_bwd_apply: null pointer");
    PB_REQUIRE(B > 0 && R > 0 && C > 0 && C % 8 == 0, "bn_act_bwd_apply: bad dims");
    long long total = (long long)B * R * (C / 8);
    PB_DISPATCH_DTYPE(dtype, {
        BnBwdCommon<T> k{dout, dout_bcast, (const T*)z, scale, shift, mean, invstd, mask, R, C, act, slope};
        bn_act_bwd_apply_kernel<T><<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(k, coef, (T*)dz, total);
    });
    PB_CHECK_LAUNCH("bn_act_bwd_apply");
    return PB_OK;
}
