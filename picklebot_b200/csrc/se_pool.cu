// Squeeze-excite (SEBlock3D, mobilenet.py:11-26) and global average pooling
// (AdaptiveAvgPool3d, mobilenet.py:16,186,252; movinet.py:147) on NDHWC matrices [B][R][C].
#include "reduce.cuh"

namespace pb {

template <typename T>
struct PoolF {
    const T* x; long long R; int C;
    __device__ void operator()(int b, long long r, int c0, float (&out)[1][8]) const {
        F8 v = load8(x + ((long long)b * R + r) * C + c0);
#pragma unroll
        for (int i = 0; i < 8; ++i) out[0][i] = v.v[i];
    }
};

template <typename T>
struct DotF {
    const T* g; const T* y; long long R; int C;
    __device__ void operator()(int b, long long r, int c0, float (&out)[1][8]) const {
        long long o = ((long long)b * R + r) * C + c0;
        F8 a = load8(g + o), v = load8(y + o);
#pragma unroll
        for (int i = 0; i < 8; ++i) out[0][i] = a.v[i] * v.v[i];
    }
};

// one CTA per sample
__global__ void __launch_bounds__(256)
se_fc_fwd_kernel(const float* __restrict__ mean, const float* __restrict__ W1, const float* __restrict__ b1,
                 const float* __restrict__ W2, const float* __restrict__ b2, float* __restrict__ hidden,
                 float* __restrict__ gate, int C, int Ch) {
    extern __shared__ float sm[];   // mean[C] | hidden[Ch]
    float* m_s = sm;
    float* h_s = sm + C;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
    for (int c = tid; c < C; c += blockDim.x) m_s[c] = mean[(long long)b * C + c];
    __syncthreads();
    for (int j = warp; j < Ch; j += nwarp) {
        const float* w = W1 + (long long)j * C;
        float s = 0.f;
        for (int c = lane; c < C; c += 32) s = fmaf(w[c], m_s[c], s);
        s = warp_sum(s);
        if (lane == 0) {
            float h = fmaxf(s + b1[j], 0.f);
            h_s[j] = h;
            hidden[(long long)b * Ch + j] = h;
        }
    }
    __syncthreads();
    for (int c = warp; c < C; c += nwarp) {
        const float* w = W2 + (long long)c * Ch;
        float s = 0.f;
        for (int j = lane; j < Ch; j += 32) s = fmaf(w[j], h_s[j], s);
        s = warp_sum(s);
        if (lane == 0) gate[(long long)b * C + c] = act_fwd(s + b2[c], PB_ACT_HSIGMOID, 0.f);
    }
}

// per sample: da2 = dgate * hsig'(.), da1 = (W2^T da2) * relu'(.), dmean = W1^T da1 * inv_R
// work: da2 [B][C] | da1 [B][Ch]
__global__ void __launch_bounds__(256)
se_fc_bwd_sample_kernel(const float* __restrict__ dgate, const float* __restrict__ hidden,
                        const float* __restrict__ gate, const float* __restrict__ W1,
                        const float* __restrict__ W2, float inv_R, float* __restrict__ dmean,
                        float* __restrict__ work, int B, int C, int Ch) {
    extern __shared__ float sm[];   // da2[C] | da1[Ch]
    float* a2 = sm;
    float* a1 = sm + C;
    const int b = blockIdx.x, tid = threadIdx.x;
    for (int c = tid; c < C; c += blockDim.x) {
        float gt = gate[(long long)b * C + c];
        float d = (gt > 0.f && gt < 1.f) ? dgate[(long long)b * C + c] * (1.f / 6.f) : 0.f;
        a2[c] = d;
        work[(long long)b * C + c] = d;
    }
    __syncthreads();
    for (int j = tid; j < Ch; j += blockDim.x) {
        float s = 0.f;
        for (int c = 0; c < C; ++c) s = fmaf(W2[(long long)c * Ch + j], a2[c], s);
        float d = hidden[(long long)b * Ch + j] > 0.f ? s : 0.f;
        a1[j] = d;
        work[(long long)B * C + (long long)b * Ch + j] = d;
    }
    __syncthreads();
    for (int c = tid; c < C; c += blockDim.x) {
        float s = 0.f;
        for (int j = 0; j < Ch; ++j) s = fmaf(W1[(long long)j * C + c], a1[j], s);
        dmean[(long long)b * C + c] = s * inv_R;
    }
}

// parameter gradients: sums over the B samples
__global__ void se_fc_bwd_param_kernel(const float* __restrict__ work, const float* __restrict__ mean,
                                       const float* __restrict__ hidden, float* __restrict__ dW1,
                                       float* __restrict__ db1, float* __restrict__ dW2, float* __restrict__ db2,
                                       int B, int C, int Ch) {
    const float* a2 = work;
    const float* a1 = work + (long long)B * C;
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long nW = (long long)C * Ch;
    if (idx < nW) {                       // dW2[c][j] = sum_b da2[b][c] * hidden[b][j]
        int c = (int)(idx / Ch), j = (int)(idx % Ch);
        float s = 0.f;
        for (int b = 0; b < B; ++b) s = fmaf(a2[(long long)b * C + c], hidden[(long long)b * Ch + j], s);
        dW2[idx] = s;
    } else if (idx < 2 * nW) {            // dW1[j][c] = sum_b da1[b][j] * mean[b][c]
        long long i = idx - nW;
        int j = (int)(i / C), c = (int)(i % C);
        float s = 0.f;
        for (int b = 0; b < B; ++b) s = fmaf(a1[(long long)b * Ch + j], mean[(long long)b * C + c], s);
        dW1[i] = s;
    } else if (idx < 2 * nW + C) {
        int c = (int)(idx - 2 * nW);
        float s = 0.f;
        for (int b = 0; b < B; ++b) s += a2[(long long)b * C + c];
        db2[c] = s;
    } else if (idx < 2 * nW + C + Ch) {
        int j = (int)(idx - 2 * nW - C);
        float s = 0.f;
        for (int b = 0; b < B; ++b) s += a1[(long long)b * Ch + j];
        db1[j] = s;
    }
}

template <typename T, bool ADD>
__global__ void __launch_bounds__(256)
rowscale_kernel(const T* x, const float* __restrict__ gate, const float* __restrict__ add,
                T* y, long long R, int C, long long total) {   // x may alias y (in-place)
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int G = C >> 3;
    int c0 = (int)(idx % G) << 3;
    long long b = (idx / G) / R;
    F8 v = load8(x + idx * 8);
    const float* gp = gate + b * C + c0;
#pragma unroll
    for (int i = 0; i < 8; ++i) v.v[i] *= __ldg(gp + i);
    if (ADD) {
        const float* ap = add + b * C + c0;
#pragma unroll
        for (int i = 0; i < 8; ++i) v.v[i] += __ldg(ap + i);
    }
    store8(y + idx * 8, v);
}

}  // namespace pb

using namespace pb;

extern "C" int pb_pool_fwd(const void* x, int dtype, int B, long long R, int C, float* mean, pb_stream_t stream) {
    PB_REQUIRE(x && mean && B > 0 && B <= 65535 && R > 0 && C > 0 && C % 8 == 0 && C <= 2048, "pool_fwd: bad args");
    cudaStream_t st = (cudaStream_t)stream;
    PB_CUDA(cudaMemsetAsync(mean, 0, sizeof(float) * (size_t)B * C, st));
    dim3 grid = colreduce_grid(R, C, B);
    PB_DISPATCH_DTYPE(dtype, {
        PoolF<T> f{(const T*)x, R, C};
        colreduce_kernel<PoolF<T>, 1, float><<<grid, 256, sizeof(float) * C, st>>>(f, R, C, mean, B, 1.f / (float)R);
    });
    PB_CHECK_LAUNCH("pool_fwd");
    return PB_OK;
}

extern "C" int pb_rowdot(const void* g, const void* y, int dtype, int B, long long R, int C, float* out,
                         pb_stream_t stream) {
    PB_REQUIRE(g && y && out && B > 0 && B <= 65535 && R > 0 && C > 0 && C % 8 == 0 && C <= 2048, "rowdot: bad args");
    cudaStream_t st = (cudaStream_t)stream;
    PB_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)B * C, st));
    dim3 grid = colreduce_grid(R, C, B);
    PB_DISPATCH_DTYPE(dtype, {
        DotF<T> f{(const T*)g, (const T*)y, R, C};
        colreduce_kernel<DotF<T>, 1, float><<<grid, 256, sizeof(float) * C, st>>>(f, R, C, out, B, 1.f);
    });
    PB_CHECK_LAUNCH("rowdot");
    return PB_OK;
}

extern "C" int pb_se_fc_fwd(const float* mean, const float* W1, const float* b1, const float* W2, const float* b2,
                            float* hidden, float* gate, int B, int C, int Ch, pb_stream_t stream) {
    PB_REQUIRE(mean && W1 && b1 && W2 && b2 && hidden && gate && B > 0 && C > 0 && Ch > 0, "se_fc_fwd: bad args");
    se_fc_fwd_kernel<<<B, 256, sizeof(float) * (C + Ch), (cudaStream_t)stream>>>(mean, W1, b1, W2, b2, hidden, gate, C, Ch);
    PB_CHECK_LAUNCH("se_fc_fwd");
    return PB_OK;
}

extern "C" int pb_se_fc_bwd(const float* dgate, const float* mean, const float* hidden, const float* gate,
                            const float* W1, const float* W2, float inv_R, float* dmean, float* work,
                            float* dW1, float* db1, float* dW2, float* db2, int B, int C, int Ch,
                            pb_stream_t stream) {
    PB_REQUIRE(dgate && mean && hidden && gate && W1 && W2 && dmean && work && dW1 && db1 && dW2 && db2,
               "se_fc_bwd: null pointer");
    PB_REQUIRE(B > 0 && C > 0 && Ch > 0, "se_fc_bwd: bad dims");
    cudaStream_t st = (cudaStream_t)stream;
    se_fc_bwd_sample_kernel<<<B, 256, sizeof(float) * (C + Ch), st>>>(dgate, hidden, gate, W1, W2, inv_R, dmean, work, B, C, Ch);
    PB_CHECK_LAUNCH("se_fc_bwd_sample");
    long long n = 2LL * C * Ch + C + Ch;
    se_fc_bwd_param_kernel<<<ceil_div(n, 256), 256, 0, st>>>(work, mean, hidden, dW1, db1, dW2, db2, B, C, Ch);
    PB_CHECK_LAUNCH("se_fc_bwd_param");
    return PB_OK;
}

extern "C" int pb_rowscale(const void* x, const float* gate, void* y, int dtype, int B, long long R, int C,
                           pb_stream_t stream) {
    PB_REQUIRE(x && gate && y && B > 0 && R > 0 && C > 0 && C % 8 == 0, "rowscale: bad args");
    long long total = (long long)B * R * (C / 8);
    PB_DISPATCH_DTYPE(dtype, {
        rowscale_kernel<T, false><<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>((const T*)x, gate, nullptr, (T*)y, R, C, total);
    });
    PB_CHECK_LAUNCH("rowscale");
    return PB_OK;
}

extern "C" int pb_scale_add(void* g, const float* gate, const float* add, int dtype, int B, long long R, int C,
                            pb_stream_t stream) {
    PB_REQUIRE(g && gate && add && B > 0 && R > 0 && C > 0 && C % 8 == 0, "scale_add: bad args");
    long long total = (long long)B * R * (C / 8);
    PB_DISPATCH_DTYPE(dtype, {
        rowscale_kernel<T, true><<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>((const T*)g, gate, add, (T*)g, R, C, total);
    });
    PB_CHECK_LAUNCH("scale_add");
    return PB_OK;
}
