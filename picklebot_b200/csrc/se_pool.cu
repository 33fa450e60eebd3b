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

template <typename T, bool ADD>
__global__ void __launch_bounds__(256)
rowscale_kernel(const T* x, const float* __restrict__ gate, const float* __restrict__ add,
                T* y, long long R, int C, long long total) {   // x may alias y (in-place)
    pdl_trigger();
    pdl_wait();
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

// Cumulative mean for the streaming mode (MoViNet stream buffers, config 4): one block, so that every thread reads
// the old row count before thread 0 advances it.
__global__ void __launch_bounds__(1024)
stream_pool_update_kernel(const float* __restrict__ chunk_mean, long long R, float* __restrict__ sum,
                          long long* __restrict__ rows, float* __restrict__ mean_out, int n) {
    pdl_trigger();
    pdl_wait();
    const long long seen = *rows + R;
    __syncthreads();
    const float inv = 1.0f / (float)seen, fr = (float)R;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float sacc = sum[i] + chunk_mean[i] * fr;
        sum[i] = sacc;
        mean_out[i] = sacc * inv;
    }
    if (threadIdx.x == 0) *rows = seen;
}

}  // namespace pb

using namespace pb;

extern "C" int pb_stream_pool_update(const float* chunk_mean, long long R, float* sum, long long* rows, float* mean_out,
                                     int B, int C, pb_stream_t stream) {
    PB_REQUIRE(chunk_mean && sum && rows && mean_out && B > 0 && C > 0 && R > 0, "stream_pool_update: bad args");
    (void)launch_pdl(stream_pool_update_kernel, dim3(1), dim3(1024), 0, (cudaStream_t)stream, chunk_mean, R, sum, rows,
                     mean_out, B * C);
    PB_CHECK_LAUNCH("stream_pool_update_kernel");
    return PB_OK;
}

extern "C" int pb_pool_fwd(const void* x, int dtype, int B, long long R, int C, float* mean, pb_stream_t stream) {
    PB_REQUIRE(x && mean && B > 0 && B <= 65535 && R > 0 && C > 0 && C % 8 == 0 && C <= 2048, "pool_fwd: bad args");
    cudaStream_t st = (cudaStream_t)stream;
    PB_CUDA(cudaMemsetAsync(mean, 0, sizeof(float) * (size_t)B * C, st));
    dim3 grid = colreduce_grid(R, C, B);
    PB_DISPATCH_DTYPE(dtype, {
        PoolF<T> f{(const T*)x, R, C};
        (void)launch_pdl(colreduce_kernel<PoolF<T>, 1, float>, dim3(grid), dim3(256), 0, st, f, R, C, mean, B,
                         1.f / (float)R);
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
        (void)launch_pdl(colreduce_kernel<DotF<T>, 1, float>, dim3(grid), dim3(256), 0, st, f, R, C, out, B, 1.f);
    });
    PB_CHECK_LAUNCH("rowdot");
    return PB_OK;
}

extern "C" int pb_rowscale(const void* x, const float* gate, void* y, int dtype, int B, long long R, int C,
                           pb_stream_t stream) {
    PB_REQUIRE(x && gate && y && B > 0 && R > 0 && C > 0 && C % 8 == 0, "rowscale: bad args");
    long long total = (long long)B * R * (C / 8);
    PB_DISPATCH_DTYPE(dtype, {
        (void)launch_pdl(rowscale_kernel<T, false>, dim3(ceil_div(total, 256)), dim3(256), 0, (cudaStream_t)stream,
                         (const T*)x, gate, nullptr, (T*)y, R, C, total);
    });
    PB_CHECK_LAUNCH("rowscale");
    return PB_OK;
}

extern "C" int pb_scale_add(void* g, const float* gate, const float* add, int dtype, int B, long long R, int C,
                            pb_stream_t stream) {
    PB_REQUIRE(g && gate && add && B > 0 && R > 0 && C > 0 && C % 8 == 0, "scale_add: bad args");
    long long total = (long long)B * R * (C / 8);
    PB_DISPATCH_DTYPE(dtype, {
        (void)launch_pdl(rowscale_kernel<T, true>, dim3(ceil_div(total, 256)), dim3(256), 0, (cudaStream_t)stream,
                         (const T*)g, gate, add, (T*)g, R, C, total);
    });
    PB_CHECK_LAUNCH("scale_add");
    return PB_OK;
}
