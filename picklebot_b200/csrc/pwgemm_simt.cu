// Pointwise-conv / FC GEMMs on CUDA cores: the fp32 parity path (1e-4 tolerance) and the small-M
// classifier layers.  The bf16 production path is the tcgen05 kernel in pwgemm_tc.cu.
// Replaces nn.Conv3d(kernel_size=1) (mobilenet.py:64,79; movinet.py:47,63) and nn.Linear (movinet.py:149,153).
#include <algorithm>

#include "common.cuh"

namespace pb {

constexpr int BM = 64, BN = 64, BK = 16;

// C[b][r][n] = (sum_k A[b][r][k]*ascale[b][k] * W(n,k) + bias[n]) * colscale[b][n] + coladd[b][n]
template <typename T>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const T* __restrict__ A, const float* __restrict__ W, long long w_sn, long long w_sk,
                 const float* __restrict__ bias, const float* __restrict__ ascale,
                 const float* __restrict__ colscale, const float* __restrict__ coladd,
                 T* __restrict__ C, long long R, int K, int N) {
    pdl_trigger();
    pdl_wait();
    __shared__ float As[BK][BM + 4];
    __shared__ float Ws[BK][BN + 4];
    const int b = blockIdx.z;
    const long long m0 = (long long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const T* Ab = A + (long long)b * R * K;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int e = tid + i * 256;
            int kk = e & 15, mm = e >> 4;
            long long m = m0 + mm;
            int k = k0 + kk;
            float v = 0.f;
            if (m < R && k < K) {
                v = to_float(Ab[m * K + k]);
                if (ascale) v = round_to<T>(v * ascale[(long long)b * K + k]);
            }
            As[kk][mm] = v;
            int n = n0 + mm;
            float w = 0.f;
            if (n < N && k < K) w = round_to<T>(W[n * w_sn + k * w_sk]);
            Ws[kk][mm] = w;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[4], w[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) w[j] = Ws[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
        }
        __syncthreads();
    }
    T* Cb = C + (long long)b * R * N;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        long long m = m0 + ty * 4 + i;
        if (m >= R) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float v = acc[i][j];
            if (bias) v += bias[n];
            if (colscale) v *= colscale[(long long)b * N + n];
            if (coladd) v += coladd[(long long)b * N + n];
            Cb[m * N + n] = from_float<T>(v);
        }
    }
}

// dW[n][k] += sum over this CTA's rows of dC[m][n] * A[m][k] * ascale[b(m)][k]
template <typename T>
__global__ void __launch_bounds__(256)
wgrad_simt_kernel(const T* __restrict__ A, const T* __restrict__ dC, const float* __restrict__ ascale,
                  float* __restrict__ dW, float* __restrict__ dbias, long long M, long long R, int K, int N,
                  long long rows_per_cta) {
    pdl_trigger();
    pdl_wait();
    __shared__ float Ds[BK][BN + 4];
    __shared__ float As[BK][BM + 4];
    const int n0 = blockIdx.y * BN, k0 = blockIdx.z * BM;
    const long long r_begin = (long long)blockIdx.x * rows_per_cta;
    const long long r_end = min(M, r_begin + rows_per_cta);
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    float bsum = 0.f;
    for (long long r0 = r_begin; r0 < r_end; r0 += BK) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int e = tid + i * 256;
            int cc = e & 63, mm = e >> 6;
            long long m = r0 + mm;
            float dv = 0.f, av = 0.f;
            if (m < r_end) {
                if (n0 + cc < N) dv = to_float(dC[m * N + n0 + cc]);
                if (k0 + cc < K) {
                    av = to_float(A[m * K + k0 + cc]);
                    if (ascale) av = round_to<T>(av * ascale[(m / R) * K + k0 + cc]);
                }
            }
            Ds[mm][cc] = dv;
            As[mm][cc] = av;
        }
        __syncthreads();
#pragma unroll
        for (int mm = 0; mm < BK; ++mm) {
            float dvv[4], avv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) dvv[i] = Ds[mm][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) avv[j] = As[mm][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(dvv[i], avv[j], acc[i][j]);
        }
        if (dbias && blockIdx.z == 0 && tid < BN) {
#pragma unroll
            for (int mm = 0; mm < BK; ++mm) bsum += Ds[mm][tid];
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int n = n0 + ty * 4 + i;
        if (n >= N) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int k = k0 + tx * 4 + j;
            if (k >= K) continue;
            atomicAdd(&dW[(long long)n * K + k], acc[i][j]);
        }
    }
    if (dbias && blockIdx.z == 0 && tid < BN && n0 + tid < N) atomicAdd(&dbias[n0 + tid], bsum);
}

template <typename TD>
__global__ void cast_matrix_kernel(const float* __restrict__ src, TD* __restrict__ dst, int rows, int cols,
                                   int transpose, int round_bf16) {
    pdl_trigger();
    pdl_wait();
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)rows * cols) return;
    int r = (int)(idx / cols), c = (int)(idx % cols);
    float v = src[idx];
    if (round_bf16) v = __bfloat162float(__float2bfloat16_rn(v));
    long long o = transpose ? (long long)c * rows + r : idx;
    dst[o] = from_float<TD>(v);
}

// 8 consecutive k per thread: two float4 loads of W and of the gate, one 16-byte bf16 store
__global__ void fold_gate_kernel(const float* __restrict__ W, const float* __restrict__ gate,
                                 __nv_bfloat16* __restrict__ dst, int N, int K, long long total8) {
    pdl_trigger();
    pdl_wait();
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total8) return;
    const int K8 = K >> 3;
    const int k = (int)(idx % K8) << 3;
    const long long q = idx / K8;
    const int n = (int)(q % N);
    const int b = (int)(q / N);
    const float4* wp = reinterpret_cast<const float4*>(W + (long long)n * K + k);
    const float4* gp = reinterpret_cast<const float4*>(gate + (long long)b * K + k);
    const float4 w0 = __ldg(wp), w1 = __ldg(wp + 1), g0 = __ldg(gp), g1 = __ldg(gp + 1);
    uint4 o;
    o.x = pack_bf16x2(w0.x * g0.x, w0.y * g0.y); o.y = pack_bf16x2(w0.z * g0.z, w0.w * g0.w);
    o.z = pack_bf16x2(w1.x * g1.x, w1.y * g1.y); o.w = pack_bf16x2(w1.z * g1.z, w1.w * g1.w);
    *reinterpret_cast<uint4*>(dst + idx * 8) = o;
}

// Inference: dst[b][n][k] = W[n][k] * gate[b][k] * rowscale[n] -- the squeeze-excite gate (optional) and the scale of
// the eval-mode BatchNorm that follows the convolution (optional) folded into one bf16 weight matrix per sample.
__global__ void fold_scaled_kernel(const float* __restrict__ W, const float* __restrict__ gate,
                                   const float* __restrict__ rowscale, __nv_bfloat16* __restrict__ dst, int N, int K,
                                   long long total8) {
    pdl_trigger();
    pdl_wait();
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total8) return;
    const int K8 = K >> 3;
    const int k = (int)(idx % K8) << 3;
    const long long q = idx / K8;
    const int n = (int)(q % N);
    const int b = (int)(q / N);
    const float4* wp = reinterpret_cast<const float4*>(W + (long long)n * K + k);
    const float4 w0 = __ldg(wp), w1 = __ldg(wp + 1);
    float4 g0 = make_float4(1.f, 1.f, 1.f, 1.f), g1 = g0;
    if (gate) {
        const float4* gp = reinterpret_cast<const float4*>(gate + (long long)b * K + k);
        g0 = __ldg(gp); g1 = __ldg(gp + 1);
    }
    const float rs = rowscale ? __ldg(rowscale + n) : 1.f;
    uint4 o;
    o.x = pack_bf16x2(w0.x * g0.x * rs, w0.y * g0.y * rs); o.y = pack_bf16x2(w0.z * g0.z * rs, w0.w * g0.w * rs);
    o.z = pack_bf16x2(w1.x * g1.x * rs, w1.y * g1.y * rs); o.w = pack_bf16x2(w1.z * g1.z * rs, w1.w * g1.w * rs);
    *reinterpret_cast<uint4*>(dst + idx * 8) = o;
}

// Row-scaled fold from an already TRANSPOSED fp32 weight: dst[b][r][c] = Wt[r][c] * gate[b][r]  (Wt fp32 [R][C]).
// Same result as fold_gate_t_kernel, but every access is a coalesced 16/32-byte vector (the strided gather of the
// transposing kernel took 16 us per launch on the 112 x 672 layer).
__global__ void fold_rows_kernel(const float* __restrict__ Wt, const float* __restrict__ gate,
                                 __nv_bfloat16* __restrict__ dst, int R, int C, long long total8) {
    pdl_trigger();
    pdl_wait();
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total8) return;
    const int C8 = C >> 3;
    const int c = (int)(idx % C8) << 3;
    const long long q = idx / C8;
    const int r = (int)(q % R);
    const int b = (int)(q / R);
    const float4* wp = reinterpret_cast<const float4*>(Wt + (long long)r * C + c);
    const float4 w0 = __ldg(wp), w1 = __ldg(wp + 1);
    const float g = __ldg(gate + (long long)b * R + r);
    uint4 o;
    o.x = pack_bf16x2(w0.x * g, w0.y * g); o.y = pack_bf16x2(w0.z * g, w0.w * g);
    o.z = pack_bf16x2(w1.x * g, w1.y * g); o.w = pack_bf16x2(w1.z * g, w1.w * g);
    *reinterpret_cast<uint4*>(dst + idx * 8) = o;
}

// Transposed twin for the input-gradient GEMM of a squeeze-excite block: dst[b][k][n] = W[n][k] * gate[b][k]
// (W fp32 [N][K] as stored by the layer, dst bf16 [B][K][N]): the gate scales the ROWS of the transposed weight,
// i.e. the output columns of  dy2 = dz W  -- folded here so that GEMM needs no per-column epilogue vector.
__global__ void fold_gate_t_kernel(const float* __restrict__ W, const float* __restrict__ gate,
                                   __nv_bfloat16* __restrict__ dst, int N, int K, long long total8) {
    pdl_trigger();
    pdl_wait();
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total8) return;
    const int N8 = N >> 3;
    const int n = (int)(idx % N8) << 3;
    const long long q = idx / N8;
    const int k = (int)(q % K);
    const int b = (int)(q / K);
    const float g = __ldg(gate + (long long)b * K + k);
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __ldg(W + (long long)(n + i) * K + k) * g;
    uint4 o;
    o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
    o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(dst + idx * 8) = o;
}

// dst[(i*N + n)][(j*K + k)] = (i == j) ? W[n][k] : 0  -- the weight of a row-folded GEMM (pwgemm_tc.cu)
__global__ void block_diag_kernel(const __nv_bfloat16* __restrict__ W, __nv_bfloat16* __restrict__ dst, int F, int N,
                                  int K, long long total) {
    pdl_trigger();
    pdl_wait();
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int col = (int)(idx % ((long long)F * K));
    const int row = (int)(idx / ((long long)F * K));
    const int i = row / N, j = col / K;
    dst[idx] = i == j ? W[(long long)(row - i * N) * K + (col - j * K)] : __float2bfloat16_rn(0.f);
}

}  // namespace pb

using namespace pb;

extern "C" int pb_pw_gemm_simt(const void* A, const float* W, long long w_sn, long long w_sk, const float* bias,
                               const float* ascale, const float* colscale, const float* coladd, void* C,
                               int dtype, int Bt, long long R, int K, int N, pb_stream_t stream) {
    PB_REQUIRE(A && W && C, "pw_gemm_simt: null pointer");
    PB_REQUIRE(Bt > 0 && R > 0 && K > 0 && N > 0, "pw_gemm_simt: empty problem Bt=%d R=%lld K=%d N=%d", Bt, R, K, N);
    PB_REQUIRE(Bt <= 65535 && ceil_div(N, BN) <= 65535, "pw_gemm_simt: grid too large");
    dim3 grid(ceil_div(R, BM), ceil_div(N, BN), Bt);
    PB_DISPATCH_DTYPE(dtype, {
        (void)launch_pdl(gemm_simt_kernel<T>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const T*)A, W, w_sn,
                         w_sk, bias, ascale, colscale, coladd, (T*)C, R, K, N);
    });
    PB_CHECK_LAUNCH("gemm_simt_kernel");
    count_path(PB_PATH_GEMM_SIMT);
    return PB_OK;
}

extern "C" int pb_pw_wgrad_simt(const void* A, const void* dC, const float* ascale, float* dW, float* dbias,
                                int dtype, int Bt, long long R, int K, int N, pb_stream_t stream) {
    PB_REQUIRE(A && dC && dW, "pw_wgrad_simt: null pointer");
    PB_REQUIRE(Bt > 0 && R > 0 && K > 0 && N > 0, "pw_wgrad_simt: empty problem");
    cudaStream_t st = (cudaStream_t)stream;
    PB_CUDA(cudaMemsetAsync(dW, 0, sizeof(float) * (size_t)N * K, st));
    if (dbias) PB_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * (size_t)N, st));
    long long M = (long long)Bt * R;
    int tiles = ceil_div(N, BN) * ceil_div(K, BM);
    long long want_ctas = std::max<long long>(1, (148LL * 6) / tiles);
    long long rows = std::max<long long>(256, (M + want_ctas - 1) / want_ctas);
    rows = (rows + BK - 1) / BK * BK;
    dim3 grid(ceil_div(M, rows), ceil_div(N, BN), ceil_div(K, BM));
    PB_DISPATCH_DTYPE(dtype, {
        (void)launch_pdl(wgrad_simt_kernel<T>, dim3(grid), dim3(256), 0, st, (const T*)A, (const T*)dC, ascale, dW,
                         dbias, M, R, K, N, rows);
    });
    PB_CHECK_LAUNCH("wgrad_simt_kernel");
    count_path(PB_PATH_WGRAD_SIMT);
    return PB_OK;
}

extern "C" int pb_cast_matrix(const float* src, void* dst, int dst_dtype, int rows, int cols, int transpose,
                              pb_stream_t stream) {
    PB_REQUIRE(src && dst && rows > 0 && cols > 0, "cast_matrix: bad args");
    long long n = (long long)rows * cols;
    cudaStream_t st = (cudaStream_t)stream;
    if (dst_dtype == PB_BF16)
        (void)launch_pdl(cast_matrix_kernel<__nv_bfloat16>, dim3(ceil_div(n, 256)), dim3(256), 0, st, src,
                         (__nv_bfloat16*)dst, rows, cols, transpose, 0);
    else if (dst_dtype == PB_F32)
        (void)launch_pdl(cast_matrix_kernel<float>, dim3(ceil_div(n, 256)), dim3(256), 0, st, src, (float*)dst, rows,
                         cols, transpose, 0);
    else if (dst_dtype == PB_F32_RBF16)   // fp32 storage, values rounded through bf16 (autocast's weight cast)
        (void)launch_pdl(cast_matrix_kernel<float>, dim3(ceil_div(n, 256)), dim3(256), 0, st, src, (float*)dst, rows,
                         cols, transpose, 1);
    else { set_error("cast_matrix: bad dst dtype %d", dst_dtype); return PB_ERR_BAD_ARG; }
    PB_CHECK_LAUNCH("cast_matrix_kernel");
    return PB_OK;
}

extern "C" int pb_fold_gate_bf16(const float* W, const float* gate, void* dst, int Bt, int N, int K,
                                 pb_stream_t stream) {
    PB_REQUIRE(W && gate && dst && Bt > 0 && N > 0 && K > 0 && K % 8 == 0, "fold_gate: bad args (K %% 8 == 0)");
    PB_REQUIRE(((reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(gate) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0,
               "fold_gate: pointers must be 16-byte aligned");
    long long n = (long long)Bt * N * (K / 8);
    (void)launch_pdl(fold_gate_kernel, dim3(ceil_div(n, 256)), dim3(256), 0, (cudaStream_t)stream, W, gate,
                     (__nv_bfloat16*)dst, N, K, n);
    PB_CHECK_LAUNCH("fold_gate_kernel");
    return PB_OK;
}

extern "C" int pb_fold_scaled_bf16(const float* W, const float* gate, const float* rowscale, void* dst, int Bt, int N,
                                   int K, pb_stream_t stream) {
    PB_REQUIRE(W && dst && Bt > 0 && N > 0 && K > 0 && K % 8 == 0, "fold_scaled: bad args (K %% 8 == 0)");
    PB_REQUIRE(gate || Bt == 1, "fold_scaled: Bt > 1 needs a gate");
    PB_REQUIRE(((reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(gate) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0,
               "fold_scaled: pointers must be 16-byte aligned");
    long long n = (long long)Bt * N * (K / 8);
    (void)launch_pdl(fold_scaled_kernel, dim3(ceil_div(n, 256)), dim3(256), 0, (cudaStream_t)stream, W, gate, rowscale,
                     (__nv_bfloat16*)dst, N, K, n);
    PB_CHECK_LAUNCH("fold_scaled_kernel");
    return PB_OK;
}

extern "C" int pb_fold_rows_bf16(const float* Wt, const float* gate, void* dst, int Bt, int R, int C, pb_stream_t stream) {
    PB_REQUIRE(Wt && gate && dst && Bt > 0 && R > 0 && C > 0 && C % 8 == 0, "fold_rows: bad args (C must be a multiple of 8)");
    PB_REQUIRE(((reinterpret_cast<uintptr_t>(Wt) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0, "fold_rows: pointers must be 16-byte aligned");
    long long n = (long long)Bt * R * (C / 8);
    (void)launch_pdl(fold_rows_kernel, dim3(ceil_div(n, 256)), dim3(256), 0, (cudaStream_t)stream, Wt, gate,
                     (__nv_bfloat16*)dst, R, C, n);
    PB_CHECK_LAUNCH("fold_rows_kernel");
    return PB_OK;
}

extern "C" int pb_fold_gate_t_bf16(const float* W, const float* gate, void* dst, int Bt, int N, int K, pb_stream_t stream) {
    PB_REQUIRE(W && gate && dst && Bt > 0 && N > 0 && K > 0 && N % 8 == 0, "fold_gate_t: bad args (N must be a multiple of 8)");
    PB_REQUIRE((reinterpret_cast<uintptr_t>(dst) & 15) == 0, "fold_gate_t: dst must be 16-byte aligned");
    long long n = (long long)Bt * K * (N / 8);
    (void)launch_pdl(fold_gate_t_kernel, dim3(ceil_div(n, 256)), dim3(256), 0, (cudaStream_t)stream, W, gate,
                     (__nv_bfloat16*)dst, N, K, n);
    PB_CHECK_LAUNCH("fold_gate_t_kernel");
    return PB_OK;
}

extern "C" int pb_block_diag_bf16(const void* W, void* dst, int F, int N, int K, pb_stream_t stream) {
    PB_REQUIRE(W && dst && F > 0 && N > 0 && K > 0, "block_diag: bad args");
    long long n = (long long)F * N * F * K;
    (void)launch_pdl(block_diag_kernel, dim3(ceil_div(n, 256)), dim3(256), 0, (cudaStream_t)stream,
                     (const __nv_bfloat16*)W, (__nv_bfloat16*)dst, F, N, K, n);
    PB_CHECK_LAUNCH("block_diag_kernel");
    return PB_OK;
}
