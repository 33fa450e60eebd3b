// Squeeze-excite fully connected layers (SEBlock3D.se[1], se[3]; mobilenet.py:17-20) for a batch of pooled
// vectors: tiny GEMMs (B = 64 samples, C <= 960, C/4 hidden units) that must not cost a kernel per sample.
//   hidden = relu(W1 mean + b1)            [B][Ch]
//   gate   = hardsigmoid(W2 hidden + b2)   [B][C]
// and their backward.  Two kernel shapes cover everything:
//   fc_rows: W stored [N][K] (outputs are rows): one warp per output, lanes stride over K, BT samples per CTA
//   fc_cols: W stored [K][N] (outputs are columns): one thread per output, sequential over K
#include <algorithm>

#include "common.cuh"

namespace pb {

constexpr int FC_BT = 8;       // samples per CTA; their input vectors are staged in shared memory

// The weights of these layers are parameters: nothing in flight writes them, but by the time a layer runs again they
// have long left L2.  Each CTA asks for its slice BEFORE griddepcontrol.wait, i.e. while the kernel that produces
// its inputs is still running (PDL), so the cold misses overlap that kernel instead of heading a latency chain.
__device__ __forceinline__ void prefetch_rows_l2(const float* base, int rows, long long row_stride, int row_floats) {
    const int lines = (row_floats * 4 + 127) >> 7;
    for (int i = threadIdx.x; i < rows * lines; i += blockDim.x) {
        const int r = i / lines, l = i - r * lines;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (long long)r * row_stride + l * 32));
    }
}

// The FC_BT input rows are contiguous in X ([B][K] row-major), so the staging is one flat copy.  Loads are issued
// four at a time into registers BEFORE the stores: with one load in flight per thread this loop was a 20 us
// latency chain for K = 960 (ncu: fc_rows 24.5 us, of which the products are ~2 us).
__device__ __forceinline__ void stage_x(const float* __restrict__ X, float* xs, int b0, int B, int K) {
    const int nb = min(FC_BT, B - b0);
    const int n = nb * K, total = FC_BT * K;
    const float* src = X + (long long)b0 * K;
    const int nt = blockDim.x;
    if ((K & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        const int n4 = n >> 2, t4 = total >> 2;
        for (int base = threadIdx.x; base < t4; base += 4 * nt) {
            float4 v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int i = base + j * nt;
                v[j] = i < n4 ? __ldg(reinterpret_cast<const float4*>(src) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int i = base + j * nt;
                if (i < t4) reinterpret_cast<float4*>(xs)[i] = v[j];
            }
        }
    } else {
        for (int base = threadIdx.x; base < total; base += 8 * nt) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int i = base + j * nt;
                v[j] = i < n ? __ldg(src + i) : 0.f;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int i = base + j * nt;
                if (i < total) xs[i] = v[j];
            }
        }
    }
    __syncthreads();
}

// Y[b][n] = act(bias[n] + sum_k X[b][k] * W[n][k]).  CTA = 32 outputs x FC_BT samples; a warp owns 4 outputs,
// lanes stride over K (coalesced weight rows, conflict-free shared-memory reads of X).
template <int ACT>
__global__ void __launch_bounds__(256)
fc_rows_kernel(const float* __restrict__ X, const float* __restrict__ W, const float* __restrict__ bias,
               float* __restrict__ Y, int B, int N, int K) {
    pdl_trigger();
    prefetch_rows_l2(W + (long long)blockIdx.x * 32 * K, min(32, N - (int)blockIdx.x * 32), K, K);
    pdl_wait();
    extern __shared__ float xs[];                       // [FC_BT][K]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b0 = blockIdx.y * FC_BT;
    stage_x(X, xs, b0, B, K);
    const int n0 = blockIdx.x * 32 + warp * 4;
    float acc[4][FC_BT];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < FC_BT; ++i) acc[j][i] = 0.f;
    // weights are cold (HBM) in a training step and this kernel is a latency chain: 16 loads per lane are issued
    // before the first product (the compiler would not batch them across the loop's bounds checks on its own)
    const float* wrow[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) wrow[j] = W + (long long)min(n0 + j, N - 1) * K;
    int k = lane;
    for (; k + 96 < K; k += 128) {
        float wv[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int j = 0; j < 4; ++j) wv[u][j] = __ldg(wrow[j] + k + 32 * u);
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < FC_BT; ++i) {
                const float xv = xs[i * K + k + 32 * u];
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[j][i] = fmaf(xv, wv[u][j], acc[j][i]);
            }
    }
    for (; k < K; k += 32) {
        float wv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) wv[j] = __ldg(wrow[j] + k);
#pragma unroll
        for (int i = 0; i < FC_BT; ++i) {
            const float xv = xs[i * K + k];
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j][i] = fmaf(xv, wv[j], acc[j][i]);
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < FC_BT; ++i) acc[j][i] = warp_sum(acc[j][i]);
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + j;
            if (n >= N) continue;
            const float bv = bias ? bias[n] : 0.f;
#pragma unroll
            for (int i = 0; i < FC_BT; ++i)
                if (b0 + i < B) Y[(long long)(b0 + i) * N + n] = act_fwd(acc[j][i] + bv, ACT, 0.f);
        }
    }
}

// Y[b][n] = scale * sum_k X[b][k] * Wt[k][n], optionally masked by (relu_ref[b][n] > 0)  (relu backward).
// blockDim = (32 outputs, 8 k-slices): coalesced weight reads along n, X broadcast from shared memory, K split
// over the 8 slices, then a shared-memory reduction over the slices.
__global__ void __launch_bounds__(256)
fc_cols_kernel(const float* __restrict__ X, const float* __restrict__ Wt, const float* __restrict__ relu_ref,
               float* __restrict__ Y, int B, int N, int K, float scale) {
    pdl_trigger();
    prefetch_rows_l2(Wt + blockIdx.x * 32, K, N, min(32, N - (int)blockIdx.x * 32));
    pdl_wait();
    extern __shared__ float xs[];                       // [FC_BT][K] | red[8][FC_BT][33]
    float* red = xs + FC_BT * K;
    const int nl = threadIdx.x & 31, ks = threadIdx.x >> 5;
    const int n = blockIdx.x * 32 + nl;
    const int b0 = blockIdx.y * FC_BT;
    stage_x(X, xs, b0, B, K);
    float acc[FC_BT];
#pragma unroll
    for (int i = 0; i < FC_BT; ++i) acc[i] = 0.f;
    if (n < N) {
        int k = ks;
        for (; k + 56 < K; k += 64) {                       // eight weight rows in flight per thread
            float wv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) wv[u] = __ldg(Wt + (long long)(k + 8 * u) * N + n);
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
                for (int i = 0; i < FC_BT; ++i) acc[i] = fmaf(xs[i * K + k + 8 * u], wv[u], acc[i]);
        }
        for (; k < K; k += 8) {
            const float wv = __ldg(Wt + (long long)k * N + n);
#pragma unroll
            for (int i = 0; i < FC_BT; ++i) acc[i] = fmaf(xs[i * K + k], wv, acc[i]);
        }
    }
#pragma unroll
    for (int i = 0; i < FC_BT; ++i) red[(ks * FC_BT + i) * 33 + nl] = acc[i];
    __syncthreads();
    {   // thread (nl, ks) finishes sample ks
        const int i = ks, b = b0 + i;
        if (i < FC_BT && n < N && b < B) {
            float v = 0.f;
#pragma unroll
            for (int s2 = 0; s2 < 8; ++s2) v += red[(s2 * FC_BT + i) * 33 + nl];
            v *= scale;
            if (relu_ref && !(relu_ref[(long long)b * N + n] > 0.f)) v = 0.f;
            Y[(long long)b * N + n] = v;
        }
    }
}

// da2 = dgate * hardsigmoid'(.) : 1/6 where 0 < gate < 1
__global__ void hsig_bwd_kernel(const float* __restrict__ dgate, const float* __restrict__ gate,
                                float* __restrict__ da2, long long n) {
    pdl_trigger();
    pdl_wait();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float g = gate[i];
    da2[i] = (g > 0.f && g < 1.f) ? dgate[i] * (1.f / 6.f) : 0.f;
}

// Parameter gradients: outer products summed over the B samples,
//   dW2[c][j] = sum_b da2[b][c] * hidden[b][j]   db2[c] = sum_b da2[b][c]       (blockIdx.z == 0)
//   dW1[j][c] = sum_b da1[b][j] * mean[b][c]     db1[j] = sum_b da1[b][j]       (blockIdx.z == 1)
// i.e. out[r][c] = sum_b X[b][r] * Y[b][c] twice.  A CTA owns a 64 x 64 tile of `out`: the samples' X and Y
// columns are staged 32 samples at a time in shared memory, a thread keeps a 4 x 4 register tile (two LDS.128 per
// 16 FMAs); the row sums (bias gradients) come from the threads of the first column of tiles.
__global__ void __launch_bounds__(256)
se_fc_bwd_param_kernel(const float* __restrict__ a2, const float* __restrict__ a1,
                       const float* __restrict__ mean, const float* __restrict__ hidden,
                       float* __restrict__ dW1, float* __restrict__ db1, float* __restrict__ dW2,
                       float* __restrict__ db2, int B, int C, int Ch) {
    pdl_trigger();
    pdl_wait();
    constexpr int BC = 32;
    __shared__ __align__(16) float xs[BC][64], ys[BC][64];
    const bool second = blockIdx.z == 1;
    const float* X = second ? a1 : a2;
    const float* Y = second ? mean : hidden;
    const int Rn = second ? Ch : C, Cn = second ? C : Ch;
    float* out = second ? dW1 : dW2;
    float* rowsum = second ? db1 : db2;
    const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
    if (r0 >= Rn || c0 >= Cn) return;
    const int tr = threadIdx.x >> 4, tc = threadIdx.x & 15;
    float acc[4][4], rs[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        rs[i] = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    }
    for (int b0 = 0; b0 < B; b0 += BC) {
#pragma unroll
        for (int e = threadIdx.x; e < BC * 64; e += 256) {      // 8 iterations: all 16 loads of a thread in flight
            const int bi = e >> 6, k = e & 63, b = b0 + bi;
            xs[bi][k] = (b < B && r0 + k < Rn) ? __ldg(X + (long long)b * Rn + r0 + k) : 0.f;
            ys[bi][k] = (b < B && c0 + k < Cn) ? __ldg(Y + (long long)b * Cn + c0 + k) : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int bi = 0; bi < BC; ++bi) {
            const float4 xv = *reinterpret_cast<const float4*>(&xs[bi][tr * 4]);
            const float4 yv = *reinterpret_cast<const float4*>(&ys[bi][tc * 4]);
            const float x[4] = {xv.x, xv.y, xv.z, xv.w}, y[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                rs[i] += x[i];
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(x[i], y[j], acc[i][j]);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = r0 + tr * 4 + i;
        if (r >= Rn) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c0 + tc * 4 + j;
            if (c < Cn) out[(long long)r * Cn + c] = acc[i][j];
        }
        if (blockIdx.x == 0 && tc == 0) rowsum[r] = rs[i];
    }
}

}  // namespace pb

using namespace pb;

// shared-memory staging of FC_BT input vectors (+ the k-slice reduction scratch): allow up to 96 KB
constexpr int FC_MAX_K = 2560;
static int fc_smem_attrs() {
    static unsigned long long done[4] = {0, 0, 0, 0};
    const int bytes = 96 * 1024;
    cudaError_t e = ensure_dyn_smem(fc_rows_kernel<PB_ACT_NONE>, bytes, &done[0]);
    if (e == cudaSuccess) e = ensure_dyn_smem(fc_rows_kernel<PB_ACT_RELU>, bytes, &done[1]);
    if (e == cudaSuccess) e = ensure_dyn_smem(fc_rows_kernel<PB_ACT_HSIGMOID>, bytes, &done[2]);
    if (e == cudaSuccess) e = ensure_dyn_smem(fc_cols_kernel, bytes, &done[3]);
    return e == cudaSuccess ? PB_OK : cuda_fail(e, "cudaFuncSetAttribute(fc kernels)");
}

// Classifier-head linear layers (mobilenet.py:184-190, movinet.py:146-154): B = clips per call (64), so the
// work is reading the weight matrix once; same kernels as the squeeze-excite layers.
extern "C" int pb_fc_fwd(const float* X, const float* W, const float* bias, float* Y, int B, int N, int K,
                         pb_stream_t stream) {
    PB_REQUIRE(X && W && Y && B > 0 && N > 0 && K > 0, "fc_fwd: bad args");
    PB_REQUIRE(K <= FC_MAX_K, "fc_fwd: K=%d too large for the shared-memory staging", K);
    if (int e = fc_smem_attrs()) return e;
    dim3 g(ceil_div(N, 32), ceil_div(B, FC_BT));
    (void)launch_pdl(fc_rows_kernel<PB_ACT_NONE>, g, dim3(256), sizeof(float) * FC_BT * K, (cudaStream_t)stream, X, W, bias,
                     Y, B, N, K);
    PB_CHECK_LAUNCH("fc_fwd");
    return PB_OK;
}

extern "C" int pb_fc_dgrad(const float* dY, const float* W, float* dX, int B, int N, int K, float scale,
                           pb_stream_t stream) {
    PB_REQUIRE(dY && W && dX && B > 0 && N > 0 && K > 0, "fc_dgrad: bad args");
    PB_REQUIRE(N <= FC_MAX_K - 8 * 33, "fc_dgrad: N=%d too large for the shared-memory staging", N);
    if (int e = fc_smem_attrs()) return e;
    // dX[b][k] = scale * sum_n dY[b][n] W[n][k]: W [N][K] is the "Wt" of fc_cols with (K', N') = (N, K)
    dim3 g(ceil_div(K, 32), ceil_div(B, FC_BT));
    const size_t red_bytes = sizeof(float) * 8 * FC_BT * 33;
    (void)launch_pdl(fc_cols_kernel, g, dim3(256), sizeof(float) * FC_BT * N + red_bytes, (cudaStream_t)stream, dY, W,
                     (const float*)nullptr, dX, B, K, N, scale);
    PB_CHECK_LAUNCH("fc_dgrad");
    return PB_OK;
}

extern "C" int pb_se_fc_fwd(const float* mean, const float* W1, const float* b1, const float* W2, const float* b2,
                            float* hidden, float* gate, int B, int C, int Ch, pb_stream_t stream) {
    PB_REQUIRE(mean && W1 && b1 && W2 && b2 && hidden && gate && B > 0 && C > 0 && Ch > 0, "se_fc_fwd: bad args");
    cudaStream_t st = (cudaStream_t)stream;
    PB_REQUIRE(C <= FC_MAX_K && Ch <= FC_MAX_K, "se_fc_fwd: channel count too large for the shared-memory staging");
    if (int e = fc_smem_attrs()) return e;
    dim3 g1(ceil_div(Ch, 32), ceil_div(B, FC_BT));
    (void)launch_pdl(fc_rows_kernel<PB_ACT_RELU>, dim3(g1), dim3(256), sizeof(float) * FC_BT * C, st, mean, W1, b1,
                     hidden, B, Ch, C);
    PB_CHECK_LAUNCH("se_fc_fwd(1)");
    dim3 g2(ceil_div(C, 32), ceil_div(B, FC_BT));
    (void)launch_pdl(fc_rows_kernel<PB_ACT_HSIGMOID>, dim3(g2), dim3(256), sizeof(float) * FC_BT * Ch, st, hidden, W2,
                     b2, gate, B, C, Ch);
    PB_CHECK_LAUNCH("se_fc_fwd(2)");
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
    float* a2 = work;                          // [B][C]
    float* a1 = work + (long long)B * C;       // [B][Ch]
    const long long nBC = (long long)B * C;
    (void)launch_pdl(hsig_bwd_kernel, dim3(ceil_div(nBC, 256)), dim3(256), 0, st, dgate, gate, a2, nBC);
    PB_CHECK_LAUNCH("se_fc_bwd(hsig)");
    // da1[b][j] = relu'(hidden) * sum_c da2[b][c] * W2[c][j]   (W2 is [C][Ch] = "Wt" with K=C, N=Ch)
    dim3 g1(ceil_div(Ch, 32), ceil_div(B, FC_BT));
    PB_REQUIRE(C <= FC_MAX_K - 8 * 33 && Ch <= FC_MAX_K - 8 * 33,
               "se_fc_bwd: channel count too large for the shared-memory staging");
    if (int e = fc_smem_attrs()) return e;
    const size_t red_bytes = sizeof(float) * 8 * FC_BT * 33;
    (void)launch_pdl(fc_cols_kernel, dim3(g1), dim3(256), sizeof(float) * FC_BT * C + red_bytes, st, a2, W2, hidden,
                     a1, B, Ch, C, 1.f);
    PB_CHECK_LAUNCH("se_fc_bwd(da1)");
    // dmean[b][c] = inv_R * sum_j da1[b][j] * W1[j][c]         (W1 is [Ch][C] = "Wt" with K=Ch, N=C)
    dim3 g2(ceil_div(C, 32), ceil_div(B, FC_BT));
    (void)launch_pdl(fc_cols_kernel, dim3(g2), dim3(256), sizeof(float) * FC_BT * Ch + red_bytes, st, a1, W1, nullptr,
                     dmean, B, C, Ch, inv_R);
    PB_CHECK_LAUNCH("se_fc_bwd(dmean)");
    const int big = std::max(C, Ch);
    (void)launch_pdl(se_fc_bwd_param_kernel, dim3(ceil_div(big, 64), ceil_div(big, 64), 2), dim3(256), 0, st, a2, a1, mean, hidden, dW1, db1,
                     dW2, db2, B, C, Ch);
    PB_CHECK_LAUNCH("se_fc_bwd(param)");
    return PB_OK;
}
