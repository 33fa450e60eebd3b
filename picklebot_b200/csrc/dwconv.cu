// Depthwise Conv3d (groups == C) forward / dgrad / wgrad, NDHWC, 8 channels per thread.
// Replaces Bottleneck3D.depthwise_conv (mobilenet.py:67-75) and MoviNetBottleneck.conv
// (movinet.py:52-61).  This file holds the general-shape kernels: any (kT,kH,kW), stride and padding,
// bounds handled per tap.  The tiled fast paths for the MobileNet (1,k,k) classes live in
// dwconv_tiled.cu and are selected by the same entry points.
#include <algorithm>

#include "common.cuh"
#include <cstdlib>

#include "dwconv.cuh"

namespace pb {

// ---------------------------------------------------------------------------------------------
// forward: one thread = one output pixel x 8 channels
// ---------------------------------------------------------------------------------------------
template <typename T, bool STREAM>
__global__ void __launch_bounds__(256)
dw_fwd_generic(const T* __restrict__ x, const T* __restrict__ sbuf, const float* __restrict__ w_tc,
               T* __restrict__ y, DwDims d, long long total) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int G = d.C >> 3;
    int g = (int)(idx % G);
    long long p = idx / G;
    int wo = (int)(p % d.Wo); p /= d.Wo;
    int ho = (int)(p % d.Ho); p /= d.Ho;
    int to = (int)(p % d.To);
    int b  = (int)(p / d.To);
    const int c0 = g << 3;
    F8 acc = zero8();
    for (int kt = 0; kt < d.kT; ++kt) {
        int ti = to * d.sT - d.pT + kt;
        const T* frame;
        if (STREAM) {
            // causal: pT == kT-1 on the left only; negative frames come from the stream buffer
            if (ti < 0) frame = sbuf + ((long long)b * (d.kT - 1) + (d.kT - 1 + ti)) * d.H * d.W * d.C;
            else        frame = x + ((long long)b * d.T + ti) * d.H * d.W * d.C;
        } else {
            if (ti < 0 || ti >= d.T) continue;
            frame = x + ((long long)b * d.T + ti) * d.H * d.W * d.C;
        }
        for (int kh = 0; kh < d.kH; ++kh) {
            int hi = ho * d.sH - d.pH + kh;
            if (hi < 0 || hi >= d.H) continue;
            for (int kw = 0; kw < d.kW; ++kw) {
                int wi = wo * d.sW - d.pW + kw;
                if (wi < 0 || wi >= d.W) continue;
                F8 xv = load8(frame + ((long long)hi * d.W + wi) * d.C + c0);
                const float* wp = w_tc + (long long)((kt * d.kH + kh) * d.kW + kw) * d.C + c0;
                float4 w0 = __ldg(reinterpret_cast<const float4*>(wp));
                float4 w1 = __ldg(reinterpret_cast<const float4*>(wp + 4));
                acc.v[0] = fmaf(xv.v[0], w0.x, acc.v[0]); acc.v[1] = fmaf(xv.v[1], w0.y, acc.v[1]);
                acc.v[2] = fmaf(xv.v[2], w0.z, acc.v[2]); acc.v[3] = fmaf(xv.v[3], w0.w, acc.v[3]);
                acc.v[4] = fmaf(xv.v[4], w1.x, acc.v[4]); acc.v[5] = fmaf(xv.v[5], w1.y, acc.v[5]);
                acc.v[6] = fmaf(xv.v[6], w1.z, acc.v[6]); acc.v[7] = fmaf(xv.v[7], w1.w, acc.v[7]);
            }
        }
    }
    store8(y + idx * 8, acc);
}

// ---------------------------------------------------------------------------------------------
// dgrad: one thread = one INPUT pixel x 8 channels, gathers the output pixels that used it
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
dw_dgrad_generic(const T* __restrict__ dy, const float* __restrict__ w_tc, T* __restrict__ dx,
                 DwDims d, long long total) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int G = d.C >> 3;
    int g = (int)(idx % G);
    long long p = idx / G;
    int wi = (int)(p % d.W); p /= d.W;
    int hi = (int)(p % d.H); p /= d.H;
    int ti = (int)(p % d.T);
    int b  = (int)(p / d.T);
    const int c0 = g << 3;
    F8 acc = zero8();
    for (int kt = 0; kt < d.kT; ++kt) {
        int tn = ti + d.pT - kt;
        if (tn < 0 || tn % d.sT) continue;
        int to = tn / d.sT;
        if (to >= d.To) continue;
        for (int kh = 0; kh < d.kH; ++kh) {
            int hn = hi + d.pH - kh;
            if (hn < 0 || hn % d.sH) continue;
            int ho = hn / d.sH;
            if (ho >= d.Ho) continue;
            for (int kw = 0; kw < d.kW; ++kw) {
                int wn = wi + d.pW - kw;
                if (wn < 0 || wn % d.sW) continue;
                int wo = wn / d.sW;
                if (wo >= d.Wo) continue;
                F8 gv = load8(dy + ((((long long)b * d.To + to) * d.Ho + ho) * d.Wo + wo) * d.C + c0);
                const float* wp = w_tc + (long long)((kt * d.kH + kh) * d.kW + kw) * d.C + c0;
                float4 w0 = __ldg(reinterpret_cast<const float4*>(wp));
                float4 w1 = __ldg(reinterpret_cast<const float4*>(wp + 4));
                acc.v[0] = fmaf(gv.v[0], w0.x, acc.v[0]); acc.v[1] = fmaf(gv.v[1], w0.y, acc.v[1]);
                acc.v[2] = fmaf(gv.v[2], w0.z, acc.v[2]); acc.v[3] = fmaf(gv.v[3], w0.w, acc.v[3]);
                acc.v[4] = fmaf(gv.v[4], w1.x, acc.v[4]); acc.v[5] = fmaf(gv.v[5], w1.y, acc.v[5]);
                acc.v[6] = fmaf(gv.v[6], w1.z, acc.v[6]); acc.v[7] = fmaf(gv.v[7], w1.w, acc.v[7]);
            }
        }
    }
    store8(dx + idx * 8, acc);
}

// ---------------------------------------------------------------------------------------------
// wgrad: blockIdx.y selects a chunk of <= TAPC taps; each thread keeps TAPC x 8 fp32 accumulators
// for its channel group while striding over output pixels; CTA-level reduction in shared memory,
// then one fp32 atomicAdd per (tap, channel) per CTA.
// ---------------------------------------------------------------------------------------------
constexpr int TAPC = 5;

template <typename T>
__global__ void __launch_bounds__(256)
dw_wgrad_generic(const T* __restrict__ x, const T* __restrict__ dy, float* __restrict__ dw_tc,
                 DwDims d, long long P) {
    extern __shared__ float sm[];   // [TAPC][C]
    const int G = d.C >> 3;
    const int RPI = blockDim.x / G;
    const int g = threadIdx.x % G, rr = threadIdx.x / G;
    const int taps = d.kT * d.kH * d.kW;
    const int tap0 = blockIdx.y * TAPC;
    const int ntap = min(TAPC, taps - tap0);
    for (int i = threadIdx.x; i < TAPC * d.C; i += blockDim.x) sm[i] = 0.f;
    __syncthreads();
    float acc[TAPC][8];
#pragma unroll
    for (int t = 0; t < TAPC; ++t)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[t][i] = 0.f;
    const int c0 = g << 3;
    // decode the tap chunk once
    int kts[TAPC], khs[TAPC], kws[TAPC];
#pragma unroll
    for (int t = 0; t < TAPC; ++t) {
        int tap = min(tap0 + t, taps - 1);
        kws[t] = tap % d.kW; khs[t] = (tap / d.kW) % d.kH; kts[t] = tap / (d.kW * d.kH);
    }
    if (rr < RPI) {
        for (long long p = (long long)blockIdx.x * RPI + rr; p < P; p += (long long)gridDim.x * RPI) {
            long long q = p;
            int wo = (int)(q % d.Wo); q /= d.Wo;
            int ho = (int)(q % d.Ho); q /= d.Ho;
            int to = (int)(q % d.To);
            int b  = (int)(q / d.To);
            F8 gv = load8(dy + p * d.C + c0);
#pragma unroll
            for (int t = 0; t < TAPC; ++t) {
                if (t >= ntap) break;
                int ti = to * d.sT - d.pT + kts[t];
                int hi = ho * d.sH - d.pH + khs[t];
                int wi = wo * d.sW - d.pW + kws[t];
                if (ti < 0 || ti >= d.T || hi < 0 || hi >= d.H || wi < 0 || wi >= d.W) continue;
                F8 xv = load8(x + ((((long long)b * d.T + ti) * d.H + hi) * d.W + wi) * d.C + c0);
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[t][i] = fmaf(xv.v[i], gv.v[i], acc[t][i]);
            }
        }
#pragma unroll
        for (int t = 0; t < TAPC; ++t) {
            if (t >= ntap) break;
#pragma unroll
            for (int i = 0; i < 8; ++i) atomicAdd(&sm[t * d.C + c0 + i], acc[t][i]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < ntap * d.C; i += blockDim.x)
        atomicAdd(&dw_tc[(long long)tap0 * d.C + i], sm[i]);
}

// copy the last kT-1 frames of concat(stream_buf, x) into stream_buf_out
template <typename T>
__global__ void stream_tail_copy(const T* __restrict__ x, const T* __restrict__ sbuf, T* __restrict__ out,
                                 int B, int T_, int keep, long long frame_elems8) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)B * keep * frame_elems8;
    if (idx >= total) return;
    long long e = idx % frame_elems8;
    long long q = idx / frame_elems8;
    int f = (int)(q % keep);
    int b = (int)(q / keep);
    int src = T_ + f - keep;   // frame index in x if >= 0, else frame (keep + src) of the old buffer
    const T* s = src >= 0 ? x + ((long long)b * T_ + src) * frame_elems8 * 8
                          : sbuf + ((long long)b * keep + (keep + src)) * frame_elems8 * 8;
    store8(out + ((long long)b * keep + f) * frame_elems8 * 8 + e * 8, load8(s + e * 8));
}

static int check_dims(const DwDims& d) {
    PB_REQUIRE(d.B > 0 && d.C > 0 && d.T > 0 && d.H > 0 && d.W > 0, "dwconv3d: empty tensor");
    PB_REQUIRE(d.C % 8 == 0, "dwconv3d: C=%d must be a multiple of 8", d.C);
    PB_REQUIRE(d.kT > 0 && d.kH > 0 && d.kW > 0 && d.sT > 0 && d.sH > 0 && d.sW > 0, "dwconv3d: bad kernel/stride");
    PB_REQUIRE(d.To == (d.T + 2 * d.pT - d.kT) / d.sT + 1 && d.Ho == (d.H + 2 * d.pH - d.kH) / d.sH + 1 &&
               d.Wo == (d.W + 2 * d.pW - d.kW) / d.sW + 1, "dwconv3d: output dims (%d,%d,%d) inconsistent", d.To, d.Ho, d.Wo);
    return PB_OK;
}

}  // namespace pb

using namespace pb;

#define DW_ARGS int B, int C, int T_, int H, int W, int kT, int kH, int kW, int sT, int sH, int sW, \
                int pT, int pH, int pW, int To, int Ho, int Wo
#define DW_PACK DwDims d{B, C, T_, H, W, kT, kH, kW, sT, sH, sW, pT, pH, pW, To, Ho, Wo}

extern "C" int pb_dwconv3d_fwd(const void* x, const float* w_tc, void* y, int dtype, DW_ARGS, pb_stream_t stream) {
    DW_PACK;
    if (int e = check_dims(d)) return e;
    cudaStream_t st = (cudaStream_t)stream;
    PB_DISPATCH_DTYPE(dtype, {
        if (dw_fwd_tiled<T>((const T*)x, w_tc, (T*)y, d, st)) { PB_CHECK_LAUNCH("dw_fwd_tiled"); count_path(PB_PATH_DW_FWD_TMA); return PB_OK; }
        long long total = (long long)B * To * Ho * Wo * (C / 8);
        dw_fwd_generic<T, false><<<ceil_div(total, 256), 256, 0, st>>>((const T*)x, nullptr, w_tc, (T*)y, d, total);
    });
    PB_CHECK_LAUNCH("dw_fwd_generic");
    count_path(PB_PATH_DW_FWD_GENERIC);
    return PB_OK;
}

extern "C" int pb_pool_fwd(const void* x, int dtype, int B, long long R, int C, float* mean, pb_stream_t stream);

// Forward + global average pool of the output in one pass (squeeze-excite blocks, mobilenet.py:84-88): the strip
// kernels add the rounded outputs into pool[B][C] while they store them; every other path pools in a second pass.
extern "C" int pb_dwconv3d_fwd_pool(const void* x, const float* w_tc, void* y, float* pool, int dtype, DW_ARGS,
                                    pb_stream_t stream) {
    DW_PACK;
    if (int e = check_dims(d)) return e;
    PB_REQUIRE(pool != nullptr, "dwconv3d_fwd_pool: null pool pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == PB_BF16 && !getenv("PB_DW_POOL_SEPARATE")) {
        PB_CUDA(cudaMemsetAsync(pool, 0, sizeof(float) * (size_t)B * C, st));
        if (dw_fwd_tiled<__nv_bfloat16>((const __nv_bfloat16*)x, w_tc, (__nv_bfloat16*)y, d, st, pool)) {
            PB_CHECK_LAUNCH("dw_fwd_tiled(pool)");
            count_path(PB_PATH_DW_FWD_TMA);
            return PB_OK;
        }
    }
    if (int e = pb_dwconv3d_fwd(x, w_tc, y, dtype, B, C, T_, H, W, kT, kH, kW, sT, sH, sW, pT, pH, pW, To, Ho, Wo, stream)) return e;
    return pb_pool_fwd(y, dtype, B, (long long)To * Ho * Wo, C, pool, stream);
}

extern "C" int pb_dwconv3d_dgrad(const void* dy, const float* w_tc, void* dx, int dtype, DW_ARGS, pb_stream_t stream) {
    DW_PACK;
    if (int e = check_dims(d)) return e;
    cudaStream_t st = (cudaStream_t)stream;
    PB_DISPATCH_DTYPE(dtype, {
        if (dw_dgrad_tiled<T>((const T*)dy, w_tc, (T*)dx, d, st)) { PB_CHECK_LAUNCH("dw_dgrad_tiled"); count_path(PB_PATH_DW_DGRAD_TMA); return PB_OK; }
        long long total = (long long)B * T_ * H * W * (C / 8);
        dw_dgrad_generic<T><<<ceil_div(total, 256), 256, 0, st>>>((const T*)dy, w_tc, (T*)dx, d, total);
    });
    PB_CHECK_LAUNCH("dw_dgrad_generic");
    count_path(PB_PATH_DW_DGRAD_GENERIC);
    return PB_OK;
}

extern "C" int pb_dwconv3d_wgrad(const void* x, const void* dy, float* dw_tc, int dtype, DW_ARGS, pb_stream_t stream) {
    DW_PACK;
    if (int e = check_dims(d)) return e;
    PB_REQUIRE(C / 8 <= 256, "dwconv3d_wgrad: C=%d too large", C);
    cudaStream_t st = (cudaStream_t)stream;
    const int taps = kT * kH * kW;
    PB_CUDA(cudaMemsetAsync(dw_tc, 0, sizeof(float) * (size_t)taps * C, st));
    PB_DISPATCH_DTYPE(dtype, {
        if (dw_wgrad_tiled<T>((const T*)x, (const T*)dy, dw_tc, d, st)) { PB_CHECK_LAUNCH("dw_wgrad_tiled"); count_path(PB_PATH_DW_WGRAD_TMA); return PB_OK; }
        long long P = (long long)B * To * Ho * Wo;
        int G = C / 8, RPI = 256 / G;
        int gx = (int)std::min<long long>(ceil_div(P, RPI), 148 * 8);
        dim3 grid(gx, ceil_div(taps, TAPC));
        dw_wgrad_generic<T><<<grid, 256, sizeof(float) * TAPC * C, st>>>((const T*)x, (const T*)dy, dw_tc, d, P);
    });
    PB_CHECK_LAUNCH("dw_wgrad_generic");
    count_path(PB_PATH_DW_WGRAD_GENERIC);
    return PB_OK;
}

extern "C" int pb_stream_dwconv3d_fwd(const void* x, const void* stream_buf, const float* w_tc, void* y,
                                      void* stream_buf_out, int dtype, int B, int C, int T_, int H, int W,
                                      int kT, int kH, int kW, int sH, int sW, int pH, int pW, int Ho, int Wo,
                                      pb_stream_t stream) {
    DwDims d{B, C, T_, H, W, kT, kH, kW, 1, sH, sW, kT - 1, pH, pW, T_, Ho, Wo};
    PB_REQUIRE(B > 0 && C > 0 && C % 8 == 0 && T_ > 0, "stream_dwconv3d: bad dims");
    PB_REQUIRE(Ho == (H + 2 * pH - kH) / sH + 1 && Wo == (W + 2 * pW - kW) / sW + 1, "stream_dwconv3d: bad output dims");
    PB_REQUIRE(kT == 1 || (stream_buf && stream_buf_out), "stream_dwconv3d: stream buffers required for kT>1");
    PB_REQUIRE(stream_buf_out != stream_buf || kT - 1 <= T_, "stream_dwconv3d: in-place tail update needs T >= kT-1");
    cudaStream_t st = (cudaStream_t)stream;
    bool tiled = false;
    if (dtype == PB_BF16 && kT == 1) {      // no temporal taps: the stateless (1,k,k) kernels serve the chunk
        DwDims d1 = d;
        d1.pT = 0;
        tiled = dw_fwd_tiled<__nv_bfloat16>((const __nv_bfloat16*)x, w_tc, (__nv_bfloat16*)y, d1, st);
    }
    if (dtype == PB_BF16 && kT > 1)
        tiled = dw_stream_fwd_tiled((const __nv_bfloat16*)x, (const __nv_bfloat16*)stream_buf, w_tc, (__nv_bfloat16*)y, d, st);
    if (tiled) { PB_CHECK_LAUNCH("dw_stream_fwd_tiled"); count_path(PB_PATH_DW_STREAM_TMA); }
    PB_DISPATCH_DTYPE(dtype, {
        if (!tiled) {
            long long total = (long long)B * T_ * Ho * Wo * (C / 8);
            dw_fwd_generic<T, true><<<ceil_div(total, 256), 256, 0, st>>>((const T*)x, (const T*)stream_buf, w_tc, (T*)y, d, total);
            PB_CHECK_LAUNCH("dw_fwd_stream");
            count_path(PB_PATH_DW_STREAM_GENERIC);
        }
        if (kT > 1) {
            long long fe8 = (long long)H * W * (C / 8);
            long long n = (long long)B * (kT - 1) * fe8;
            stream_tail_copy<T><<<ceil_div(n, 256), 256, 0, st>>>((const T*)x, (const T*)stream_buf, (T*)stream_buf_out, B, T_, kT - 1, fe8);
            PB_CHECK_LAUNCH("stream_tail_copy");
        }
    });
    return PB_OK;
}
