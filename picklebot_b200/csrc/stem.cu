// Stem: dense Conv3d with a tiny input-channel count (RGB), reading the clip with arbitrary strides
// (the reference hands over a (B,C,T,H,W) view of a (B,T,H,W,C) uint8 batch, train.py:102-108) and
// writing NDHWC.  Replaces block1.0 of MobileNetLarge3D/Small3D (mobilenet.py:141,221) and of
// MoViNetA2 (movinet.py:92).  uint8 input folds the `/255` of train.py:106 through a 256-entry table.
#include <algorithm>

#include "common.cuh"
#include "stem_tc.cuh"

namespace pb {

struct StemDims {
    int B, Cin, T, H, W, Cout;
    int kT, kH, kW, sT, sH, sW, pT, pH, pW;
    int To, Ho, Wo;
    long long xs_b, xs_c, xs_t, xs_h, xs_w;
    float in_scale;
};

template <typename TX> struct InLoader;
template <> struct InLoader<unsigned char> {
    static constexpr bool kLut = true;
    static __device__ __forceinline__ float get(const unsigned char* p, const float* lut) { return lut[*p]; }
};
template <> struct InLoader<float> {
    static constexpr bool kLut = false;
    static __device__ __forceinline__ float get(const float* p, const float*) { return *p; }
};
template <> struct InLoader<__nv_bfloat16> {
    static constexpr bool kLut = false;
    static __device__ __forceinline__ float get(const __nv_bfloat16* p, const float*) { return __bfloat162float(*p); }
};

constexpr int STEM_COUT = 16;

template <typename TX, typename TY>
__global__ void __launch_bounds__(128)
stem_fwd_kernel(const TX* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                TY* __restrict__ y, StemDims d, long long P) {
    extern __shared__ float sm[];           // weights [taps*Cin][16] | lut[256]
    const int taps = d.kT * d.kH * d.kW;
    const int nw = taps * d.Cin * STEM_COUT;
    float* ws = sm;
    float* lut = sm + nw;
    for (int i = threadIdx.x; i < nw; i += blockDim.x) {
        int co = i % STEM_COUT;
        int rest = i / STEM_COUT;
        int ci = rest % d.Cin, tap = rest / d.Cin;
        ws[i] = round_to<TY>(w[((long long)co * d.Cin + ci) * taps + tap]);
    }
    if (InLoader<TX>::kLut)
        for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = round_to<TY>((float)i / d.in_scale);
    __syncthreads();
    long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    long long q = p;
    int wo = (int)(q % d.Wo); q /= d.Wo;
    int ho = (int)(q % d.Ho); q /= d.Ho;
    int to = (int)(q % d.To);
    int b  = (int)(q / d.To);
    float acc[STEM_COUT];
#pragma unroll
    for (int c = 0; c < STEM_COUT; ++c) acc[c] = bias ? bias[c] : 0.f;
    const TX* xb = x + (long long)b * d.xs_b;
    for (int kt = 0; kt < d.kT; ++kt) {
        int ti = to * d.sT - d.pT + kt;
        if (ti < 0 || ti >= d.T) continue;
        for (int kh = 0; kh < d.kH; ++kh) {
            int hi = ho * d.sH - d.pH + kh;
            if (hi < 0 || hi >= d.H) continue;
            for (int kw = 0; kw < d.kW; ++kw) {
                int wi = wo * d.sW - d.pW + kw;
                if (wi < 0 || wi >= d.W) continue;
                const TX* px = xb + ti * d.xs_t + hi * d.xs_h + wi * d.xs_w;
                const int tap = (kt * d.kH + kh) * d.kW + kw;
                for (int ci = 0; ci < d.Cin; ++ci) {
                    float xv = InLoader<TX>::get(px + ci * d.xs_c, lut);
                    if (!InLoader<TX>::kLut) xv = round_to<TY>(xv);
                    const float4* wr = reinterpret_cast<const float4*>(ws + (tap * d.Cin + ci) * STEM_COUT);
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        float4 wv = wr[v];
                        acc[v * 4 + 0] = fmaf(xv, wv.x, acc[v * 4 + 0]);
                        acc[v * 4 + 1] = fmaf(xv, wv.y, acc[v * 4 + 1]);
                        acc[v * 4 + 2] = fmaf(xv, wv.z, acc[v * 4 + 2]);
                        acc[v * 4 + 3] = fmaf(xv, wv.w, acc[v * 4 + 3]);
                    }
                }
            }
        }
    }
    F8 o0, o1;
#pragma unroll
    for (int i = 0; i < 8; ++i) { o0.v[i] = acc[i]; o1.v[i] = acc[8 + i]; }
    store8(y + p * STEM_COUT, o0);
    store8(y + p * STEM_COUT + 8, o1);
}

// wgrad: thread = (one (tap,ci) pair, one of SUBS pixel lanes); 16 accumulators (one per output channel).
constexpr int STEM_TP = 64;

template <typename TX, typename TY>
__global__ void __launch_bounds__(256)
stem_wgrad_kernel(const TX* __restrict__ x, const TY* __restrict__ dy, float* __restrict__ dw,
                  float* __restrict__ dbias, StemDims d, long long P, long long pix_per_cta) {
    __shared__ float dys[STEM_TP][STEM_COUT];
    __shared__ int   pb_[STEM_TP], pt_[STEM_TP], ph_[STEM_TP], pw_[STEM_TP];
    __shared__ float lut[256];
    const int taps = d.kT * d.kH * d.kW;
    const int NP = taps * d.Cin;
    const int SUBS = blockDim.x / NP;
    const int tid = threadIdx.x;
    const int pair = tid % NP, sub = tid / NP;
    const bool active = sub < SUBS;
    const int ci = pair % d.Cin, tap = pair / d.Cin;
    const int kw = tap % d.kW, kh = (tap / d.kW) % d.kH, kt = tap / (d.kW * d.kH);
    if (InLoader<TX>::kLut)
        for (int i = tid; i < 256; i += blockDim.x) lut[i] = round_to<TY>((float)i / d.in_scale);
    float acc[STEM_COUT];
#pragma unroll
    for (int c = 0; c < STEM_COUT; ++c) acc[c] = 0.f;
    float bsum = 0.f;
    const long long p_begin = (long long)blockIdx.x * pix_per_cta;
    const long long p_end = min(P, p_begin + pix_per_cta);
    for (long long p0 = p_begin; p0 < p_end; p0 += STEM_TP) {
        __syncthreads();
        for (int i = tid; i < STEM_TP * STEM_COUT; i += blockDim.x) {
            int pi = i / STEM_COUT, c = i % STEM_COUT;
            long long p = p0 + pi;
            dys[pi][c] = p < p_end ? to_float(dy[p * STEM_COUT + c]) : 0.f;
        }
        if (tid < STEM_TP) {
            long long q = p0 + tid;
            if (q < p_end) {
                int wo = (int)(q % d.Wo); q /= d.Wo;
                int ho = (int)(q % d.Ho); q /= d.Ho;
                int to = (int)(q % d.To);
                pb_[tid] = (int)(q / d.To);
                pt_[tid] = to * d.sT - d.pT; ph_[tid] = ho * d.sH - d.pH; pw_[tid] = wo * d.sW - d.pW;
            } else {
                pb_[tid] = -1;
            }
        }
        __syncthreads();
        if (active) {
            for (int pi = sub; pi < STEM_TP; pi += SUBS) {
                int b = pb_[pi];
                if (b < 0) continue;
                int ti = pt_[pi] + kt, hi = ph_[pi] + kh, wi = pw_[pi] + kw;
                if (ti < 0 || ti >= d.T || hi < 0 || hi >= d.H || wi < 0 || wi >= d.W) continue;
                float xv = InLoader<TX>::get(x + b * d.xs_b + ci * d.xs_c + ti * d.xs_t + hi * d.xs_h + wi * d.xs_w, lut);
                if (!InLoader<TX>::kLut) xv = round_to<TY>(xv);
                const float4* g4 = reinterpret_cast<const float4*>(&dys[pi][0]);
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    float4 g = g4[v];
                    acc[v * 4 + 0] = fmaf(xv, g.x, acc[v * 4 + 0]);
                    acc[v * 4 + 1] = fmaf(xv, g.y, acc[v * 4 + 1]);
                    acc[v * 4 + 2] = fmaf(xv, g.z, acc[v * 4 + 2]);
                    acc[v * 4 + 3] = fmaf(xv, g.w, acc[v * 4 + 3]);
                }
            }
        }
        if (dbias && tid < STEM_COUT) {
            for (int pi = 0; pi < STEM_TP; ++pi) bsum += dys[pi][tid];
        }
    }
    if (active) {
#pragma unroll
        for (int c = 0; c < STEM_COUT; ++c)
            atomicAdd(&dw[((long long)c * d.Cin + ci) * taps + tap], acc[c]);
    }
    if (dbias && tid < STEM_COUT) atomicAdd(&dbias[tid], bsum);
}

static int stem_check(const StemDims& d) {
    PB_REQUIRE(d.Cout == STEM_COUT, "stem: Cout=%d unsupported (kernel is specialised for 16)", d.Cout);
    PB_REQUIRE(d.B > 0 && d.Cin > 0 && d.Cin <= 4 && d.T > 0 && d.H > 0 && d.W > 0, "stem: bad input dims");
    PB_REQUIRE(d.kT * d.kH * d.kW * d.Cin <= 128, "stem: kernel too large");
    PB_REQUIRE(d.To == (d.T + 2 * d.pT - d.kT) / d.sT + 1 && d.Ho == (d.H + 2 * d.pH - d.kH) / d.sH + 1 &&
               d.Wo == (d.W + 2 * d.pW - d.kW) / d.sW + 1, "stem: output dims inconsistent");
    PB_REQUIRE(d.in_scale != 0.f, "stem: in_scale (divisor) must be non-zero");
    return PB_OK;
}

// tcgen05 path (stem_tc.cu): bf16 activations, RGB channels-last clip, 3x3 spatial kernel, 16 output channels
static bool stem_tc_eligible(const StemDims& d, int y_dtype) {
    return y_dtype == PB_BF16 && d.Cin == 3 && d.Cout == STEM_COUT && d.kH == 3 && d.kW == 3 && (d.kT == 3 || d.kT == 1) &&
           d.xs_c == 1 && d.xs_w == 3 && d.in_scale == 255.f;
}
static StemTc stem_tc_dims(const StemDims& d) {
    StemTc t{d.B, d.T, d.H, d.W, d.To, d.Ho, d.Wo, d.sT, d.sH, d.sW, d.pT, d.pH, d.pW, d.xs_b, d.xs_t, d.xs_h, 0, 0};
    t.P = (long long)d.B * d.To * d.Ho * d.Wo;
    t.steps = (t.P + 255) / 256;
    t.inv_scale = 1.0f / d.in_scale;
    t.act = PB_ACT_NONE; t.slope = 0.f;
    return t;
}

}  // namespace pb

using namespace pb;

#define STEM_ARGS long long xs_b, long long xs_c, long long xs_t, long long xs_h, long long xs_w, float in_scale
#define STEM_DIMS int B, int Cin, int T_, int H, int W, int Cout, int kT, int kH, int kW, int sT, int sH, int sW, \
                  int pT, int pH, int pW, int To, int Ho, int Wo
#define STEM_PACK StemDims d{B, Cin, T_, H, W, Cout, kT, kH, kW, sT, sH, sW, pT, pH, pW, To, Ho, Wo, \
                             xs_b, xs_c, xs_t, xs_h, xs_w, in_scale}

// Dispatch over (input dtype, output dtype); body sees TX, TY.
#define STEM_DISPATCH(x_dtype, y_dtype, ...)                                                        \
    do {                                                                                            \
        if ((y_dtype) == PB_BF16) {                                                                 \
            using TY = __nv_bfloat16;                                                               \
            if ((x_dtype) == PB_U8) { using TX = unsigned char; __VA_ARGS__; }                      \
            else if ((x_dtype) == PB_BF16) { using TX = __nv_bfloat16; __VA_ARGS__; }               \
            else if ((x_dtype) == PB_F32) { using TX = float; __VA_ARGS__; }                        \
            else { set_error("stem: bad x dtype"); return PB_ERR_BAD_ARG; }                         \
        } else if ((y_dtype) == PB_F32) {                                                           \
            using TY = float;                                                                       \
            if ((x_dtype) == PB_U8) { using TX = unsigned char; __VA_ARGS__; }                      \
            else if ((x_dtype) == PB_BF16) { using TX = __nv_bfloat16; __VA_ARGS__; }               \
            else if ((x_dtype) == PB_F32) { using TX = float; __VA_ARGS__; }                        \
            else { set_error("stem: bad x dtype"); return PB_ERR_BAD_ARG; }                         \
        } else { set_error("stem: bad y dtype"); return PB_ERR_BAD_ARG; }                           \
    } while (0)

// Inference form: y = act(conv(x) + bias) with the eval-mode BatchNorm of the stem already folded into w / bias by the
// caller (blocks.py stem_eval).  tcgen05 kernels only (bf16 output); PB_ERR_UNSUPPORTED otherwise.
extern "C" int pb_stem_conv_fwd_act(const void* x, int x_dtype, STEM_ARGS, const float* w, const float* bias, void* y,
                                    int y_dtype, STEM_DIMS, int act, float slope, pb_stream_t stream) {
    STEM_PACK;
    if (int e = stem_check(d)) return e;
    PB_REQUIRE(x && w && y, "stem_conv_fwd_act: null pointer");
    StemTc t = stem_tc_dims(d);
    t.act = act; t.slope = slope;
    if (stem_tc_eligible(d, y_dtype) && stem_tc_fwd(x, x_dtype, w, bias, y, kT, t, (cudaStream_t)stream)) {
        PB_CHECK_LAUNCH("stem_tc_fwd_kernel");
        count_path(PB_PATH_STEM_TC);
        return PB_OK;
    }
    set_error("stem_conv_fwd_act: shape / dtype not covered by the tensor-core stem kernels");
    return PB_ERR_UNSUPPORTED;
}

extern "C" int pb_stem_conv_fwd(const void* x, int x_dtype, STEM_ARGS, const float* w, const float* bias, void* y,
                                int y_dtype, STEM_DIMS, pb_stream_t stream) {
    STEM_PACK;
    if (int e = stem_check(d)) return e;
    PB_REQUIRE(x && w && y, "stem_conv_fwd: null pointer");
    if (stem_tc_eligible(d, y_dtype) && stem_tc_fwd(x, x_dtype, w, bias, y, kT, stem_tc_dims(d), (cudaStream_t)stream)) {
        PB_CHECK_LAUNCH("stem_tc_fwd_kernel");
        count_path(PB_PATH_STEM_TC);
        return PB_OK;
    }
    long long P = (long long)B * To * Ho * Wo;
    size_t smem = sizeof(float) * ((size_t)kT * kH * kW * Cin * STEM_COUT + 256);
    STEM_DISPATCH(x_dtype, y_dtype, {
        stem_fwd_kernel<TX, TY><<<ceil_div(P, 128), 128, smem, (cudaStream_t)stream>>>((const TX*)x, w, bias, (TY*)y, d, P);
    });
    PB_CHECK_LAUNCH("stem_fwd_kernel");
    count_path(PB_PATH_STEM_SIMT);
    return PB_OK;
}

extern "C" int pb_stem_conv_wgrad(const void* x, int x_dtype, STEM_ARGS, const void* dy, int y_dtype, float* dw,
                                  float* dbias, STEM_DIMS, pb_stream_t stream) {
    STEM_PACK;
    if (int e = stem_check(d)) return e;
    PB_REQUIRE(x && dy && dw, "stem_conv_wgrad: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int taps = kT * kH * kW;
    PB_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)Cout * Cin * taps, st));
    if (dbias) PB_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * Cout, st));
    if (stem_tc_eligible(d, y_dtype) && stem_tc_wgrad(x, x_dtype, dy, dw, dbias, kT, stem_tc_dims(d), st)) {
        PB_CHECK_LAUNCH("stem_tc_wgrad_kernel");
        count_path(PB_PATH_STEM_TC);
        return PB_OK;
    }
    long long P = (long long)B * To * Ho * Wo;
    long long per = std::max<long long>(STEM_TP, (P + 148 * 8 - 1) / (148 * 8));
    per = (per + STEM_TP - 1) / STEM_TP * STEM_TP;
    STEM_DISPATCH(x_dtype, y_dtype, {
        stem_wgrad_kernel<TX, TY><<<ceil_div(P, per), 256, 0, st>>>((const TX*)x, (const TY*)dy, dw, dbias, d, P, per);
    });
    PB_CHECK_LAUNCH("stem_wgrad_kernel");
    count_path(PB_PATH_STEM_SIMT);
    return PB_OK;
}
