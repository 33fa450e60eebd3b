// Stem convolution on the tensor cores: implicit im2col built by the threads in shared memory + tcgen05.
//
//   forward : Y[p][co]  = sum_k Xcol[p][k] * W[co][k] + bias      (M = 128 pixels per MMA, N = 16, K = 96)
//   wgrad   : dW[k][co] = sum_p Xcol[p][k] * dY[p][co]            (M = 128 rows of k, N = 16, K = pixels)
//
// Xcol[p][k] (k = ((kt*3+kh)*3+kw)*3+ci, 81 real columns for the 3x3x3 stem, 27 for MoViNet's 1x3x3) is never
// materialised in HBM: every thread gathers the 9 contiguous values of each (kt,kh) patch row of its pixel from
// the channels-last clip (uint8 / bf16 / fp32; uint8 folds train.py:106's `/255`), packs them to bf16 and
// writes 16-byte chunks in the canonical NO-swizzle core-matrix layout
//        offset(p, k) = (p/8)*2048 + (k/8)*128 + (p%8)*16 + (k%8)*2
// which tcgen05 reads K-major for the forward product and -- the very same bytes -- MN-major for the weight
// gradient (there the reduction runs over pixels).  Column 81 of the wgrad tile is a column of ones, so row
// 81 of the accumulator is the bias gradient.  Accumulators live in TMEM; the forward double-buffers them so
// the epilogue of one 256-pixel step overlaps the MMAs of the next; the weight gradient keeps one accumulator
// for the whole kernel and adds it to dW with one atomic per weight and CTA.
// Replaces block1.0 of the three models (mobilenet.py:141,221; movinet.py:92) for bf16 activations and
// channels-last input; everything else goes to the direct kernels in stem.cu.
#include <algorithm>

#include "stem_tc.cuh"
#include "tc_common.cuh"

namespace pb {

using namespace tc;

constexpr int STC_KPAD = 128;                 // k columns per pixel row in the tile (16 chunks of 8)
constexpr int STC_GROUP_BYTES = 16 * 128;     // 8 pixels x 16 chunks x 16 B
constexpr int STC_TILE_BYTES = 16 * STC_GROUP_BYTES;   // 128 pixels = 32 KB

template <typename TX> struct StcLoad;
template <> struct StcLoad<unsigned char> {
    static constexpr bool IS_U8 = true;
    static __device__ __forceinline__ float get(const unsigned char* p) { return (float)(*p) * (1.0f / 255.0f); }
};
template <> struct StcLoad<__nv_bfloat16> {
    static constexpr bool IS_U8 = false;
    static __device__ __forceinline__ float get(const __nv_bfloat16* p) { return __bfloat162float(*p); }
};
template <> struct StcLoad<float> {
    static constexpr bool IS_U8 = false;
    static __device__ __forceinline__ float get(const float* p) { return *p; }
};

__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// no-swizzle descriptor: LBO / SBO in bytes
__device__ __forceinline__ uint64_t desc_ns(uint32_t addr, uint32_t lbo, uint32_t sbo) { return make_desc(addr, lbo, sbo, 0u); }

// Gather the im2col row of output pixel `p` into the tile row `row` (0..127) at `tile` (shared address).
// ONES: also write 1.0 into column NK (bias gradient trick).
template <typename TX, int KT, bool ONES>
__device__ __forceinline__ void gather_row(const TX* __restrict__ x, const StemTc& d, long long p, bool valid,
                                           uint32_t tile, int row) {
    constexpr int NK = KT * 3 * 9;                         // 81 or 27 real columns
    constexpr int NCH = ONES ? 16 : (NK + 15) / 16 * 2;    // chunks to write (wgrad: all 16, fwd: K rounded to 16)
    const uint32_t base = tile + (uint32_t)(row >> 3) * STC_GROUP_BYTES + (uint32_t)(row & 7) * 16;
    const TX* rowp[KT * 3];
    unsigned rowmask = 0, colmask = 0;
    if (valid) {
        long long q = p;
        const int wo = (int)(q % d.Wo); q /= d.Wo;
        const int ho = (int)(q % d.Ho); q /= d.Ho;
        const int to = (int)(q % d.To);
        const int b = (int)(q / d.To);
        const int t0 = to * d.sT - d.pT, h0 = ho * d.sH - d.pH, w0 = wo * d.sW - d.pW;
        const TX* xb = x + (long long)b * d.xs_b + (long long)w0 * 3;
#pragma unroll
        for (int kt = 0; kt < KT; ++kt)
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                const int ti = t0 + kt, hi = h0 + kh;
                const bool ok = ti >= 0 && ti < d.T && hi >= 0 && hi < d.H;
                rowp[kt * 3 + kh] = xb + (long long)(ok ? ti : 0) * d.xs_t + (long long)(ok ? hi : 0) * d.xs_h;
                rowmask |= (ok ? 1u : 0u) << (kt * 3 + kh);
            }
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) colmask |= ((w0 + kw >= 0 && w0 + kw < d.W) ? 1u : 0u) << kw;
    } else {
#pragma unroll
        for (int i = 0; i < KT * 3; ++i) rowp[i] = x;
    }
    const bool interior = rowmask == ((1u << (KT * 3)) - 1u) && colmask == 7u;
    uint32_t cur[4];
    float prev = 0.f;
    const bool all_interior = __all_sync(0xffffffffu, interior);
    // uint8 clip, word path: taken unless a lane's patch sticks out on the RIGHT (kw = 1 or 2 invalid: reading the
    // 9 bytes could then run past the clip's end).  Left padding (w0 = -1), padded rows / frames and the invalid
    // pixels of a partial tile are handled by masks, so the ~30 % of warps that touch a border do not fall back to
    // 81 byte loads + I2F per pixel.
    const bool word_ok = !valid || (colmask & 6u) == 6u;
    if (StcLoad<TX>::IS_U8 && __all_sync(0xffffffffu, word_ok)) {
        // The 9 bytes of a patch row come from three aligned 32-bit loads; byte -> float exactly through the 2^23
        // mantissa trick, then the same * (1/255) as the scalar path, so both paths produce identical bf16 values.
        // The last word read may extend up to 3 bytes past the patch row but never past the 4-byte word holding
        // the clip's last byte.  With left padding the row is read from its first real pixel and shifted by 3 bytes.
        const bool left_pad = valid && !(colmask & 1u);
        const uint32_t m0 = left_pad ? 0xFF000000u : 0xFFFFFFFFu;          // bytes 0..2 belong to kw = 0
        uint32_t by[KT * 3][3];
#pragma unroll
        for (int r = 0; r < KT * 3; ++r) {
            const uintptr_t pa = reinterpret_cast<uintptr_t>(rowp[r]) + (left_pad ? 3 : 0);
            const uint32_t* wp = reinterpret_cast<const uint32_t*>(pa & ~uintptr_t(3));
            const unsigned sh = (unsigned)(pa & 3) * 8;
            const uint32_t w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = __ldg(wp + 2);
            uint32_t b0 = __funnelshift_r(w0, w1, sh), b1 = __funnelshift_r(w1, w2, sh), b2 = w2 >> sh;
            if (left_pad) {                                    // shift the 9-byte string up by 3 bytes
                b2 = b1 >> 8;
                b1 = __funnelshift_l(b0, b1, 24);
                b0 = b0 << 24;
            }
            const bool rok = (rowmask >> r) & 1u;
            by[r][0] = rok ? (b0 & m0) : 0u;
            by[r][1] = rok ? b1 : 0u;
            by[r][2] = rok ? b2 : 0u;
        }
#pragma unroll
        for (int k = 0; k < NCH * 8; ++k) {
            float v = 0.f;
            if (k < NK) {
                const uint32_t m = __byte_perm(by[k / 9][(k % 9) >> 2], 0x4B000000u, 0x7540u + (uint32_t)((k % 9) & 3));
                v = (__uint_as_float(m) - 8388608.f) * (1.0f / 255.0f);
            } else if (ONES && k == NK) {
                v = valid ? 1.f : 0.f;
            }
            if (k & 1) cur[(k & 7) >> 1] = pack_bf16x2(prev, v); else prev = v;
            if ((k & 7) == 7) sts128(base + (uint32_t)(k >> 3) * 128, cur[0], cur[1], cur[2], cur[3]);
        }
    } else if (all_interior) {
#pragma unroll
        for (int k = 0; k < NCH * 8; ++k) {
            float v = 0.f;
            if (k < NK) v = StcLoad<TX>::get(rowp[k / 9] + (k % 9));
            else if (ONES && k == NK) v = 1.f;
            if (k & 1) cur[(k & 7) >> 1] = pack_bf16x2(prev, v); else prev = v;
            if ((k & 7) == 7) sts128(base + (uint32_t)(k >> 3) * 128, cur[0], cur[1], cur[2], cur[3]);
        }
    } else {
#pragma unroll
        for (int k = 0; k < NCH * 8; ++k) {
            float v = 0.f;
            if (k < NK) {
                const bool ok = ((rowmask >> (k / 9)) & 1u) && ((colmask >> ((k % 9) / 3)) & 1u);
                if (ok) v = StcLoad<TX>::get(rowp[k / 9] + (k % 9));
            } else if (ONES && k == NK) {
                v = valid ? 1.f : 0.f;
            }
            if (k & 1) cur[(k & 7) >> 1] = pack_bf16x2(prev, v); else prev = v;
            if ((k & 7) == 7) sts128(base + (uint32_t)(k >> 3) * 128, cur[0], cur[1], cur[2], cur[3]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <typename TX, int KT>
__global__ void __launch_bounds__(256, 2)
stem_tc_fwd_kernel(const TX* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                   __nv_bfloat16* __restrict__ y, const StemTc d) {
    constexpr int NK = KT * 27;
    constexpr int KSTEPS = (NK + 15) / 16;                 // 6 (81 -> 96) or 2 (27 -> 32)
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t mma_bar[2];
    __shared__ uint32_t tmem_base_s;
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t a_tiles = (raw + 1023u) & ~1023u;       // 2 x 32 KB im2col tiles
    const uint32_t w_tile = a_tiles + 2 * STC_TILE_BYTES;  // weights: 2 groups of 8 co x 16 chunks x 16 B = 4 KB
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) { mbar_init(&mma_bar[0], 1); mbar_init(&mma_bar[1], 1); fence_barrier_init(); }
    if (warp == 1) tmem_alloc(&tmem_base_s, 64);
    // weights -> bf16 core-matrix layout: offset(co,k) = (co/8)*2048 + (k/8)*128 + (co%8)*16 + (k%8)*2
    for (int i = tid; i < 16 * STC_KPAD; i += 256) {
        const int co = i / STC_KPAD, k = i % STC_KPAD;
        float v = 0.f;
        if (k < NK) { const int ci = k % 3, tap = k / 3; v = w[((long long)co * 3 + ci) * (KT * 9) + tap]; }
        const uint32_t off = (uint32_t)(co >> 3) * STC_GROUP_BYTES + (uint32_t)(k >> 3) * 128 + (uint32_t)(co & 7) * 16 + (uint32_t)(k & 7) * 2;
        asm volatile("st.shared.u16 [%0], %1;" ::"r"(w_tile + off), "h"(__bfloat16_as_ushort(__float2bfloat16_rn(v))) : "memory");
    }
    fence_proxy_async();                                   // weights are read by the tensor core (async proxy)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t idesc = make_idesc(128, 16, 0, 0);
    float bv[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) bv[j] = bias ? __ldg(bias + j) : 0.f;

    const int tile_sel = tid >> 7, row = tid & 127;        // threads 0-127 -> tile 0, 128-255 -> tile 1
    long long it = 0;
    for (long long s = blockIdx.x; ; s += gridDim.x, ++it) {
        const bool have = s < d.steps;
        const int st = (int)(it & 1);
        if (it >= 1) {
            // the MMAs of the previous step must be complete before their tiles are overwritten
            mbar_wait(&mma_bar[st ^ 1], (uint32_t)(((it - 1) >> 1) & 1));
            tc_fence_after();
        }
        if (have) {
            const long long p = s * 256 + tid;
            gather_row<TX, KT, false>(x, d, p, p < d.P, a_tiles + (uint32_t)tile_sel * STC_TILE_BYTES, row);
            fence_proxy_async();
        }
        __syncthreads();
        if (have && tid == 0) {
            tc_fence_after();
#pragma unroll
            for (int t2 = 0; t2 < 2; ++t2) {
                const uint32_t a0 = a_tiles + (uint32_t)t2 * STC_TILE_BYTES;
#pragma unroll
                for (int ks = 0; ks < KSTEPS; ++ks)      // 16 k = 2 chunks = 256 B; K-major: LBO = chunk pitch, SBO = group pitch
                    umma_bf16(tmem_base + (uint32_t)(st * 32 + t2 * 16), desc_ns(a0 + ks * 256, 128, STC_GROUP_BYTES),
                              desc_ns(w_tile + ks * 256, 128, STC_GROUP_BYTES), idesc, ks != 0);
            }
            umma_commit(&mma_bar[st]);
        }
        // epilogue of the previous step (its accumulator stage was waited for above)
        if (it >= 1) {
            const long long sp = s - gridDim.x;
            const int pst = st ^ 1;
            const long long p = sp * 256 + (long long)(warp >> 2) * 128 + (warp & 3) * 32 + lane;
            uint32_t r[16];
            tmem_ld16(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(pst * 32 + (warp >> 2) * 16), r);
            tmem_ld_wait();
            if (p < d.P) {
                uint4 o0, o1;
                o0.x = pack_bf16x2(__uint_as_float(r[0]) + bv[0], __uint_as_float(r[1]) + bv[1]);
                o0.y = pack_bf16x2(__uint_as_float(r[2]) + bv[2], __uint_as_float(r[3]) + bv[3]);
                o0.z = pack_bf16x2(__uint_as_float(r[4]) + bv[4], __uint_as_float(r[5]) + bv[5]);
                o0.w = pack_bf16x2(__uint_as_float(r[6]) + bv[6], __uint_as_float(r[7]) + bv[7]);
                o1.x = pack_bf16x2(__uint_as_float(r[8]) + bv[8], __uint_as_float(r[9]) + bv[9]);
                o1.y = pack_bf16x2(__uint_as_float(r[10]) + bv[10], __uint_as_float(r[11]) + bv[11]);
                o1.z = pack_bf16x2(__uint_as_float(r[12]) + bv[12], __uint_as_float(r[13]) + bv[13]);
                o1.w = pack_bf16x2(__uint_as_float(r[14]) + bv[14], __uint_as_float(r[15]) + bv[15]);
                uint4* dst = reinterpret_cast<uint4*>(y + p * 16);
                dst[0] = o0; dst[1] = o1;
            }
            tc_fence_before();
        }
        if (!have) break;
    }
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 64); }
}

// ------------------------------------------------------------------------------------------------
// weight (and bias) gradient
// ------------------------------------------------------------------------------------------------
template <typename TX, int KT>
__global__ void __launch_bounds__(256, 2)
stem_tc_wgrad_kernel(const TX* __restrict__ x, const __nv_bfloat16* __restrict__ dy, float* __restrict__ dw,
                     float* __restrict__ dbias, const StemTc d) {
    constexpr int NK = KT * 27;
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t mma_bar;
    __shared__ uint32_t tmem_base_s;
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t a_tiles = (raw + 1023u) & ~1023u;       // 2 x 32 KB im2col tiles
    const uint32_t g_tiles = a_tiles + 2 * STC_TILE_BYTES; // 2 x 4 KB dy tiles: offset(p,co) = (p/8)*256 + (co/8)*128 + (p%8)*16
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) { mbar_init(&mma_bar, 1); fence_barrier_init(); }
    if (warp == 1) tmem_alloc(&tmem_base_s, 32);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t idesc = make_idesc(128, 16, 1, 1);       // both operands MN-major

    const int tile_sel = tid >> 7, row = tid & 127;
    long long it = 0;
    for (long long s = blockIdx.x; s < d.steps; s += gridDim.x, ++it) {
        if (it >= 1) {                                      // previous MMAs done -> tiles may be overwritten
            mbar_wait(&mma_bar, (uint32_t)((it - 1) & 1));
            tc_fence_after();
        }
        const long long p = s * 256 + tid;
        const bool valid = p < d.P;
        gather_row<TX, KT, true>(x, d, p, valid, a_tiles + (uint32_t)tile_sel * STC_TILE_BYTES, row);
        {
            uint4 g0 = make_uint4(0u, 0u, 0u, 0u), g1 = g0;
            if (valid) {
                const uint4* src = reinterpret_cast<const uint4*>(dy + p * 16);
                g0 = __ldg(src); g1 = __ldg(src + 1);
            }
            const uint32_t gb = g_tiles + (uint32_t)tile_sel * 4096 + (uint32_t)(row >> 3) * 256 + (uint32_t)(row & 7) * 16;
            sts128(gb, g0.x, g0.y, g0.z, g0.w);
            sts128(gb + 128, g1.x, g1.y, g1.z, g1.w);
        }
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
#pragma unroll
            for (int t2 = 0; t2 < 2; ++t2) {
                const uint32_t a0 = a_tiles + (uint32_t)t2 * STC_TILE_BYTES;
                const uint32_t b0 = g_tiles + (uint32_t)t2 * 4096;
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)            // 16 pixels = 2 groups; MN-major: SBO = chunk pitch (MN), LBO = group pitch (K)
                    umma_bf16(tmem_base, desc_ns(a0 + ks * 2 * STC_GROUP_BYTES, STC_GROUP_BYTES, 128),
                              desc_ns(b0 + ks * 512, 256, 128), idesc, (it | t2 | ks) != 0);
            }
            umma_commit(&mma_bar);
        }
    }
    // drain: D[k][co], k = TMEM lane
    if (it >= 1) {
        mbar_wait(&mma_bar, (uint32_t)((it - 1) & 1));
        tc_fence_after();
        if (warp < 4) {
            uint32_t r[16];
            tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16), r);
            tmem_ld_wait();
            const int k = warp * 32 + lane;
            if (k < NK) {
                const int ci = k % 3, tap = k / 3;
#pragma unroll
                for (int co = 0; co < 16; ++co)
                    atomicAdd(&dw[((long long)co * 3 + ci) * (KT * 9) + tap], __uint_as_float(r[co]));
            } else if (k == NK && dbias) {
#pragma unroll
                for (int co = 0; co < 16; ++co) atomicAdd(&dbias[co], __uint_as_float(r[co]));
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 32); }
}

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------
template <typename K>
static bool stc_attr(K kernel) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024) == cudaSuccess;
}

static int stc_grid(const StemTc& d) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return (int)std::min<long long>(d.steps, 2LL * sms);      // two CTAs per SM hide the gather latency
}

// Both return true if they launched; false = shape not covered (caller uses the direct kernels).
bool stem_tc_fwd(const void* x, int x_dtype, const float* w, const float* bias, void* y, int kT, const StemTc& d,
                 cudaStream_t st) {
    const size_t smem = 2 * STC_TILE_BYTES + 4096 + 1024;
#define STC_LAUNCH_FWD(TX, KT)                                                                         \
    do {                                                                                               \
        if (!stc_attr(stem_tc_fwd_kernel<TX, KT>)) return false;                                       \
        stem_tc_fwd_kernel<TX, KT><<<stc_grid(d), 256, smem, st>>>((const TX*)x, w, bias, (__nv_bfloat16*)y, d); \
        return true;                                                                                   \
    } while (0)
    if (kT == 3) {
        if (x_dtype == PB_U8) STC_LAUNCH_FWD(unsigned char, 3);
        if (x_dtype == PB_BF16) STC_LAUNCH_FWD(__nv_bfloat16, 3);
        if (x_dtype == PB_F32) STC_LAUNCH_FWD(float, 3);
    } else if (kT == 1) {
        if (x_dtype == PB_U8) STC_LAUNCH_FWD(unsigned char, 1);
        if (x_dtype == PB_BF16) STC_LAUNCH_FWD(__nv_bfloat16, 1);
        if (x_dtype == PB_F32) STC_LAUNCH_FWD(float, 1);
    }
#undef STC_LAUNCH_FWD
    return false;
}

bool stem_tc_wgrad(const void* x, int x_dtype, const void* dy, float* dw, float* dbias, int kT, const StemTc& d,
                   cudaStream_t st) {
    const size_t smem = 2 * STC_TILE_BYTES + 2 * 4096 + 1024;
#define STC_LAUNCH_WG(TX, KT)                                                                          \
    do {                                                                                               \
        if (!stc_attr(stem_tc_wgrad_kernel<TX, KT>)) return false;                                     \
        stem_tc_wgrad_kernel<TX, KT><<<stc_grid(d), 256, smem, st>>>((const TX*)x, (const __nv_bfloat16*)dy, dw, dbias, d); \
        return true;                                                                                   \
    } while (0)
    if (kT == 3) {
        if (x_dtype == PB_U8) STC_LAUNCH_WG(unsigned char, 3);
        if (x_dtype == PB_BF16) STC_LAUNCH_WG(__nv_bfloat16, 3);
        if (x_dtype == PB_F32) STC_LAUNCH_WG(float, 3);
    } else if (kT == 1) {
        if (x_dtype == PB_U8) STC_LAUNCH_WG(unsigned char, 1);
        if (x_dtype == PB_BF16) STC_LAUNCH_WG(__nv_bfloat16, 1);
        if (x_dtype == PB_F32) STC_LAUNCH_WG(float, 1);
    }
#undef STC_LAUNCH_WG
    return false;
}

}  // namespace pb
