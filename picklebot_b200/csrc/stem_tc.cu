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
#include <cstdio>
#include <cstdlib>

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
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]) + bv[j];
                act_fwd_vec(v, d.act, d.slope);
                uint4 o0, o1;
                o0.x = pack_bf16x2(v[0], v[1]);   o0.y = pack_bf16x2(v[2], v[3]);
                o0.z = pack_bf16x2(v[4], v[5]);   o0.w = pack_bf16x2(v[6], v[7]);
                o1.x = pack_bf16x2(v[8], v[9]);   o1.y = pack_bf16x2(v[10], v[11]);
                o1.z = pack_bf16x2(v[12], v[13]); o1.w = pack_bf16x2(v[14], v[15]);
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
// uint8 clips through TMA (round 2)
//
// The gather kernels above fetch every pixel's 9 patch rows straight from HBM: ~1 us of latency per 256-pixel
// step with nothing to hide it behind (1.1 TB/s).  Here a producer warp streams the clip rows a step needs --
// RT output rows of one output frame = (RT-1)*sH+3 input rows of KT frames -- into a shared-memory ring with one
// cp.async.bulk.tensor per step (the clip is described as a 4-D tensor of 32-bit words [W*3/4][H][T][B]; rows and
// frames outside the clip arrive as zeros, which IS the convolution's zero padding, so the border masks of the
// gather path disappear), and the 256 gather threads build the im2col rows from shared memory.
//
// Bytes become EXACT integers and the /255 of train.py:106 moves behind the fp32 accumulator, so nothing is rounded
// before the products.  Forward: fp16 integers with one PRMT + one HSUB2 per pair (0x6400 | b = 1024 + b) against
// fp16 weights: 1.1 instructions per value instead of 3.5.  Weight gradient: the upstream gradient is bf16 and
// kind::f16 does NOT take mixed operand formats (fp16 x bf16 faults with "illegal instruction" on B200, measured),
// so there the bytes become bf16 integers through the 2^23 trick with packed fp32 adds: 2.1 instructions per value.  K order per pixel: patch row pr = kt*3+kh -> 10 columns
// (9 values kw*3+ci and a zero), i.e. 5 packed words per patch row; the bias-gradient ones column is k = 10*KT*3.
// ------------------------------------------------------------------------------------------------
struct StemTma {
    int TW, RT, NR;          // pixels per output row (= Wo), output rows per step, staged input rows per frame
    int rowb;                // bytes per staged row (box width * 4)
    int boff;                // byte offset of output column 0's patch inside a staged row
    int e0;                  // first 32-bit word of the box (negative: left padding; a multiple of 4 -- the
                             // innermost TMA coordinate must land on a 16-byte boundary or the load faults)
    int groups;              // row groups per output frame
    int nst;                 // ring stages
    uint32_t xbytes;         // bytes of one clip box
    uint32_t stage_pitch;    // ring pitch (clip box [+ dy block for the weight gradient]), multiple of 128
    long long steps;
};

constexpr int STM_THREADS = 288;              // 8 gather/epilogue warps + 1 producer warp
constexpr int STM_MAX_STAGES = 4;

__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t bytes_to_h2(uint32_t w, uint32_t sel) {      // two bytes -> two exact fp16 integers
    uint32_t r;
    const uint32_t h = __byte_perm(w, 0x00000064u, sel);
    asm("sub.f16x2 %0, %1, %2;" : "=r"(r) : "r"(h), "r"(0x64006400u));
    return r;
}
// four bytes -> four exact bf16 integers: 0x4B0000bb = 2^23 + b as fp32, minus 2^23 (two lanes per add.f32x2),
// and the high halves of the two floats ARE the bf16 values (an 8-bit integer leaves the low 16 bits zero)
__device__ __forceinline__ void bytes_to_bf16x4(uint32_t w, uint32_t& lo, uint32_t& hi) {
    uint32_t f0 = __byte_perm(w, 0x4B000000u, 0x7540u), f1 = __byte_perm(w, 0x4B000000u, 0x7541u);
    uint32_t f2 = __byte_perm(w, 0x4B000000u, 0x7542u), f3 = __byte_perm(w, 0x4B000000u, 0x7543u);
    uint64_t a, b;
    const uint64_t m = 0xCB000000CB000000ull;                  // (-2^23, -2^23)
    asm("{.reg .b64 t; mov.b64 t, {%1, %2}; add.rn.f32x2 %0, t, %3;}" : "=l"(a) : "r"(f0), "r"(f1), "l"(m));
    asm("{.reg .b64 t; mov.b64 t, {%1, %2}; add.rn.f32x2 %0, t, %3;}" : "=l"(b) : "r"(f2), "r"(f3), "l"(m));
    lo = __byte_perm((uint32_t)a, (uint32_t)(a >> 32), 0x7632u);
    hi = __byte_perm((uint32_t)b, (uint32_t)(b >> 32), 0x7632u);
}
__device__ __forceinline__ uint32_t byte_to_bf16(uint32_t w) {  // byte 0 -> {bf16 value, 0}
    const float f = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540u)) - 8388608.f;
    return __float_as_uint(f) >> 16;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// im2col row of the pixel (row r of the step, column wo) from the staged clip rows -> tile row `row`.
// NCH 16-byte chunks are written; ONES puts 1.0 (valid pixel) / 0 at column 10*KT*3.
template <int KT, int NCH, bool ONES, bool BF16>
__device__ __forceinline__ void gather_row_smem(uint32_t stage, const StemTma& g, int sH, int sW, int r, int wo,
                                                bool valid, uint32_t tile, int row) {
    constexpr int NPR = KT * 3;
    constexpr int NW = NPR * 5;
    const uint32_t base = tile + (uint32_t)(row >> 3) * STC_GROUP_BYTES + (uint32_t)(row & 7) * 16;
    const uint32_t off0 = (uint32_t)g.boff + (uint32_t)(wo * sW * 3);
    const uint32_t sh = (off0 & 3u) * 8u;
    const uint32_t src = stage + (uint32_t)(r * sH) * (uint32_t)g.rowb + (off0 & ~3u);
    uint32_t wd[NCH * 4];
#pragma unroll
    for (int pr = 0; pr < NPR; ++pr) {
        const uint32_t a = src + (uint32_t)((pr / 3) * g.NR + (pr % 3)) * (uint32_t)g.rowb;
        const uint32_t w0 = lds32(a), w1 = lds32(a + 4), w2 = lds32(a + 8);
        const uint32_t b0 = __funnelshift_r(w0, w1, sh), b1 = __funnelshift_r(w1, w2, sh), b2 = w2 >> sh;
        if (BF16) {
            bytes_to_bf16x4(b0, wd[pr * 5 + 0], wd[pr * 5 + 1]);
            bytes_to_bf16x4(b1, wd[pr * 5 + 2], wd[pr * 5 + 3]);
            wd[pr * 5 + 4] = byte_to_bf16(b2);
        } else {
            wd[pr * 5 + 0] = bytes_to_h2(b0, 0x4140u);
            wd[pr * 5 + 1] = bytes_to_h2(b0, 0x4342u);
            wd[pr * 5 + 2] = bytes_to_h2(b1, 0x4140u);
            wd[pr * 5 + 3] = bytes_to_h2(b1, 0x4342u);
            wd[pr * 5 + 4] = bytes_to_h2(b2, 0x4540u);         // {value 8, 0}
        }
    }
#pragma unroll
    for (int i = NW; i < NCH * 4; ++i) wd[i] = 0u;
    if (ONES) wd[NW] = valid ? (BF16 ? 0x00003F80u : 0x00003C00u) : 0u;   // 1.0 in the low half
#pragma unroll
    for (int c = 0; c < NCH; ++c) sts128(base + (uint32_t)c * 128, wd[c * 4], wd[c * 4 + 1], wd[c * 4 + 2], wd[c * 4 + 3]);
}

struct StmStep { int b, to, hg; };
__device__ __forceinline__ StmStep stm_decode(uint32_t s, const StemTc& d, const StemTma& g) {   // steps < 2^31 (host)
    StmStep c;
    const uint32_t t2 = s / (uint32_t)g.groups;
    c.hg = (int)(s - t2 * (uint32_t)g.groups);
    c.b = (int)(t2 / (uint32_t)d.To);
    c.to = (int)(t2 - (uint32_t)c.b * (uint32_t)d.To);
    return c;
}

template <int KT>
__global__ void __launch_bounds__(STM_THREADS, 2)
stem_tma_fwd_kernel(const __grid_constant__ CUtensorMap tmX, const float* __restrict__ w, const float* __restrict__ bias,
                    __nv_bfloat16* __restrict__ y, const StemTc d, const StemTma g) {
    constexpr int NPR = KT * 3;
    constexpr int NCH = (NPR * 5 + 3) / 4;                 // 12 chunks (96 k) or 4 (32 k)
    constexpr int KSTEPS = NCH / 2;
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t mma_bar[2], full[STM_MAX_STAGES], empty[STM_MAX_STAGES];
    __shared__ uint32_t tmem_base_s;
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t a_tiles = (raw + 1023u) & ~1023u;       // 2 x 32 KB im2col tiles
    const uint32_t w_tile = a_tiles + 2 * STC_TILE_BYTES;  // fp16 weights, 4 KB
    const uint32_t ring = w_tile + 4096;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        mbar_init(&mma_bar[0], 1); mbar_init(&mma_bar[1], 1);
        for (int i = 0; i < g.nst; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 8); }
        fence_barrier_init();
        tma_prefetch_desc(&tmX);
    }
    if (warp == 1) tmem_alloc(&tmem_base_s, 64);
    // weights -> fp16, K order k' = pr*10 + kw*3 + ci
    for (int i = tid; i < 16 * STC_KPAD; i += STM_THREADS) {
        const int co = i / STC_KPAD, k = i % STC_KPAD;
        const int pr = k / 10, j = k % 10;
        float v = 0.f;
        if (pr < NPR && j < 9) v = w[((long long)co * 3 + (j % 3)) * (KT * 9) + pr * 3 + j / 3];
        const uint32_t off = (uint32_t)(co >> 3) * STC_GROUP_BYTES + (uint32_t)(k >> 3) * 128 + (uint32_t)(co & 7) * 16 + (uint32_t)(k & 7) * 2;
        asm volatile("st.shared.u16 [%0], %1;" ::"r"(w_tile + off), "h"(__half_as_ushort(__float2half_rn(v))) : "memory");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 8) {                                       // producer
        if (lane == 0) {
            const uint32_t nsteps = (uint32_t)g.steps;
            int sg = 0;
            uint32_t ph = 1;                               // parity of "slot was never used": passes at once
            for (uint32_t s = blockIdx.x; s < nsteps; s += gridDim.x) {
                mbar_wait_parked(&empty[sg], ph);
                const StmStep c = stm_decode(s, d, g);
                mbar_expect_tx(&full[sg], g.xbytes);
                tma_load_4d(reinterpret_cast<void*>(__cvta_shared_to_generic(ring + (uint32_t)sg * g.stage_pitch)), &tmX,
                            &full[sg], g.e0, c.hg * g.RT * d.sH - d.pH, c.to * d.sT - d.pT, c.b);
                if (++sg == g.nst) { sg = 0; ph ^= 1u; }
            }
        }
    } else {
        const uint32_t idesc = make_idesc_fmt(128, 16, 0, 0, 0, 0);      // fp16 x fp16
        float bv[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) bv[j] = bias ? __ldg(bias + j) : 0.f;
        const float inv = d.inv_scale;
        const int tile_sel = tid >> 7, row = tid & 127;
        const int r = tid / g.TW, wo = tid % g.TW;
        // epilogue: this thread reads TMEM lane (warp&3)*32+lane of tile warp>>2 = pixel q of the step
        const int q = (warp >> 2) * 128 + (warp & 3) * 32 + lane;
        const int qr = q / g.TW, qw = q % g.TW;
        const uint32_t nsteps = (uint32_t)g.steps;
        uint32_t it = 0, ph = 0;
        int sg = 0;
        for (uint32_t s = blockIdx.x; ; s += gridDim.x, ++it) {
            const bool have = s < nsteps;
            const int st = (int)(it & 1);
            if (it >= 1) {                                 // previous MMAs done: tiles free, accumulators ready
                mbar_wait_parked(&mma_bar[st ^ 1], ((it - 1) >> 1) & 1u);
                tc_fence_after();
            }
            if (have) {
                mbar_wait_parked(&full[sg], ph);
                if (r < g.RT)
                    gather_row_smem<KT, NCH, false, false>(ring + (uint32_t)sg * g.stage_pitch, g, d.sH, d.sW, r, wo, true,
                                                    a_tiles + (uint32_t)tile_sel * STC_TILE_BYTES, row);
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[sg]);
                if (++sg == g.nst) { sg = 0; ph ^= 1u; }
            }
            named_bar_sync(1, 256);
            if (have && tid == 0) {
                tc_fence_after();
#pragma unroll
                for (int t2 = 0; t2 < 2; ++t2) {
                    const uint32_t a0 = a_tiles + (uint32_t)t2 * STC_TILE_BYTES;
#pragma unroll
                    for (int ks = 0; ks < KSTEPS; ++ks)
                        umma_bf16(tmem_base + (uint32_t)(st * 32 + t2 * 16), desc_ns(a0 + ks * 256, 128, STC_GROUP_BYTES),
                                  desc_ns(w_tile + ks * 256, 128, STC_GROUP_BYTES), idesc, ks != 0);
                }
                umma_commit(&mma_bar[st]);
            }
            if (it >= 1) {                                 // epilogue of the previous step
                const StmStep c = stm_decode(s - gridDim.x, d, g);
                const int pst = st ^ 1;
                uint32_t rg[16];
                tmem_ld16(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(pst * 32 + (warp >> 2) * 16), rg);
                tmem_ld_wait();
                const int ho = c.hg * g.RT + qr;
                if (qr < g.RT && ho < d.Ho) {
                    const long long p = (((long long)c.b * d.To + c.to) * d.Ho + ho) * d.Wo + qw;
                    float v[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = fmaf(__uint_as_float(rg[j]), inv, bv[j]);
                    act_fwd_vec(v, d.act, d.slope);
                    uint4 o0, o1;
                    o0.x = pack_bf16x2(v[0], v[1]);   o0.y = pack_bf16x2(v[2], v[3]);
                    o0.z = pack_bf16x2(v[4], v[5]);   o0.w = pack_bf16x2(v[6], v[7]);
                    o1.x = pack_bf16x2(v[8], v[9]);   o1.y = pack_bf16x2(v[10], v[11]);
                    o1.z = pack_bf16x2(v[12], v[13]); o1.w = pack_bf16x2(v[14], v[15]);
                    uint4* dst = reinterpret_cast<uint4*>(y + p * 16);
                    dst[0] = o0; dst[1] = o1;
                }
                tc_fence_before();
            }
            if (!have) break;
        }
    }
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 64); }
}

// weight (and bias) gradient: dW[k][co] = sum_p Xcol[p][k] * dY[p][co]; A = the im2col tiles (bf16 integers here) read
// MN-major, B = dY (bf16) staged by a bulk copy next to the clip box and re-laid into core matrices by the gather threads.
template <int KT>
__global__ void __launch_bounds__(STM_THREADS, 2)
stem_tma_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __nv_bfloat16* __restrict__ dy,
                      float* __restrict__ dw, float* __restrict__ dbias, const StemTc d, const StemTma g) {
    constexpr int NPR = KT * 3;
    constexpr int NK = NPR * 10;                           // ones column
    constexpr int NCH = (NPR * 5 + 1 + 3) / 4;             // chunks covering the values and the ones column (12 / 4)
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t mma_bar, full[STM_MAX_STAGES], empty[STM_MAX_STAGES];
    __shared__ uint32_t tmem_base_s;
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t a_tiles = (raw + 1023u) & ~1023u;
    const uint32_t g_tiles = a_tiles + 2 * STC_TILE_BYTES; // 2 x 4 KB dy tiles: offset(p,co) = (p/8)*256 + (co/8)*128 + (p%8)*16
    const uint32_t ring = g_tiles + 2 * 4096;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        mbar_init(&mma_bar, 1);
        for (int i = 0; i < g.nst; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 8); }
        fence_barrier_init();
        tma_prefetch_desc(&tmX);
    }
    if (warp == 1) tmem_alloc(&tmem_base_s, 32);
    // rows no thread ever writes (q >= TW*RT) must hold finite values: they meet dy = 0
    for (uint32_t i = (uint32_t)tid * 16; i < 2 * STC_TILE_BYTES; i += STM_THREADS * 16) sts128(a_tiles + i, 0u, 0u, 0u, 0u);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    const uint32_t nsteps = (uint32_t)g.steps;
    uint32_t it = 0;
    int sg = 0;
    if (warp == 8) {
        if (lane == 0) {
            uint32_t ph = 1;
            for (uint32_t s = blockIdx.x; s < nsteps; s += gridDim.x) {
                mbar_wait_parked(&empty[sg], ph);
                const StmStep c = stm_decode(s, d, g);
                const int rows = min(g.RT, d.Ho - c.hg * g.RT);
                const uint32_t dyb = (uint32_t)(rows * g.TW) * 32u;
                const long long p0 = (((long long)c.b * d.To + c.to) * d.Ho + (long long)c.hg * g.RT) * d.Wo;
                const uint32_t dst = ring + (uint32_t)sg * g.stage_pitch;
                mbar_expect_tx(&full[sg], g.xbytes + dyb);
                tma_load_4d(reinterpret_cast<void*>(__cvta_shared_to_generic(dst)), &tmX, &full[sg], g.e0,
                            c.hg * g.RT * d.sH - d.pH, c.to * d.sT - d.pT, c.b);
                bulk_load_1d(dst + ((g.xbytes + 127u) & ~127u), dy + p0 * 16, dyb, &full[sg]);
                if (++sg == g.nst) { sg = 0; ph ^= 1u; }
            }
        }
    } else {
        const uint32_t idesc = make_idesc(128, 16, 1, 1);               // bf16 x bf16, both MN-major
        const int tile_sel = tid >> 7, row = tid & 127;
        const int r = tid / g.TW, wo = tid % g.TW;
        uint32_t ph = 0;
        for (uint32_t s = blockIdx.x; s < nsteps; s += gridDim.x, ++it) {
            if (it >= 1) {
                mbar_wait_parked(&mma_bar, (it - 1) & 1u);
                tc_fence_after();
            }
            mbar_wait_parked(&full[sg], ph);
            const int hg = (int)(s % (uint32_t)g.groups);
            const bool valid = r < g.RT && hg * g.RT + r < d.Ho;
            const uint32_t stage = ring + (uint32_t)sg * g.stage_pitch;
            if (r < g.RT)
                gather_row_smem<KT, NCH, true, true>(stage, g, d.sH, d.sW, r, wo, valid, a_tiles + (uint32_t)tile_sel * STC_TILE_BYTES, row);
            {
                uint4 g0 = make_uint4(0u, 0u, 0u, 0u), g1 = g0;
                if (valid) {
                    const uint32_t src = stage + ((g.xbytes + 127u) & ~127u) + (uint32_t)tid * 32u;
                    g0 = lds128(src); g1 = lds128(src + 16);
                }
                const uint32_t gb = g_tiles + (uint32_t)tile_sel * 4096 + (uint32_t)(row >> 3) * 256 + (uint32_t)(row & 7) * 16;
                sts128(gb, g0.x, g0.y, g0.z, g0.w);
                sts128(gb + 128, g1.x, g1.y, g1.z, g1.w);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[sg]);
            if (++sg == g.nst) { sg = 0; ph ^= 1u; }
            named_bar_sync(1, 256);
            if (tid == 0) {
                tc_fence_after();
#pragma unroll
                for (int t2 = 0; t2 < 2; ++t2) {
                    const uint32_t a0 = a_tiles + (uint32_t)t2 * STC_TILE_BYTES;
                    const uint32_t b0 = g_tiles + (uint32_t)t2 * 4096;
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks)
                        umma_bf16(tmem_base, desc_ns(a0 + ks * 2 * STC_GROUP_BYTES, STC_GROUP_BYTES, 128),
                                  desc_ns(b0 + ks * 512, 256, 128), idesc, (it | t2 | ks) != 0);
                }
                umma_commit(&mma_bar);
            }
        }
        if (it >= 1) {                                     // drain: D[k][co], k = TMEM lane
            mbar_wait_parked(&mma_bar, (it - 1) & 1u);
            tc_fence_after();
            if (warp < 4) {
                uint32_t rg[16];
                tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16), rg);
                tmem_ld_wait();
                const int k = warp * 32 + lane;
                const int pr = k / 10, j = k % 10;
                if (k < NK && j < 9) {
                    const int ci = j % 3, tap = pr * 3 + j / 3;
#pragma unroll
                    for (int co = 0; co < 16; ++co)
                        atomicAdd(&dw[((long long)co * 3 + ci) * (KT * 9) + tap], __uint_as_float(rg[co]) * d.inv_scale);
                } else if (k == NK && dbias) {
#pragma unroll
                    for (int co = 0; co < 16; ++co) atomicAdd(&dbias[co], __uint_as_float(rg[co]));
                }
            }
            tc_fence_before();
        }
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

// Geometry + tensor map of the TMA path; false = not covered (strided or unaligned clip, wide frames): the gather
// kernels take over.
static bool stm_plan(const void* x, int x_dtype, const void* yptr, int kT, const StemTc& d, bool wgrad, StemTma* g,
                     CUtensorMap* tm, size_t* smem) {
    if (x_dtype != PB_U8 || (kT != 3 && kT != 1)) return false;
    const long long rowbytes = (long long)d.W * 3;
    if (d.xs_h != rowbytes || d.xs_t != rowbytes * d.H || d.xs_b != rowbytes * d.H * d.T) return false;
    if (rowbytes % 16 != 0 || (reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(yptr) & 15)) return false;
    if (d.Wo > 128 || d.Wo < 1) return false;
    g->TW = d.Wo;
    const int bs = -3 * d.pW;
    g->e0 = (bs >= 0 ? bs / 16 : -((-bs + 15) / 16)) * 4;     // the box must start on a 16-byte boundary
    g->boff = bs - 4 * g->e0;
    const int words = (g->boff + (g->TW - 1) * d.sW * 3 + 12 + 3) / 4;
    const int IW = (words + 3) / 4 * 4;
    if (IW > 256) return false;
    g->rowb = IW * 4;
    const size_t fixed = 1024 + 2 * STC_TILE_BYTES + (wgrad ? 2 * 4096 : 4096);
    const size_t budget = 110 * 1024;                      // two CTAs per SM
    for (g->RT = std::min(256 / g->TW, d.Ho);; --g->RT) {  // as many output rows per step as two ring stages allow
        if (g->RT < 1) return false;
        g->NR = (g->RT - 1) * d.sH + 3;
        g->xbytes = (uint32_t)g->rowb * g->NR * kT;
        g->stage_pitch = (g->xbytes + 127u) & ~127u;
        if (wgrad) g->stage_pitch += ((uint32_t)(g->TW * g->RT) * 32u + 127u) & ~127u;
        if (g->NR <= 256 && fixed + 2 * (size_t)g->stage_pitch <= budget) break;
    }
    g->nst = (int)std::min<size_t>(STM_MAX_STAGES, (budget - fixed) / g->stage_pitch);
    g->groups = (d.Ho + g->RT - 1) / g->RT;
    g->steps = (long long)d.B * d.To * g->groups;
    if (g->steps >= (1LL << 31)) return false;
    *smem = fixed + (size_t)g->nst * g->stage_pitch;
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t gdim[4] = {(cuuint64_t)(rowbytes / 4), (cuuint64_t)d.H, (cuuint64_t)d.T, (cuuint64_t)d.B};
    cuuint64_t gstr[3] = {(cuuint64_t)rowbytes, (cuuint64_t)(rowbytes * d.H), (cuuint64_t)(rowbytes * d.H * d.T)};
    cuuint32_t box[4] = {(cuuint32_t)IW, (cuuint32_t)g->NR, (cuuint32_t)kT, 1u};
    cuuint32_t es[4] = {1, 1, 1, 1};
    return fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, const_cast<void*>(x), gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int stm_grid(const StemTma& g) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return (int)std::min<long long>(g.steps, 2LL * sms);
}

// Both return true if they launched; false = shape not covered (caller uses the direct kernels).
bool stem_tc_fwd(const void* x, int x_dtype, const float* w, const float* bias, void* y, int kT, const StemTc& d,
                 cudaStream_t st) {
    {
        StemTma g; CUtensorMap tm; size_t sm = 0;
        static unsigned long long attr3 = 0, attr1 = 0;
        if (!getenv("PB_STEM_GATHER") && stm_plan(x, x_dtype, y, kT, d, false, &g, &tm, &sm)) {
            if (kT == 3) {
                if (ensure_dyn_smem(stem_tma_fwd_kernel<3>, 112 * 1024, &attr3) != cudaSuccess) return false;
                stem_tma_fwd_kernel<3><<<stm_grid(g), STM_THREADS, sm, st>>>(tm, w, bias, (__nv_bfloat16*)y, d, g);
            } else {
                if (ensure_dyn_smem(stem_tma_fwd_kernel<1>, 112 * 1024, &attr1) != cudaSuccess) return false;
                stem_tma_fwd_kernel<1><<<stm_grid(g), STM_THREADS, sm, st>>>(tm, w, bias, (__nv_bfloat16*)y, d, g);
            }
            count_path(PB_PATH_STEM_TMA);
            return true;
        }
    }
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
    {
        StemTma g; CUtensorMap tm; size_t sm = 0;
        static unsigned long long attr3 = 0, attr1 = 0;
        if (!getenv("PB_STEM_GATHER") && stm_plan(x, x_dtype, dy, kT, d, true, &g, &tm, &sm)) {
            if (kT == 3) {
                if (ensure_dyn_smem(stem_tma_wgrad_kernel<3>, 112 * 1024, &attr3) != cudaSuccess) return false;
                stem_tma_wgrad_kernel<3><<<stm_grid(g), STM_THREADS, sm, st>>>(tm, (const __nv_bfloat16*)dy, dw, dbias, d, g);
            } else {
                if (ensure_dyn_smem(stem_tma_wgrad_kernel<1>, 112 * 1024, &attr1) != cudaSuccess) return false;
                stem_tma_wgrad_kernel<1><<<stm_grid(g), STM_THREADS, sm, st>>>(tm, (const __nv_bfloat16*)dy, dw, dbias, d, g);
            }
            count_path(PB_PATH_STEM_TMA);
            return true;
        }
    }
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
