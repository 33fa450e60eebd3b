// Depthwise (1,k,k) convolution fast paths for bf16 NDHWC activations: TMA-staged halo tiles.
//
// A persistent CTA owns one channel block (<= 128 channels) and walks over (sample, frame, h-tile, w-tile)
// work items.  For each item cp.async.bulk.tensor.5d fetches the halo tile [Hi][Wi][Cb] of the input frame
// (and, for the weight gradient, the matching dy tile) into a shared-memory ring; out-of-bounds
// coordinates are zero-filled by the TMA unit, so spatial padding costs nothing and needs no bounds
// checks, and the frames created by the reference's scalar temporal padding (mobilenet.py:67-75) are
// written as zeros without reading anything.  Each thread owns 4 channels (8-byte vectors) and a strip of
// output pixels: every staged vector is unpacked once and fed to all the outputs of the strip that use it
// with packed fp32 FMAs (fma.rn.f32x2); filter taps stay in registers.
//
//   forward            : y  = conv(x, w)                     (also the stride-1 input gradient: flipped w)
//   input gradient, s=2: dx = gather of dy over the taps whose parity matches
//   weight gradient    : dw accumulated in registers over all tiles of the CTA, reduced once at the end
//                        (warp shuffles, shared memory, then one fp32 atomic per tap and channel per CTA)
#include <algorithm>
#include <mutex>

#include "dwconv.cuh"
#include "tc_common.cuh"

namespace pb {

using namespace tc;

constexpr int DWT_MAX_STAGES = 4;
constexpr int DWT_THREADS = 256;
constexpr int DWT_MAX_ZF = 64;

struct DwTile {
    int B, C;
    int Ti, Hin, Win;          // tensor that is staged through TMA ("source")
    int To, Ho, Wo;            // tensor the tiles are defined on ("destination")
    int pS;                    // spatial padding of the convolution
    int Cb, Gb, nblk;          // channel block, 4-channel groups per block, number of blocks
    int Ht, Wt, Hi, Wi, tiles_h, tiles_w;
    int f_first, f_step, f_count;      // destination frames with a source frame: f = f_first + n*f_step
    int src_first, src_step;           // ... and the source frame of the n-th one
    int nzf;                           // destination frames without a source frame (all zero)
    unsigned char zf[DWT_MAX_ZF];
    long long ntiles;
    int stage_bytes, box_bytes, box2_bytes, stages;
    int flip;
};

__device__ __forceinline__ void ffma2(float2& d, const float2 a, const float2 b) {
    unsigned long long dd = *reinterpret_cast<unsigned long long*>(&d);
    asm("fma.rn.f32x2 %0, %1, %2, %0;"
        : "+l"(dd)
        : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)));
    d = *reinterpret_cast<float2*>(&dd);
}
__device__ __forceinline__ float2 unpack2(uint32_t u) { return make_float2(bf16_lo(u), bf16_hi(u)); }

struct TileCoord { int b, n, h0, w0; };

__device__ __forceinline__ TileCoord decode_tile(const DwTile& p, long long t) {
    TileCoord c;
    int tw = (int)(t % p.tiles_w); t /= p.tiles_w;
    int th = (int)(t % p.tiles_h); t /= p.tiles_h;
    c.n = (int)(t % p.f_count);
    c.b = (int)(t / p.f_count);
    c.h0 = th * p.Ht; c.w0 = tw * p.Wt;
    return c;
}

// zero the destination frames that have no source frame (only CTAs with blockIdx.y == 0 call this)
__device__ __forceinline__ void zero_frames(const DwTile& p, __nv_bfloat16* y) {
    if (p.nzf <= 0) return;
    const long long frame16 = (long long)p.Ho * p.Wo * p.C / 8;
    const long long total = (long long)p.B * p.nzf * frame16;
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    for (long long i = (long long)blockIdx.x * DWT_THREADS + threadIdx.x; i < total;
         i += (long long)gridDim.x * DWT_THREADS) {
        long long e = i % frame16;
        long long q = i / frame16;
        int f = p.zf[(int)(q % p.nzf)];
        int b = (int)(q / p.nzf);
        reinterpret_cast<uint4*>(y)[((long long)b * p.To + f) * frame16 + e] = z;
    }
}

// ------------------------------------------------------------------------------------------------
// forward (and stride-1 dgrad): destination tile [Ht][Wt], source halo [(Ht-1)S+K][(Wt-1)S+K]
// ------------------------------------------------------------------------------------------------
template <int K, int S, int WS>
__global__ void __launch_bounds__(DWT_THREADS, 2)
dw_fwd_tma_kernel(const __grid_constant__ CUtensorMap tmX, const float* __restrict__ w_tc,
                  __nv_bfloat16* __restrict__ y, const DwTile p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_bar[DWT_MAX_STAGES];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* ring = smem_raw + (((raw + 127u) & ~127u) - raw);
    const int tid = threadIdx.x;
    const int c_base = blockIdx.y * p.Cb;
    const int cb_bytes = p.Cb * 2;

    if (tid == 0) {
        tma_prefetch_desc(&tmX);
        for (int s = 0; s < p.stages; ++s) mbar_init(&full_bar[s], 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (blockIdx.y == 0) zero_frames(p, y);

    const int ppp = DWT_THREADS / p.Gb;
    const int g = tid % p.Gb;
    const bool active = tid < ppp * p.Gb && (c_base + g * 4) < p.C;
    const int slot = tid / p.Gb;
    // K == 3: fp32 pairs (36 registers); K == 5: packed bf16 pairs (50 registers; lossless, the tap-major
    // weights were rounded to bf16 for this dtype already)
    constexpr bool PACKED = (K == 5);
    float2 wv[PACKED ? 1 : K * K][2];
    uint32_t wp[PACKED ? K * K : 1][2];
#pragma unroll
    for (int t = 0; t < K * K; ++t) {
        const int src = p.flip ? (K * K - 1 - t) : t;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            float lo = 0.f, hi = 0.f;
            if (active) {
                const float* wsrc = w_tc + (long long)src * p.C + c_base + g * 4 + c * 2;
                lo = wsrc[0]; hi = wsrc[1];
            }
            if constexpr (PACKED) wp[t][c] = pack_bf16x2(lo, hi);
            else wv[t][c] = make_float2(lo, hi);
        }
    }

    const long long my_tiles = p.ntiles > blockIdx.x ? (p.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int nstrips = p.Wt / WS;
    const int items = p.Ht * nstrips;

    auto issue = [&](long long n) {
        TileCoord tc_ = decode_tile(p, blockIdx.x + n * (long long)gridDim.x);
        const int s = (int)(n % p.stages);
        mbar_expect_tx(&full_bar[s], (uint32_t)p.box_bytes);
        tma_load_5d(ring + (size_t)s * p.stage_bytes, &tmX, &full_bar[s], c_base, tc_.w0 * S - p.pS, tc_.h0 * S - p.pS,
                    p.src_first + tc_.n * p.src_step, tc_.b);
    };
    if (tid == 0)
        for (long long n = 0; n < my_tiles && n < p.stages; ++n) issue(n);

    for (long long n = 0; n < my_tiles; ++n) {
        const int s = (int)(n % p.stages);
        mbar_wait(&full_bar[s], (uint32_t)((n / p.stages) & 1));
        const TileCoord tc_ = decode_tile(p, blockIdx.x + n * (long long)gridDim.x);
        const int to = p.f_first + tc_.n * p.f_step;
        const uint8_t* tile = ring + (size_t)s * p.stage_bytes;
        if (active) {
            for (int it = slot; it < items; it += ppp) {
                const int strip = it % nstrips, hl = it / nstrips;
                const int ho = tc_.h0 + hl;
                if (ho >= p.Ho) continue;
                float2 acc[WS][2];
#pragma unroll
                for (int o = 0; o < WS; ++o) { acc[o][0] = make_float2(0.f, 0.f); acc[o][1] = make_float2(0.f, 0.f); }
#pragma unroll
                for (int i = 0; i < K; ++i) {
                    const uint8_t* rowp = tile + ((size_t)(hl * S + i) * p.Wi + (size_t)strip * WS * S) * cb_bytes + g * 8;
                    if constexpr (!PACKED) {
#pragma unroll
                        for (int j = 0; j < (WS - 1) * S + K; ++j) {
                            const uint2 u = *reinterpret_cast<const uint2*>(rowp + (size_t)j * cb_bytes);
                            const float2 v0 = unpack2(u.x), v1 = unpack2(u.y);
#pragma unroll
                            for (int o = 0; o < WS; ++o) {
                                const int t = j - o * S;
                                if (t >= 0 && t < K) {
                                    ffma2(acc[o][0], v0, wv[i * K + t][0]);
                                    ffma2(acc[o][1], v1, wv[i * K + t][1]);
                                }
                            }
                        }
                    } else {
                        constexpr int NJ = (WS - 1) * S + K;
                        float2 win[NJ][2];
#pragma unroll
                        for (int j = 0; j < NJ; ++j) {
                            const uint2 u = *reinterpret_cast<const uint2*>(rowp + (size_t)j * cb_bytes);
                            win[j][0] = unpack2(u.x); win[j][1] = unpack2(u.y);
                        }
#pragma unroll
                        for (int t = 0; t < K; ++t) {
                            const float2 w0 = unpack2(wp[i * K + t][0]), w1 = unpack2(wp[i * K + t][1]);
#pragma unroll
                            for (int o = 0; o < WS; ++o) {
                                ffma2(acc[o][0], win[o * S + t][0], w0);
                                ffma2(acc[o][1], win[o * S + t][1], w1);
                            }
                        }
                    }
                }
                const int wo0 = tc_.w0 + strip * WS;
                __nv_bfloat16* yp = y + ((((long long)tc_.b * p.To + to) * p.Ho + ho) * p.Wo + wo0) * p.C + c_base + g * 4;
#pragma unroll
                for (int o = 0; o < WS; ++o) {
                    if (wo0 + o < p.Wo) {
                        uint2 out;
                        out.x = pack_bf16x2(acc[o][0].x, acc[o][0].y);
                        out.y = pack_bf16x2(acc[o][1].x, acc[o][1].y);
                        *reinterpret_cast<uint2*>(yp + (long long)o * p.C) = out;
                    }
                }
            }
        }
        __syncthreads();                                  // everyone is done reading stage s
        if (tid == 0 && n + p.stages < my_tiles) issue(n + p.stages);
    }
}

// ------------------------------------------------------------------------------------------------
// stride-2 input gradient: destination tile of dx [Ht][Wt] (Ht, Wt even, tile origin even), source halo of
// dy rows ho_base .. ho_base + Ht/2 + (K-1)/2, same for columns.
//   dx[h][w] = sum_{i,j : (h+p-i), (w+p-j) even} dy[(h+p-i)/2][(w+p-j)/2] * w[i][j]
// ------------------------------------------------------------------------------------------------
template <int K, int XS>
__global__ void __launch_bounds__(DWT_THREADS, 2)
dw_dgrad_s2_tma_kernel(const __grid_constant__ CUtensorMap tmD, const float* __restrict__ w_tc,
                       __nv_bfloat16* __restrict__ dx, const DwTile p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_bar[DWT_MAX_STAGES];
    constexpr int P = K / 2;
    constexpr int HALF = (K - 1) / 2;         // extra dy rows/cols on the low side: p/2 rounded down is P/2
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* ring = smem_raw + (((raw + 127u) & ~127u) - raw);
    const int tid = threadIdx.x;
    const int c_base = blockIdx.y * p.Cb;
    const int cb_bytes = p.Cb * 2;

    if (tid == 0) {
        tma_prefetch_desc(&tmD);
        for (int s = 0; s < p.stages; ++s) mbar_init(&full_bar[s], 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (blockIdx.y == 0) zero_frames(p, dx);

    const int ppp = DWT_THREADS / p.Gb;
    const int g = tid % p.Gb;
    const bool active = tid < ppp * p.Gb && (c_base + g * 4) < p.C;
    const int slot = tid / p.Gb;
    uint32_t wp[K * K][2];                     // packed bf16 pairs
#pragma unroll
    for (int t = 0; t < K * K; ++t)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            float lo = 0.f, hi = 0.f;
            if (active) {
                const float* wsrc = w_tc + (long long)t * p.C + c_base + g * 4 + c * 2;
                lo = wsrc[0]; hi = wsrc[1];
            }
            wp[t][c] = pack_bf16x2(lo, hi);
        }

    const long long my_tiles = p.ntiles > blockIdx.x ? (p.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int nstrips = p.Wt / XS;
    const int items = p.Ht * nstrips;
    (void)HALF;

    auto issue = [&](long long n) {
        TileCoord tc_ = decode_tile(p, blockIdx.x + n * (long long)gridDim.x);
        const int s = (int)(n % p.stages);
        mbar_expect_tx(&full_bar[s], (uint32_t)p.box_bytes);
        tma_load_5d(ring + (size_t)s * p.stage_bytes, &tmD, &full_bar[s], c_base, tc_.w0 / 2 - P / 2, tc_.h0 / 2 - P / 2,
                    p.src_first + tc_.n * p.src_step, tc_.b);
    };
    if (tid == 0)
        for (long long n = 0; n < my_tiles && n < p.stages; ++n) issue(n);

    for (long long n = 0; n < my_tiles; ++n) {
        const int s = (int)(n % p.stages);
        mbar_wait(&full_bar[s], (uint32_t)((n / p.stages) & 1));
        const TileCoord tc_ = decode_tile(p, blockIdx.x + n * (long long)gridDim.x);
        const int t_dst = p.f_first + tc_.n * p.f_step;
        const uint8_t* tile = ring + (size_t)s * p.stage_bytes;
        if (active) {
            for (int it = slot; it < items; it += ppp) {
                const int strip = it % nstrips, hl = it / nstrips;
                const int h = tc_.h0 + hl;
                if (h >= p.Ho) continue;
                float2 acc[XS][2];
#pragma unroll
                for (int o = 0; o < XS; ++o) { acc[o][0] = make_float2(0.f, 0.f); acc[o][1] = make_float2(0.f, 0.f); }
                // local dy column of tap j for dx column xl (tile origin even): (xl + P - j)/2 + P/2
                // the strip starts at an even xl0, so parities are compile-time per o
#pragma unroll
                for (int i = 0; i < K; ++i) {
                    if (((hl + P - i) & 1) != 0) continue;                 // tile origin even: parity of h == hl
                    const int rl = ((hl + P - i) >> 1) + P / 2;            // local dy row
                    const uint8_t* rowp = tile + ((size_t)rl * p.Wi + (size_t)(strip * XS / 2)) * cb_bytes + g * 8;
                    constexpr int NC = XS / 2 + (K - 1) / 2 + 1;            // dy columns a strip can touch
                    float2 win[NC][2];
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        const uint2 u = *reinterpret_cast<const uint2*>(rowp + (size_t)c * cb_bytes);
                        win[c][0] = unpack2(u.x); win[c][1] = unpack2(u.y);
                    }
#pragma unroll
                    for (int o = 0; o < XS; ++o) {
#pragma unroll
                        for (int j = 0; j < K; ++j) {
                            if (((o + P - j) & 1) == 0) {
                                const int cl = ((o + P - j) >> 1) + P / 2;   // column relative to strip*XS/2
                                const float2 w0 = unpack2(wp[i * K + j][0]), w1 = unpack2(wp[i * K + j][1]);
                                ffma2(acc[o][0], win[cl][0], w0);
                                ffma2(acc[o][1], win[cl][1], w1);
                            }
                        }
                    }
                }
                const int x0 = tc_.w0 + strip * XS;
                __nv_bfloat16* xp = dx + ((((long long)tc_.b * p.To + t_dst) * p.Ho + h) * p.Wo + x0) * p.C + c_base + g * 4;
#pragma unroll
                for (int o = 0; o < XS; ++o) {
                    if (x0 + o < p.Wo) {
                        uint2 out;
                        out.x = pack_bf16x2(acc[o][0].x, acc[o][0].y);
                        out.y = pack_bf16x2(acc[o][1].x, acc[o][1].y);
                        *reinterpret_cast<uint2*>(xp + (long long)o * p.C) = out;
                    }
                }
            }
        }
        __syncthreads();
        if (tid == 0 && n + p.stages < my_tiles) issue(n + p.stages);
    }
}

// ------------------------------------------------------------------------------------------------
// weight gradient: tiles over dy [Ht][Wt]; both the x halo tile and the dy tile are staged.  A thread owns
// (4 channels, one filter row i) and keeps K x 4 fp32 sums in registers across ALL tiles of the CTA.
// ------------------------------------------------------------------------------------------------
template <int K, int S, int WS>
__global__ void __launch_bounds__(DWT_THREADS, 2)
dw_wgrad_tma_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmD,
                    float* __restrict__ dw_tc, const DwTile p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_bar[DWT_MAX_STAGES];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* ring = smem_raw + (((raw + 127u) & ~127u) - raw);
    const int tid = threadIdx.x;
    const int c_base = blockIdx.y * p.Cb;
    const int cb_bytes = p.Cb * 2;
    const int x_bytes = (p.box_bytes + 127) / 128 * 128;       // dy tile follows the x tile in a stage

    if (tid == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmD);
        for (int s = 0; s < p.stages; ++s) mbar_init(&full_bar[s], 1);
        fence_barrier_init();
    }
    __syncthreads();

    const int lanes = p.Gb * K;                         // (group, filter row) combinations
    const int ppp = DWT_THREADS / lanes;
    const int g = tid % p.Gb;
    const int i = (tid / p.Gb) % K;
    const bool active = tid < ppp * lanes && (c_base + g * 4) < p.C;
    const int slot = tid / lanes;
    float2 acc[K][2];
#pragma unroll
    for (int t = 0; t < K; ++t) { acc[t][0] = make_float2(0.f, 0.f); acc[t][1] = make_float2(0.f, 0.f); }

    const long long my_tiles = p.ntiles > blockIdx.x ? (p.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int nstrips = p.Wt / WS;
    const int items = p.Ht * nstrips;

    auto issue = [&](long long n) {
        TileCoord tc_ = decode_tile(p, blockIdx.x + n * (long long)gridDim.x);
        const int s = (int)(n % p.stages);
        mbar_expect_tx(&full_bar[s], (uint32_t)(p.box_bytes + p.box2_bytes));
        uint8_t* st = ring + (size_t)s * p.stage_bytes;
        tma_load_5d(st, &tmX, &full_bar[s], c_base, tc_.w0 * S - p.pS, tc_.h0 * S - p.pS,
                    p.src_first + tc_.n * p.src_step, tc_.b);
        tma_load_5d(st + x_bytes, &tmD, &full_bar[s], c_base, tc_.w0, tc_.h0, p.f_first + tc_.n * p.f_step, tc_.b);
    };
    if (tid == 0)
        for (long long n = 0; n < my_tiles && n < p.stages; ++n) issue(n);

    for (long long n = 0; n < my_tiles; ++n) {
        const int s = (int)(n % p.stages);
        mbar_wait(&full_bar[s], (uint32_t)((n / p.stages) & 1));
        const uint8_t* xt = ring + (size_t)s * p.stage_bytes;
        const uint8_t* dt = xt + x_bytes;
        if (active) {
            for (int it = slot; it < items; it += ppp) {
                const int strip = it % nstrips, hl = it / nstrips;
                const uint8_t* dp = dt + ((size_t)hl * p.Wt + (size_t)strip * WS) * cb_bytes + g * 8;
                const uint8_t* xp = xt + ((size_t)(hl * S + i) * p.Wi + (size_t)strip * WS * S) * cb_bytes + g * 8;
                float2 dyv[WS][2];
#pragma unroll
                for (int o = 0; o < WS; ++o) {
                    const uint2 u = *reinterpret_cast<const uint2*>(dp + (size_t)o * cb_bytes);
                    dyv[o][0] = unpack2(u.x); dyv[o][1] = unpack2(u.y);
                }
#pragma unroll
                for (int j = 0; j < (WS - 1) * S + K; ++j) {
                    const uint2 u = *reinterpret_cast<const uint2*>(xp + (size_t)j * cb_bytes);
                    const float2 v0 = unpack2(u.x), v1 = unpack2(u.y);
#pragma unroll
                    for (int o = 0; o < WS; ++o) {
                        const int t = j - o * S;
                        if (t >= 0 && t < K) {
                            ffma2(acc[t][0], v0, dyv[o][0]);
                            ffma2(acc[t][1], v1, dyv[o][1]);
                        }
                    }
                }
            }
        }
        __syncthreads();
        if (tid == 0 && n + p.stages < my_tiles) issue(n + p.stages);
    }

    // reduce over the CTA's threads that share (g, i): shared-memory sums, then one atomic per (tap, channel)
    float* red = reinterpret_cast<float*>(ring);            // [K*K][Cb]
    for (int e = tid; e < K * K * p.Cb; e += DWT_THREADS) red[e] = 0.f;
    __syncthreads();
    if (active) {
#pragma unroll
        for (int t = 0; t < K; ++t) {
            float* r = red + (size_t)(i * K + t) * p.Cb + g * 4;
            atomicAdd(r + 0, acc[t][0].x); atomicAdd(r + 1, acc[t][0].y);
            atomicAdd(r + 2, acc[t][1].x); atomicAdd(r + 3, acc[t][1].y);
        }
    }
    __syncthreads();
    for (int e = tid; e < K * K * p.Cb; e += DWT_THREADS) {
        const int tap = e / p.Cb, c = e % p.Cb;
        if (c_base + c < p.C) atomicAdd(&dw_tc[(long long)tap * p.C + c_base + c], red[e]);
    }
}

// ------------------------------------------------------------------------------------------------
// host-side planning
// ------------------------------------------------------------------------------------------------
struct PlanIn {
    int B, C;
    int Ti, Hin, Win;      // staged tensor
    int To, Ho, Wo;        // tiled tensor
    int K, S, WS;          // S: source step per destination pixel (1 or 2); for dgrad-s2 pass S = 0 (half-rate)
    int pS;
    int extra_tile_bytes_per_pixel;   // wgrad: the dy tile adds Cb*2 bytes per destination pixel
};

static bool plan_tile(const PlanIn& in, DwTile& p) {
    p.B = in.B; p.C = in.C; p.Ti = in.Ti; p.Hin = in.Hin; p.Win = in.Win; p.To = in.To; p.Ho = in.Ho; p.Wo = in.Wo;
    p.pS = in.pS;
    p.nblk = ceil_div(in.C, 128);
    p.Cb = (ceil_div(in.C, p.nblk) + 7) / 8 * 8;
    p.nblk = ceil_div(in.C, p.Cb);
    p.Gb = p.Cb / 4;
    const int K = in.K, S = in.S, WS = in.WS;
    auto src_extent = [&](int n) { return S == 0 ? n / 2 + (K - 1) / 2 + 1 : (n - 1) * S + K; };
    // Search (tiles_w, Ht): the tile must fit the per-stage budget; prefer the best ratio of useful
    // destination pixels to staged source pixels (halo efficiency), then the larger tile.
    const int budget = 34 * 1024;                       // per stage; 3 stages x 2 CTAs fit in one SM
    const int hstep = (S == 0) ? 2 : 1;
    double best_score = -1.0;
    int best_ht = 0, best_wt = 0;
    for (int tw = 1; tw <= 16; ++tw) {
        int wt = (ceil_div(in.Wo, tw) + WS - 1) / WS * WS;
        if (tw > 1 && wt >= best_wt && best_wt > 0 && (ceil_div(in.Wo, tw - 1) + WS - 1) / WS * WS == wt) continue;
        int wi = src_extent(wt);
        if (wi > 256 || wt > 256) continue;
        for (int ht = hstep; ht <= in.Ho + hstep - 1; ht += hstep) {
            long long bytes = (long long)src_extent(ht) * wi * p.Cb * 2 +
                              (long long)in.extra_tile_bytes_per_pixel * ht * wt * p.Cb * 2;
            if (bytes > budget || src_extent(ht) > 256) break;
            int th = ceil_div(in.Ho, ht);
            int ht_bal = ceil_div(in.Ho, th);
            if (S == 0) ht_bal = (ht_bal + 1) / 2 * 2;
            if (ht_bal != ht) continue;                 // only balanced heights
            int twn = ceil_div(in.Wo, wt);
            double useful = (double)in.Ho * in.Wo;
            double staged = (double)th * twn * src_extent(ht) * wi;
            double dest = (double)th * twn * ht * wt;
            if (S == 2) staged /= 4.0;                  // a stride-2 tile needs 4 source pixels per output anyway
            if (S == 0) staged *= 4.0;
            double score = useful / std::max(staged, dest) + 1e-4 * std::min(bytes, (long long)budget) / budget;
            if (score > best_score) { best_score = score; best_ht = ht; best_wt = wt; }
        }
    }
    if (best_ht == 0) return false;
    p.Wt = best_wt; p.tiles_w = ceil_div(in.Wo, best_wt); p.Wi = src_extent(best_wt);
    p.Ht = best_ht; p.tiles_h = ceil_div(in.Ho, best_ht); p.Hi = src_extent(best_ht);
    p.box_bytes = p.Hi * p.Wi * p.Cb * 2;
    p.box2_bytes = in.extra_tile_bytes_per_pixel ? p.Ht * p.Wt * p.Cb * 2 : 0;
    p.stage_bytes = (p.box_bytes + 127) / 128 * 128 + (p.box2_bytes + 127) / 128 * 128;
    p.stages = std::min(DWT_MAX_STAGES, (108 * 1024) / p.stage_bytes);
    if (p.stages < 2) return false;
    p.flip = 0;
    return true;
}

static int make_map5(CUtensorMap* tm, const void* base, int C, int W, int H, int T, int B, int bc, int bw, int bh) {
    uint64_t dims[5] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)T, (uint64_t)B};
    uint64_t str[5] = {2, (uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2, (uint64_t)T * H * W * C * 2};
    uint32_t box[5] = {(uint32_t)bc, (uint32_t)bw, (uint32_t)bh, 1, 1};
    return make_tmap_bf16(tm, base, 5, dims, str, box, false);
}

// frame bookkeeping: destination frame f has source frame  (f*num + off)/den  when divisible and in range
static bool plan_frames(DwTile& p, int n_dst, int n_src, int num, int off, int den) {
    p.nzf = 0;
    int first = -1, count = 0, step = 0, last = -1;
    for (int f = 0; f < n_dst; ++f) {
        int v = f * num + off;
        bool ok = v >= 0 && v % den == 0 && v / den < n_src;
        if (ok) {
            if (first < 0) first = f;
            else if (step == 0) step = f - last;
            else if (f - last != step) return false;
            last = f; ++count;
        } else {
            if (p.nzf >= DWT_MAX_ZF || f > 255) return false;
            p.zf[p.nzf++] = (unsigned char)f;
        }
    }
    p.f_first = first < 0 ? 0 : first;
    p.f_step = step == 0 ? 1 : step;
    p.f_count = count;
    p.src_first = first < 0 ? 0 : (first * num + off) / den;
    p.src_step = p.f_step * num / den;
    if (count > 1 && (p.f_step * num) % den != 0) return false;
    p.ntiles = (long long)p.B * p.f_count * p.tiles_h * p.tiles_w;
    return true;
}

static dim3 persistent_grid(const DwTile& p) {
    int ctas = std::max(1, ceil_div(148 * 2, p.nblk));
    ctas = (int)std::min<long long>(ctas, std::max<long long>(p.ntiles, 1));
    return dim3(ctas, p.nblk);
}

template <typename KernelT>
static void set_smem_once(KernelT k, std::once_flag& flag) {
    std::call_once(flag, [k] { cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024); });
}

template <int K, int S, int WS>
static bool launch_fwd(const __nv_bfloat16* x, const float* w_tc, __nv_bfloat16* y, const DwDims& d, int pT, int sT,
                       int flip, cudaStream_t st) {
    DwTile p;
    PlanIn in{d.B, d.C, d.T, d.H, d.W, d.To, d.Ho, d.Wo, K, S, WS, d.pH, 0};
    if (!plan_tile(in, p)) return false;
    if (!plan_frames(p, d.To, d.T, sT, -pT, 1)) return false;      // source frame = to*sT - pT
    p.flip = flip;
    CUtensorMap tm;
    if (make_map5(&tm, x, d.C, d.W, d.H, d.T, d.B, p.Cb, p.Wi, p.Hi) != PB_OK) return false;
    static std::once_flag once;
    set_smem_once(dw_fwd_tma_kernel<K, S, WS>, once);
    dw_fwd_tma_kernel<K, S, WS><<<persistent_grid(p), DWT_THREADS, (size_t)p.stages * p.stage_bytes + 128, st>>>(tm, w_tc, y, p);
    return true;
}

template <int K>
static bool launch_dgrad_s2(const __nv_bfloat16* dy, const float* w_tc, __nv_bfloat16* dx, const DwDims& d,
                            cudaStream_t st) {
    constexpr int XS = 4;
    DwTile p;
    PlanIn in{d.B, d.C, d.To, d.Ho, d.Wo, d.T, d.H, d.W, K, 0, XS, d.pH, 0};
    if (!plan_tile(in, p)) return false;
    if (!plan_frames(p, d.T, d.To, 1, d.pT, d.sT)) return false;   // source frame = (t + pT)/sT
    CUtensorMap tm;
    if (make_map5(&tm, dy, d.C, d.Wo, d.Ho, d.To, d.B, p.Cb, p.Wi, p.Hi) != PB_OK) return false;
    static std::once_flag once;
    set_smem_once(dw_dgrad_s2_tma_kernel<K, XS>, once);
    dw_dgrad_s2_tma_kernel<K, XS><<<persistent_grid(p), DWT_THREADS, (size_t)p.stages * p.stage_bytes + 128, st>>>(tm, w_tc, dx, p);
    return true;
}

template <int K, int S, int WS>
static bool launch_wgrad(const __nv_bfloat16* x, const __nv_bfloat16* dy, float* dw_tc, const DwDims& d,
                         cudaStream_t st) {
    DwTile p;
    PlanIn in{d.B, d.C, d.T, d.H, d.W, d.To, d.Ho, d.Wo, K, S, WS, d.pH, 1};
    if (!plan_tile(in, p)) return false;
    if (p.Gb * K > DWT_THREADS) return false;
    if (K * K * p.Cb * 4 > p.stages * p.stage_bytes) return false;
    if (!plan_frames(p, d.To, d.T, d.sT, -d.pT, 1)) return false;
    if (p.f_count == 0) return true;                                 // nothing contributes; dw stays zero
    CUtensorMap tmx, tmd;
    if (make_map5(&tmx, x, d.C, d.W, d.H, d.T, d.B, p.Cb, p.Wi, p.Hi) != PB_OK) return false;
    if (make_map5(&tmd, dy, d.C, d.Wo, d.Ho, d.To, d.B, p.Cb, p.Wt, p.Ht) != PB_OK) return false;
    static std::once_flag once;
    set_smem_once(dw_wgrad_tma_kernel<K, S, WS>, once);
    dw_wgrad_tma_kernel<K, S, WS><<<persistent_grid(p), DWT_THREADS, (size_t)p.stages * p.stage_bytes + 128, st>>>(tmx, tmd, dw_tc, p);
    return true;
}

static bool mobilenet_class(const DwDims& d) {
    return d.kT == 1 && d.kH == d.kW && (d.kH == 3 || d.kH == 5) && d.sH == d.sW && (d.sH == 1 || d.sH == 2) &&
           d.pH == d.pW && d.pH == d.kH / 2 && d.C % 8 == 0;
}
static bool aligned16(const void* a, const void* b) {
    return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0;
}

template <> bool dw_fwd_tiled<__nv_bfloat16>(const __nv_bfloat16* x, const float* w_tc, __nv_bfloat16* y,
                                            const DwDims& d, cudaStream_t st) {
    if (!mobilenet_class(d) || !aligned16(x, y)) return false;
    if (d.kH == 3 && d.sH == 1) return launch_fwd<3, 1, 4>(x, w_tc, y, d, d.pT, d.sT, 0, st);
    if (d.kH == 3 && d.sH == 2) return launch_fwd<3, 2, 4>(x, w_tc, y, d, d.pT, d.sT, 0, st);
    if (d.kH == 5 && d.sH == 1) return launch_fwd<5, 1, 4>(x, w_tc, y, d, d.pT, d.sT, 0, st);
    if (d.kH == 5 && d.sH == 2) return launch_fwd<5, 2, 4>(x, w_tc, y, d, d.pT, d.sT, 0, st);
    return false;
}

// stride-1 input gradient == forward correlation of dy with the flipped filter:
//   dx[t][h][w] = sum_{i,j} dy[t + pT][h + p - i][w + p - j] * w[i][j]
template <> bool dw_dgrad_tiled<__nv_bfloat16>(const __nv_bfloat16* dy, const float* w_tc, __nv_bfloat16* dx,
                                              const DwDims& d, cudaStream_t st) {
    if (!mobilenet_class(d) || !aligned16(dy, dx)) return false;
    if (d.sH == 2) {
        if (d.sT != 1 && d.sT != 2) return false;
        if ((d.H | d.W) < 2) return false;
        return d.kH == 3 ? launch_dgrad_s2<3>(dy, w_tc, dx, d, st) : launch_dgrad_s2<5>(dy, w_tc, dx, d, st);
    }
    if (d.sT != 1) return false;
    DwDims r = d;                      // roles swapped: "input" = dy (To,Ho,Wo), "output" = dx (T,H,W)
    r.T = d.To; r.H = d.Ho; r.W = d.Wo;
    r.To = d.T; r.Ho = d.H; r.Wo = d.W;
    if (d.kH == 3) return launch_fwd<3, 1, 4>(dy, w_tc, dx, r, -d.pT, 1, 1, st);
    return launch_fwd<5, 1, 4>(dy, w_tc, dx, r, -d.pT, 1, 1, st);
}

template <> bool dw_wgrad_tiled<__nv_bfloat16>(const __nv_bfloat16* x, const __nv_bfloat16* dy, float* dw_tc,
                                              const DwDims& d, cudaStream_t st) {
    if (!mobilenet_class(d) || !aligned16(x, dy)) return false;
    if (d.kH == 3 && d.sH == 1) return launch_wgrad<3, 1, 4>(x, dy, dw_tc, d, st);
    if (d.kH == 3 && d.sH == 2) return launch_wgrad<3, 2, 4>(x, dy, dw_tc, d, st);
    if (d.kH == 5 && d.sH == 1) return launch_wgrad<5, 1, 4>(x, dy, dw_tc, d, st);
    if (d.kH == 5 && d.sH == 2) return launch_wgrad<5, 2, 4>(x, dy, dw_tc, d, st);
    return false;
}

template <> bool dw_fwd_tiled<float>(const float*, const float*, float*, const DwDims&, cudaStream_t) { return false; }
template <> bool dw_dgrad_tiled<float>(const float*, const float*, float*, const DwDims&, cudaStream_t) { return false; }
template <> bool dw_wgrad_tiled<float>(const float*, const float*, float*, const DwDims&, cudaStream_t) { return false; }

}  // namespace pb
