// Depthwise (1,k,k) convolution fast paths for bf16 NDHWC activations: TMA-staged halo tiles.
//
// A persistent CTA owns one channel block (<= 128 channels) and walks over (sample, frame, h-tile, w-tile)
// work items.  A dedicated producer warp issues one cp.async.bulk.tensor.5d per item: the halo tile
// [Hi][Wi][Cb] of the source frame (plus, for the weight gradient, the matching dy tile) lands in a
// shared-memory ring guarded by full/empty mbarriers, so the eight consumer warps never wait for each
// other -- only for data.  Out-of-bounds coordinates are zero-filled by the TMA unit: spatial padding costs
// nothing and needs no bounds checks, and the frames created by the reference's scalar temporal padding
// (mobilenet.py:67-75) are written as zeros without reading anything.  Each consumer thread owns 4 channels
// (8-byte vectors) and a strip of output pixels: every staged vector is unpacked once and fed to all the
// outputs of the strip that use it with packed fp32 FMAs (fma.rn.f32x2); 3x3 taps stay in registers, 5x5 taps
// are staged per CTA in shared memory (they spilled otherwise).
//
// A tcgen05 formulation was built and measured (depthwise = GEMM against diag(w_tap), the im2col operand of
// tap (i,j) being the same swizzled halo tile read from a start address shifted by i*Wi+j rows; git history:
// "Depthwise 5x5 stride-1 forward/dgrad on tcgen05").  It is correct, but every tap re-reads the whole tile
// from shared memory (25 passes at 128 B/clk), which bounds it near 2.5 TB/s x the fraction of useful MMA rows:
// 1.8 TB/s on 28x28 planes against 2.1-2.3 TB/s for the kernels below, so it was dropped.
//
//   forward            : y  = conv(x, w)                     (also the stride-1 input gradient: flipped w)
//   input gradient, s=2: dx = gather of dy over the taps whose parity matches
//   weight gradient    : dw accumulated in registers over all tiles of the CTA, reduced once at the end
//                        (shared-memory sums, then one fp32 atomic per tap and channel per CTA)
#include <algorithm>
#include <mutex>

#include "dwconv.cuh"
#include "tc_common.cuh"

namespace pb {

using namespace tc;

constexpr int DWT_MAX_STAGES = 4;
constexpr int DWT_CONSUMERS = 224;                 // 7 consumer warps (2 CTAs x 256 threads x 128 regs per SM)
constexpr int DWT_THREADS = DWT_CONSUMERS + 32;    // + 1 producer warp
constexpr int DWT_MAX_ZF = 64;

struct DwTile {
    int B, C;
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
    int kT, padT;              // temporal taps (forward kernel only) and frames of padding before the first one
    int src_T, sbuf_T;         // frames of the source tensor / of the stream buffer (0: not streaming)
    // forward only: fused global average pool of the outputs (squeeze-excite), pool[b][c] += pool_scale * y
    float* pool;
    float pool_scale;
    int contig;                // tiles of a CTA are consecutive (few sample changes -> few pool flushes) instead of strided
};

// n-th tile of this CTA and how many it has
__device__ __forceinline__ long long cta_tile(const DwTile& p, long long n) {
    if (!p.contig) return blockIdx.x + n * (long long)gridDim.x;
    const long long per = p.ntiles / gridDim.x, rem = p.ntiles % gridDim.x;
    return blockIdx.x * per + (blockIdx.x < rem ? blockIdx.x : rem) + n;
}
__device__ __forceinline__ long long cta_tile_count(const DwTile& p) {
    if (!p.contig) return p.ntiles > blockIdx.x ? (p.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    return p.ntiles / gridDim.x + (blockIdx.x < p.ntiles % gridDim.x ? 1 : 0);
}

__device__ __forceinline__ void ffma2(float2& d, const float2 a, const float2 b) {
    unsigned long long dd = *reinterpret_cast<unsigned long long*>(&d);
    asm("fma.rn.f32x2 %0, %1, %2, %0;"
        : "+l"(dd)
        : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)));
    d = *reinterpret_cast<float2*>(&dd);
}
__device__ __forceinline__ float2 unpack2(uint32_t u) { return make_float2(bf16_lo(u), bf16_hi(u)); }
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}

__device__ __forceinline__ float4 lds128f(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

// shared bookkeeping of the tile pipeline
struct TileCtx {
    uint64_t full[DWT_MAX_STAGES];
    uint64_t empty[DWT_MAX_STAGES];
    int4 coord[DWT_MAX_STAGES];            // (b, n-th valid frame, h0, w0) of the tile in each stage
};

__device__ __forceinline__ int4 decode_tile(const DwTile& p, long long t) {
    int4 c;
    int tw = (int)(t % p.tiles_w); t /= p.tiles_w;
    int th = (int)(t % p.tiles_h); t /= p.tiles_h;
    c.y = (int)(t % p.f_count);
    c.x = (int)(t / p.f_count);
    c.z = th * p.Ht; c.w = tw * p.Wt;
    return c;
}

// zero the destination frames that have no source frame (consumer threads of CTAs with blockIdx.y == 0)
__device__ __forceinline__ void zero_frames(const DwTile& p, __nv_bfloat16* y, int tid) {
    if (p.nzf <= 0) return;
    // every CTA of the grid (all channel blocks) takes whole frames round-robin; a frame is contiguous
    const int frame16 = (int)((long long)p.Ho * p.Wo * p.C / 8);
    const int nframes = p.B * p.nzf;
    const int ncta = gridDim.x * gridDim.y;
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    for (int q = blockIdx.y * gridDim.x + blockIdx.x; q < nframes; q += ncta) {
        const int f = p.zf[q % p.nzf];
        const int b = q / p.nzf;
        uint4* dst = reinterpret_cast<uint4*>(y) + ((long long)b * p.To + f) * frame16;
        for (int e = tid; e < frame16; e += DWT_CONSUMERS) dst[e] = z;
    }
}

__device__ __forceinline__ void pipeline_init(TileCtx& cx, const DwTile& p) {
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(&cx.full[s], 1); mbar_init(&cx.empty[s], DWT_CONSUMERS / 32); }
        fence_barrier_init();
    }
    __syncthreads();
}

// 5x5 filters: taps of this CTA's channel block as fp32 [tap][Cb] in shared memory (all threads, before the
// role split); `flip` reverses the tap order (stride-1 input gradient).
template <int TAPS>
__device__ __forceinline__ void stage_taps(float* wsm, const float* __restrict__ w_tc, const DwTile& p, int c_base, int flip) {
    for (int e = threadIdx.x; e < TAPS * p.Cb; e += DWT_THREADS) {
        const int tap = e / p.Cb, c = e - tap * p.Cb;
        const int src = flip ? TAPS - 1 - tap : tap;
        wsm[e] = (c_base + c < p.C) ? w_tc[(long long)src * p.C + c_base + c] : 0.f;
    }
    __syncthreads();
}

// Producer warp (one elected lane): `load(stage_ptr, full_barrier, coord)` issues the TMA copies of one tile.
template <typename LoadFn>
__device__ __forceinline__ void producer_loop(TileCtx& cx, const DwTile& p, uint8_t* ring, long long my_tiles,
                                              uint32_t tx_bytes, LoadFn load) {
    for (long long n = 0; n < my_tiles; ++n) {
        const int s = (int)(n % p.stages);
        if (n >= p.stages) mbar_wait(&cx.empty[s], (uint32_t)(((n / p.stages) - 1) & 1));
        const int4 c = decode_tile(p, cta_tile(p, n));
        cx.coord[s] = c;
        mbar_expect_tx(&cx.full[s], tx_bytes);          // release: publishes coord[s] as well
        load(ring + (size_t)s * p.stage_bytes, &cx.full[s], c);
    }
}

// per-thread work split: thread = (slot, group); items of a tile are (row, strip) pairs, item = slot + n*ppp
struct ItemIter {
    int hl, strip, d_h, d_s, nstrips;
    __device__ __forceinline__ void start(int slot, int ppp, int nstrips_) {
        nstrips = nstrips_;
        hl = slot / nstrips; strip = slot % nstrips;
        d_h = ppp / nstrips; d_s = ppp % nstrips;
    }
    __device__ __forceinline__ void next() {
        hl += d_h; strip += d_s;
        if (strip >= nstrips) { strip -= nstrips; ++hl; }
    }
};

// ------------------------------------------------------------------------------------------------
// forward (and stride-1 dgrad): destination tile [Ht][Wt], source halo [(Ht-1)S+K][(Wt-1)S+K]
// ------------------------------------------------------------------------------------------------
template <int K, int S, int WS>
__global__ void __launch_bounds__(DWT_THREADS, 2)
dw_fwd_tma_kernel(const __grid_constant__ CUtensorMap tmX, const float* __restrict__ w_tc,
                  __nv_bfloat16* __restrict__ y, const DwTile p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ TileCtx cx;
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* ring = smem_raw + (((raw + 127u) & ~127u) - raw);
    const int tid = threadIdx.x;
    const int c_base = blockIdx.y * p.Cb;
    const int cb_bytes = p.Cb * 2;
    pdl_trigger();
    pipeline_init(cx, p);
    pdl_wait();                 // barrier setup above overlaps the previous kernel's tail
    if constexpr (K == 5) stage_taps<K * K>(reinterpret_cast<float*>(ring + (size_t)p.stages * p.stage_bytes), w_tc, p, c_base, p.flip);
    const long long my_tiles = cta_tile_count(p);

    if (tid >= DWT_CONSUMERS) {
        if (tid == DWT_CONSUMERS) {
            tma_prefetch_desc(&tmX);
            producer_loop(cx, p, ring, my_tiles, (uint32_t)p.box_bytes, [&](uint8_t* st, uint64_t* bar, const int4& c) {
                tma_load_5d(st, &tmX, bar, c_base, c.w * S - p.pS, c.z * S - p.pS, p.src_first + c.y * p.src_step, c.x);
            });
        }
        return;
    }
    zero_frames(p, y, tid);

    const int ppp = DWT_CONSUMERS / p.Gb;
    const int g = tid % p.Gb;
    const bool active = tid < ppp * p.Gb && (c_base + g * 4) < p.C;
    const int slot = tid / p.Gb;
    // K == 3: taps in registers as fp32 pairs (36 registers).  K == 5: 100 fp32 (or 50 packed) registers do not
    // fit next to the window and the accumulators under the 128-register cap of 2 CTAs/SM (they spilled), so the
    // taps are staged once per CTA as fp32 [tap][Cb] in shared memory and fetched with one LDS.128 per use.
    constexpr bool SMEMW = (K == 5);
    float2 wv[SMEMW ? 1 : K * K][2];
    if constexpr (!SMEMW) {
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
                wv[t][c] = make_float2(lo, hi);
            }
        }
    }
    const uint32_t wsm_g = smem_u32(ring + (size_t)p.stages * p.stage_bytes) + (uint32_t)(g * 16);
    const int nstrips = p.Wt / WS;
    const int items = p.Ht * nstrips;
    const uint32_t ring_u32 = smem_u32(ring);
    const uint32_t row_bytes = (uint32_t)(p.Wi * cb_bytes);
    const int lane = tid & 31;

    // fused average pool: per-thread sums of the ROUNDED outputs of its 4 channels, flushed with one atomic per channel
    // whenever the sample changes (tiles of a CTA are consecutive when p.pool is set: one or two flushes per CTA)
    float ps0 = 0.f, ps1 = 0.f, ps2 = 0.f, ps3 = 0.f;
    int pool_b = -1;
    auto pool_flush = [&]() {
        if (pool_b >= 0 && active) {
            float* dst = p.pool + (long long)pool_b * p.C + c_base + g * 4;
            atomicAdd(dst + 0, ps0 * p.pool_scale); atomicAdd(dst + 1, ps1 * p.pool_scale);
            atomicAdd(dst + 2, ps2 * p.pool_scale); atomicAdd(dst + 3, ps3 * p.pool_scale);
        }
        ps0 = ps1 = ps2 = ps3 = 0.f;
    };

    for (long long n = 0; n < my_tiles; ++n) {
        const int s = (int)(n % p.stages);
        mbar_wait(&cx.full[s], (uint32_t)((n / p.stages) & 1));
        const int4 c = cx.coord[s];
        const int to = p.f_first + c.y * p.f_step;
        const uint32_t tile = ring_u32 + (uint32_t)(s * p.stage_bytes) + (uint32_t)(g * 8);
        if (p.pool && c.x != pool_b) { pool_flush(); pool_b = c.x; }
        if (active) {
            ItemIter it;
            it.start(slot, ppp, nstrips);
            for (int idx = slot; idx < items; idx += ppp, it.next()) {
                const int ho = c.z + it.hl;
                if (ho >= p.Ho) break;
                float2 acc[WS][2];
#pragma unroll
                for (int o = 0; o < WS; ++o) { acc[o][0] = make_float2(0.f, 0.f); acc[o][1] = make_float2(0.f, 0.f); }
                uint32_t rowa = tile + (uint32_t)(it.hl * S) * row_bytes + (uint32_t)(it.strip * WS * S * cb_bytes);
#pragma unroll
                for (int i = 0; i < K; ++i, rowa += row_bytes) {
                    uint32_t a = rowa;
                    if constexpr (!SMEMW) {
#pragma unroll
                        for (int j = 0; j < (WS - 1) * S + K; ++j, a += cb_bytes) {
                            const uint2 u = lds64(a);
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
                        for (int j = 0; j < NJ; ++j, a += cb_bytes) {
                            const uint2 u = lds64(a);
                            win[j][0] = unpack2(u.x); win[j][1] = unpack2(u.y);
                        }
#pragma unroll
                        for (int t = 0; t < K; ++t) {
                            const float4 w4 = lds128f(wsm_g + (uint32_t)((i * K + t) * p.Cb * 4));
                            const float2 w0 = make_float2(w4.x, w4.y), w1 = make_float2(w4.z, w4.w);
#pragma unroll
                            for (int o = 0; o < WS; ++o) {
                                ffma2(acc[o][0], win[o * S + t][0], w0);
                                ffma2(acc[o][1], win[o * S + t][1], w1);
                            }
                        }
                    }
                }
                const int wo0 = c.w + it.strip * WS;
                __nv_bfloat16* yp = y + ((((long long)c.x * p.To + to) * p.Ho + ho) * p.Wo + wo0) * p.C + c_base + g * 4;
#pragma unroll
                for (int o = 0; o < WS; ++o) {
                    if (wo0 + o < p.Wo) {
                        uint2 out;
                        out.x = pack_bf16x2(acc[o][0].x, acc[o][0].y);
                        out.y = pack_bf16x2(acc[o][1].x, acc[o][1].y);
                        *reinterpret_cast<uint2*>(yp) = out;
                        if (p.pool) {
                            ps0 += bf16_lo(out.x); ps1 += bf16_hi(out.x);
                            ps2 += bf16_lo(out.y); ps3 += bf16_hi(out.y);
                        }
                    }
                    yp += p.C;
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&cx.empty[s]);          // this warp is done reading stage s
    }
    if (p.pool) pool_flush();
}

// ------------------------------------------------------------------------------------------------
// (kT,3,3) forward with temporal stride 1: MoviNetBottleneck.conv (movinet.py:52-61) and its causal streaming
// form (CausalConv3d, movinet.py:23-39).  A stage holds the kT source frames of one destination tile, fetched
// frame by frame: frame to - padT + kt of x, or -- streaming -- of the stream buffer when that index is negative
// (the tail of the previous chunk, resident in HBM); frames outside both are zero-filled by the TMA unit, which
// is the symmetric temporal padding of the non-causal layers.  All kT*9 taps sit in shared memory as fp32 [tap][Cb].
// ------------------------------------------------------------------------------------------------
template <int KT, int S, int WS>
__global__ void __launch_bounds__(DWT_THREADS, 2)
dw_fwd3d_tma_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmS,
                    const float* __restrict__ w_tc, __nv_bfloat16* __restrict__ y, const DwTile p) {
    constexpr int K = 3;
    extern __shared__ uint8_t smem_raw[];
    __shared__ TileCtx cx;
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* ring = smem_raw + (((raw + 127u) & ~127u) - raw);
    const int tid = threadIdx.x;
    const int c_base = blockIdx.y * p.Cb;
    const int cb_bytes = p.Cb * 2;
    const int frame_bytes = (p.box_bytes + 127) / 128 * 128;
    pdl_trigger();
    pipeline_init(cx, p);
    pdl_wait();
    stage_taps<KT * K * K>(reinterpret_cast<float*>(ring + (size_t)p.stages * p.stage_bytes), w_tc, p, c_base, 0);
    const long long my_tiles = p.ntiles > blockIdx.x ? (p.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (tid >= DWT_CONSUMERS) {
        if (tid == DWT_CONSUMERS) {
            tma_prefetch_desc(&tmX);
            if (p.sbuf_T > 0) tma_prefetch_desc(&tmS);
            producer_loop(cx, p, ring, my_tiles, (uint32_t)(KT * p.box_bytes), [&](uint8_t* st, uint64_t* bar, const int4& c) {
#pragma unroll
                for (int kt = 0; kt < KT; ++kt) {
                    const int fi = c.y - p.padT + kt;                       // source frame of tap kt (c.y = destination frame)
                    if (fi < 0 && p.sbuf_T > 0)
                        tma_load_5d(st + kt * frame_bytes, &tmS, bar, c_base, c.w * S - p.pS, c.z * S - p.pS, p.sbuf_T + fi, c.x);
                    else
                        tma_load_5d(st + kt * frame_bytes, &tmX, bar, c_base, c.w * S - p.pS, c.z * S - p.pS, fi, c.x);
                }
            });
        }
        return;
    }

    const int ppp = DWT_CONSUMERS / p.Gb;
    const int g = tid % p.Gb;
    const bool active = tid < ppp * p.Gb && (c_base + g * 4) < p.C;
    const int slot = tid / p.Gb;
    const uint32_t wsm_g = smem_u32(ring + (size_t)p.stages * p.stage_bytes) + (uint32_t)(g * 16);
    const int nstrips = p.Wt / WS;
    const int items = p.Ht * nstrips;
    const uint32_t ring_u32 = smem_u32(ring);
    const uint32_t row_bytes = (uint32_t)(p.Wi * cb_bytes);
    const int lane = tid & 31;

    for (long long n = 0; n < my_tiles; ++n) {
        const int s = (int)(n % p.stages);
        mbar_wait(&cx.full[s], (uint32_t)((n / p.stages) & 1));
        const int4 c = cx.coord[s];
        const uint32_t tile = ring_u32 + (uint32_t)(s * p.stage_bytes) + (uint32_t)(g * 8);
        if (active) {
            ItemIter it;
            it.start(slot, ppp, nstrips);
            for (int idx = slot; idx < items; idx += ppp, it.next()) {
                const int ho = c.z + it.hl;
                if (ho >= p.Ho) break;
                float2 acc[WS][2];
#pragma unroll
                for (int o = 0; o < WS; ++o) { acc[o][0] = make_float2(0.f, 0.f); acc[o][1] = make_float2(0.f, 0.f); }
#pragma unroll 1
                for (int kt = 0; kt < KT; ++kt) {
                    uint32_t rowa = tile + (uint32_t)(kt * frame_bytes) + (uint32_t)(it.hl * S) * row_bytes +
                                    (uint32_t)(it.strip * WS * S * cb_bytes);
#pragma unroll
                    for (int i = 0; i < K; ++i, rowa += row_bytes) {
                        constexpr int NJ = (WS - 1) * S + K;
                        float2 win[NJ][2];
                        uint32_t a = rowa;
#pragma unroll
                        for (int j = 0; j < NJ; ++j, a += cb_bytes) {
                            const uint2 u = lds64(a);
                            win[j][0] = unpack2(u.x); win[j][1] = unpack2(u.y);
                        }
#pragma unroll
                        for (int t = 0; t < K; ++t) {
                            const float4 w4 = lds128f(wsm_g + (uint32_t)(((kt * K + i) * K + t) * p.Cb * 4));
                            const float2 w0 = make_float2(w4.x, w4.y), w1 = make_float2(w4.z, w4.w);
#pragma unroll
                            for (int o = 0; o < WS; ++o) {
                                ffma2(acc[o][0], win[o * S + t][0], w0);
                                ffma2(acc[o][1], win[o * S + t][1], w1);
                            }
                        }
                    }
                }
                const int wo0 = c.w + it.strip * WS;
                __nv_bfloat16* yp = y + ((((long long)c.x * p.To + c.y) * p.Ho + ho) * p.Wo + wo0) * p.C + c_base + g * 4;
#pragma unroll
                for (int o = 0; o < WS; ++o) {
                    if (wo0 + o < p.Wo) {
                        uint2 out;
                        out.x = pack_bf16x2(acc[o][0].x, acc[o][0].y);
                        out.y = pack_bf16x2(acc[o][1].x, acc[o][1].y);
                        *reinterpret_cast<uint2*>(yp) = out;
                    }
                    yp += p.C;
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&cx.empty[s]);
    }
}

// ------------------------------------------------------------------------------------------------
// stride-2 input gradient: destination tile of dx [Ht][Wt] (Ht, Wt even, tile origin even), source halo of
// dy rows h0/2 - P/2 .. , same for columns.
//   dx[h][w] = sum_{i,j : (h+p-i), (w+p-j) even} dy[(h+p-i)/2][(w+p-j)/2] * w[i][j]
// ------------------------------------------------------------------------------------------------
template <int K, int XS>
__global__ void __launch_bounds__(DWT_THREADS, 2)
dw_dgrad_s2_tma_kernel(const __grid_constant__ CUtensorMap tmD, const float* __restrict__ w_tc,
                       __nv_bfloat16* __restrict__ dx, const DwTile p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ TileCtx cx;
    constexpr int P = K / 2;
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* ring = smem_raw + (((raw + 127u) & ~127u) - raw);
    const int tid = threadIdx.x;
    const int c_base = blockIdx.y * p.Cb;
    const int cb_bytes = p.Cb * 2;
    pdl_trigger();
    pipeline_init(cx, p);
    pdl_wait();                 // barrier setup above overlaps the previous kernel's tail
    if constexpr (K == 5) stage_taps<K * K>(reinterpret_cast<float*>(ring + (size_t)p.stages * p.stage_bytes), w_tc, p, c_base, 0);
    const long long my_tiles = p.ntiles > blockIdx.x ? (p.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (tid >= DWT_CONSUMERS) {
        if (tid == DWT_CONSUMERS) {
            tma_prefetch_desc(&tmD);
            producer_loop(cx, p, ring, my_tiles, (uint32_t)p.box_bytes, [&](uint8_t* st, uint64_t* bar, const int4& c) {
                tma_load_5d(st, &tmD, bar, c_base, c.w / 2 - P / 2, c.z / 2 - P / 2, p.src_first + c.y * p.src_step, c.x);
            });
        }
        return;
    }
    zero_frames(p, dx, tid);

    const int ppp = DWT_CONSUMERS / p.Gb;
    const int g = tid % p.Gb;
    const bool active = tid < ppp * p.Gb && (c_base + g * 4) < p.C;
    const int slot = tid / p.Gb;
    // K == 3: packed bf16 pairs in registers; K == 5: fp32 taps in shared memory (see dw_fwd_tma_kernel)
    constexpr bool SMEMW = (K == 5);
    uint32_t wp[SMEMW ? 1 : K * K][2];
    if constexpr (!SMEMW) {
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
    }
    const uint32_t wsm_g = smem_u32(ring + (size_t)p.stages * p.stage_bytes) + (uint32_t)(g * 16);
    const int nstrips = p.Wt / XS;
    const int items = p.Ht * nstrips;
    const uint32_t ring_u32 = smem_u32(ring);
    const uint32_t row_bytes = (uint32_t)(p.Wi * cb_bytes);
    const int lane = tid & 31;

    for (long long n = 0; n < my_tiles; ++n) {
        const int s = (int)(n % p.stages);
        mbar_wait(&cx.full[s], (uint32_t)((n / p.stages) & 1));
        const int4 c = cx.coord[s];
        const int t_dst = p.f_first + c.y * p.f_step;
        const uint32_t tile = ring_u32 + (uint32_t)(s * p.stage_bytes) + (uint32_t)(g * 8);
        if (active) {
            ItemIter it;
            it.start(slot, ppp, nstrips);
            for (int idx = slot; idx < items; idx += ppp, it.next()) {
                const int hl = it.hl;
                const int h = c.z + hl;
                if (h >= p.Ho) break;
                float2 acc[XS][2];
#pragma unroll
                for (int o = 0; o < XS; ++o) { acc[o][0] = make_float2(0.f, 0.f); acc[o][1] = make_float2(0.f, 0.f); }
                // tile origin even => parity of h equals parity of hl; the strip starts at an even column, so
                // column parities are compile-time per o.  Local dy row/col of tap (i,j): (x + P - j)/2 + P/2.
#pragma unroll
                for (int i = 0; i < K; ++i) {
                    if (((hl + P - i) & 1) != 0) continue;
                    const int rl = ((hl + P - i) >> 1) + P / 2;
                    uint32_t a = tile + (uint32_t)rl * row_bytes + (uint32_t)((it.strip * XS / 2) * cb_bytes);
                    constexpr int NC = XS / 2 + P / 2 + 1;               // dy columns a strip can touch
                    float2 win[NC][2];
#pragma unroll
                    for (int cc = 0; cc < NC; ++cc, a += cb_bytes) {
                        const uint2 u = lds64(a);
                        win[cc][0] = unpack2(u.x); win[cc][1] = unpack2(u.y);
                    }
#pragma unroll
                    for (int o = 0; o < XS; ++o) {
#pragma unroll
                        for (int j = 0; j < K; ++j) {
                            if (((o + P - j) & 1) == 0) {
                                constexpr int dummy = 0; (void)dummy;
                                const int cl = ((o + P - j) >> 1) + P / 2;
                                float2 w0, w1;
                                if constexpr (SMEMW) {
                                    const float4 w4 = lds128f(wsm_g + (uint32_t)((i * K + j) * p.Cb * 4));
                                    w0 = make_float2(w4.x, w4.y); w1 = make_float2(w4.z, w4.w);
                                } else {
                                    w0 = unpack2(wp[i * K + j][0]); w1 = unpack2(wp[i * K + j][1]);
                                }
                                ffma2(acc[o][0], win[cl][0], w0);
                                ffma2(acc[o][1], win[cl][1], w1);
                            }
                        }
                    }
                }
                const int x0 = c.w + it.strip * XS;
                __nv_bfloat16* xp = dx + ((((long long)c.x * p.To + t_dst) * p.Ho + h) * p.Wo + x0) * p.C + c_base + g * 4;
#pragma unroll
                for (int o = 0; o < XS; ++o) {
                    if (x0 + o < p.Wo) {
                        uint2 out;
                        out.x = pack_bf16x2(acc[o][0].x, acc[o][0].y);
                        out.y = pack_bf16x2(acc[o][1].x, acc[o][1].y);
                        *reinterpret_cast<uint2*>(xp) = out;
                    }
                    xp += p.C;
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&cx.empty[s]);
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
    __shared__ TileCtx cx;
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* ring = smem_raw + (((raw + 127u) & ~127u) - raw);
    const int tid = threadIdx.x;
    const int c_base = blockIdx.y * p.Cb;
    const int cb_bytes = p.Cb * 2;
    const int x_bytes = (p.box_bytes + 127) / 128 * 128;       // dy tile follows the x tile in a stage
    pdl_trigger();
    pipeline_init(cx, p);
    pdl_wait();                 // barrier setup above overlaps the previous kernel's tail
    const long long my_tiles = p.ntiles > blockIdx.x ? (p.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (tid == DWT_CONSUMERS) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmD);
        producer_loop(cx, p, ring, my_tiles, (uint32_t)(p.box_bytes + p.box2_bytes),
                      [&](uint8_t* st, uint64_t* bar, const int4& c) {
            tma_load_5d(st, &tmX, bar, c_base, c.w * S - p.pS, c.z * S - p.pS, p.src_first + c.y * p.src_step, c.x);
            tma_load_5d(st + x_bytes, &tmD, bar, c_base, c.w, c.z, p.f_first + c.y * p.f_step, c.x);
        });
    }
    const int lanes = p.Gb * K;                         // (group, filter row) combinations
    const int ppp = DWT_CONSUMERS / lanes;
    const int g = tid % p.Gb;
    const int i = (tid / p.Gb) % K;
    const bool consumer = tid < DWT_CONSUMERS;
    const bool active = consumer && tid < ppp * lanes && (c_base + g * 4) < p.C;
    const int slot = tid / lanes;
    float2 acc[K][2];
#pragma unroll
    for (int t = 0; t < K; ++t) { acc[t][0] = make_float2(0.f, 0.f); acc[t][1] = make_float2(0.f, 0.f); }
    const int nstrips = p.Wt / WS;
    const int items = p.Ht * nstrips;
    const uint32_t ring_u32 = smem_u32(ring);
    const uint32_t row_bytes = (uint32_t)(p.Wi * cb_bytes);
    const uint32_t drow_bytes = (uint32_t)(p.Wt * cb_bytes);
    const int lane = tid & 31;

    if (consumer) {
        for (long long n = 0; n < my_tiles; ++n) {
            const int s = (int)(n % p.stages);
            mbar_wait(&cx.full[s], (uint32_t)((n / p.stages) & 1));
            const uint32_t xt = ring_u32 + (uint32_t)(s * p.stage_bytes) + (uint32_t)(g * 8);
            const uint32_t dt = xt + (uint32_t)x_bytes;
            if (active) {
                ItemIter it;
                it.start(slot, ppp, nstrips);
                for (int idx = slot; idx < items; idx += ppp, it.next()) {
                    uint32_t da = dt + (uint32_t)it.hl * drow_bytes + (uint32_t)(it.strip * WS * cb_bytes);
                    uint32_t xa = xt + (uint32_t)(it.hl * S + i) * row_bytes + (uint32_t)(it.strip * WS * S * cb_bytes);
                    float2 dyv[WS][2];
#pragma unroll
                    for (int o = 0; o < WS; ++o, da += cb_bytes) {
                        const uint2 u = lds64(da);
                        dyv[o][0] = unpack2(u.x); dyv[o][1] = unpack2(u.y);
                    }
#pragma unroll
                    for (int j = 0; j < (WS - 1) * S + K; ++j, xa += cb_bytes) {
                        const uint2 u = lds64(xa);
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
            __syncwarp();
            if (lane == 0) mbar_arrive(&cx.empty[s]);
        }
    }
    // every issued tile has been consumed: the ring can be reused as scratch for the CTA-level reduction
    __syncthreads();
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
    int To, Ho, Wo;        // tiled (destination) tensor
    int K, S, WS;          // S: source step per destination pixel (1 or 2); dgrad-s2 passes S = 0 (half-rate)
    int pS;
    int with_dst_tile;     // wgrad: the dy tile is staged too
    int reserve;           // shared memory kept for other uses (5x5 taps), bytes
    int frames = 1;        // source frames staged per tile (temporal taps of the (kT,3,3) kernels)
};

static bool plan_tile(const PlanIn& in, DwTile& p) {
    p.B = in.B; p.C = in.C; p.To = in.To; p.Ho = in.Ho; p.Wo = in.Wo;
    p.pS = in.pS;
    p.nblk = ceil_div(in.C, in.frames > 1 ? 64 : 128);     // kT source frames per stage: narrower blocks, larger tiles
    p.Cb = (ceil_div(in.C, p.nblk) + 7) / 8 * 8;
    if (in.with_dst_tile && in.C > 128) {
        // weight gradient: a thread is (4 channels, filter row), so only multiples of Gb*K threads work; pick the
        // channel block (>= 64 channels: 128-byte TMA rows) that keeps most of the consumer threads busy
        double best = 0.0;
        for (int cb = 64; cb <= 128; cb += 8) {
            const int lanes = cb / 4 * in.K;
            if (lanes > DWT_CONSUMERS) continue;
            const int nb = ceil_div(in.C, cb);
            const double util = (double)(DWT_CONSUMERS / lanes * lanes) / DWT_CONSUMERS * in.C / ((double)nb * cb);
            if (util > best + 1e-9) { best = util; p.Cb = cb; }
        }
    }
    p.nblk = ceil_div(in.C, p.Cb);
    p.Gb = p.Cb / 4;
    const int K = in.K, S = in.S, WS = in.WS;
    auto src_extent = [&](int n) { return S == 0 ? n / 2 + (K - 1) / 2 + 1 : (n - 1) * S + K; };
    // Search (tiles_w, Ht): the tile must fit the per-stage budget; prefer the best ratio of useful
    // destination pixels to staged source pixels (halo efficiency), then the larger tile.
    // per stage: 3 stages x 2 CTAs fit in one SM; the 5x5 kernels (issue-bound, wide halos) trade the third stage
    // for larger tiles, i.e. less halo re-fetch
    const int budget = in.frames > 1 ? (104 * 1024 - in.reserve) / 2 : (in.K == 5 ? 46 : 34) * 1024;
    const int hstep = (S == 0) ? 2 : 1;
    double best_score = -1.0;
    int best_ht = 0, best_wt = 0;
    for (int tw = 1; tw <= 32; ++tw) {
        int wt = (ceil_div(in.Wo, tw) + WS - 1) / WS * WS;
        int wi = src_extent(wt);
        if (wi > 256 || wt > 256) continue;
        for (int ht = hstep; ht <= in.Ho + hstep - 1; ht += hstep) {
            long long bytes = (long long)in.frames * src_extent(ht) * wi * p.Cb * 2 + (long long)in.with_dst_tile * ht * wt * p.Cb * 2;
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
            double score = useful / std::max(staged, dest) + 1e-4 * (double)bytes / budget;
            if (score > best_score) { best_score = score; best_ht = ht; best_wt = wt; }
        }
    }
    if (best_ht == 0) return false;
    p.Wt = best_wt; p.tiles_w = ceil_div(in.Wo, best_wt); p.Wi = src_extent(best_wt);
    p.Ht = best_ht; p.tiles_h = ceil_div(in.Ho, best_ht); p.Hi = src_extent(best_ht);
    p.box_bytes = p.Hi * p.Wi * p.Cb * 2;
    p.box2_bytes = in.with_dst_tile ? p.Ht * p.Wt * p.Cb * 2 : 0;
    p.stage_bytes = in.frames * ((p.box_bytes + 127) / 128 * 128) + (p.box2_bytes + 127) / 128 * 128;
    p.stages = std::min(DWT_MAX_STAGES, (108 * 1024 - in.reserve) / p.stage_bytes);
    if (p.stages < 2) return false;
    p.flip = 0;
    return true;
}

static int make_map5(CUtensorMap* tm, const void* base, int C, int W, int H, int T, int B, int bc, int bw, int bh) {
    uint64_t dims[5] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)T, (uint64_t)B};
    uint64_t str[5] = {2, (uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2, (uint64_t)T * H * W * C * 2};
    uint32_t box[5] = {(uint32_t)bc, (uint32_t)bw, (uint32_t)bh, 1, 1};
    return make_tmap_bf16(tm, base, 5, dims, str, box, 0);
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
    if (count > 1 && (p.f_step * num) % den != 0) return false;
    p.src_step = p.f_step * num / den;
    p.ntiles = (long long)p.B * p.f_count * p.tiles_h * p.tiles_w;
    return true;
}

static dim3 persistent_grid(const DwTile& p) {
    int ctas = std::max(1, (148 * 2) / p.nblk);           // never more than one resident wave
    ctas = (int)std::min<long long>(ctas, std::max<long long>(p.ntiles, 1));
    return dim3(ctas, p.nblk);
}

template <typename KernelT>
static void set_smem_once(KernelT k, unsigned long long& done) {
    (void)ensure_dyn_smem(k, 112 * 1024, &done);      // a failure surfaces as a launch error (caller's PB_CHECK_LAUNCH)
}

template <int K, int S, int WS>
static bool launch_fwd(const __nv_bfloat16* x, const float* w_tc, __nv_bfloat16* y, const DwDims& d, int pT, int sT,
                       int flip, cudaStream_t st, float* pool = nullptr) {
    DwTile p{};
    constexpr int TAPS_BYTES = K == 5 ? K * K * 128 * 4 : 0;           // fp32 [tap][Cb <= 128]
    PlanIn in{d.B, d.C, d.To, d.Ho, d.Wo, K, S, WS, d.pH, 0, TAPS_BYTES};
    if (!plan_tile(in, p)) return false;
    if (!plan_frames(p, d.To, d.T, sT, -pT, 1)) return false;      // source frame = to*sT - pT
    p.flip = flip;
    p.pool = pool; p.contig = pool ? 1 : 0;
    p.pool_scale = 1.0f / ((float)d.To * (float)d.Ho * (float)d.Wo);
    CUtensorMap tm;
    if (make_map5(&tm, x, d.C, d.W, d.H, d.T, d.B, p.Cb, p.Wi, p.Hi) != PB_OK) return false;
    static unsigned long long once = 0;
    set_smem_once(dw_fwd_tma_kernel<K, S, WS>, once);
    (void)launch_pdl(dw_fwd_tma_kernel<K, S, WS>, dim3(persistent_grid(p)), dim3(DWT_THREADS),
                     (size_t)p.stages * p.stage_bytes + 128 + TAPS_BYTES, st,
                       tm, w_tc, y, p);   // errors: caller's PB_CHECK_LAUNCH
    return true;
}

template <int K>
static bool launch_dgrad_s2(const __nv_bfloat16* dy, const float* w_tc, __nv_bfloat16* dx, const DwDims& d,
                            cudaStream_t st) {
    constexpr int XS = 4;
    DwTile p{};
    constexpr int TAPS_BYTES = K == 5 ? K * K * 128 * 4 : 0;
    PlanIn in{d.B, d.C, d.T, d.H, d.W, K, 0, XS, d.pH, 0, TAPS_BYTES};
    if (!plan_tile(in, p)) return false;
    if (!plan_frames(p, d.T, d.To, 1, d.pT, d.sT)) return false;   // source frame = (t + pT)/sT
    CUtensorMap tm;
    if (make_map5(&tm, dy, d.C, d.Wo, d.Ho, d.To, d.B, p.Cb, p.Wi, p.Hi) != PB_OK) return false;
    static unsigned long long once = 0;
    set_smem_once(dw_dgrad_s2_tma_kernel<K, XS>, once);
    (void)launch_pdl(dw_dgrad_s2_tma_kernel<K, XS>, dim3(persistent_grid(p)), dim3(DWT_THREADS),
                     (size_t)p.stages * p.stage_bytes + 128 + TAPS_BYTES, st,
                       tm, w_tc, dx, p);
    return true;
}

template <int K, int S, int WS>
static bool launch_wgrad(const __nv_bfloat16* x, const __nv_bfloat16* dy, float* dw_tc, const DwDims& d,
                         cudaStream_t st) {
    DwTile p{};
    PlanIn in{d.B, d.C, d.To, d.Ho, d.Wo, K, S, WS, d.pH, 1, 0};
    if (!plan_tile(in, p)) return false;
    if (p.Gb * K > DWT_CONSUMERS) return false;
    if (K * K * p.Cb * 4 > p.stages * p.stage_bytes) return false;
    if (!plan_frames(p, d.To, d.T, d.sT, -d.pT, 1)) return false;
    if (p.f_count == 0) return true;                                 // nothing contributes; dw stays zero
    CUtensorMap tmx, tmd;
    if (make_map5(&tmx, x, d.C, d.W, d.H, d.T, d.B, p.Cb, p.Wi, p.Hi) != PB_OK) return false;
    if (make_map5(&tmd, dy, d.C, d.Wo, d.Ho, d.To, d.B, p.Cb, p.Wt, p.Ht) != PB_OK) return false;
    static unsigned long long once = 0;
    set_smem_once(dw_wgrad_tma_kernel<K, S, WS>, once);
    (void)launch_pdl(dw_wgrad_tma_kernel<K, S, WS>, dim3(persistent_grid(p)), dim3(DWT_THREADS), (size_t)p.stages * p.stage_bytes + 128, st,
                       tmx, tmd, dw_tc, p);
    return true;
}

// (kT,3,3), temporal stride 1: x [B][T][H][W][C] (+ optional stream buffer [B][kT-1][H][W][C]) -> y [B][To][Ho][Wo][C]
// with To = T + 2*pT - kT + 1 (non-causal) or To = T (streaming: padT = kT - 1 frames of history).
template <int KT, int S, int WS>
static bool launch_fwd3d(const __nv_bfloat16* x, const __nv_bfloat16* sbuf, const float* w_tc, __nv_bfloat16* y,
                         const DwDims& d, int padT, cudaStream_t st) {
    DwTile p{};
    constexpr int TAPS_BYTES = KT * 9 * 128 * 4;                            // fp32 [tap][Cb <= 128]
    PlanIn in{d.B, d.C, d.To, d.Ho, d.Wo, 3, S, WS, d.pH, 0, TAPS_BYTES};
    in.frames = KT;
    if (!plan_tile(in, p)) return false;
    p.nzf = 0; p.f_first = 0; p.f_step = 1; p.f_count = d.To; p.src_first = 0; p.src_step = 1;
    p.ntiles = (long long)p.B * p.f_count * p.tiles_h * p.tiles_w;
    p.kT = KT; p.padT = padT; p.src_T = d.T; p.sbuf_T = sbuf ? KT - 1 : 0;
    p.flip = 0;
    CUtensorMap tm, ts;
    if (make_map5(&tm, x, d.C, d.W, d.H, d.T, d.B, p.Cb, p.Wi, p.Hi) != PB_OK) return false;
    ts = tm;
    if (sbuf && make_map5(&ts, sbuf, d.C, d.W, d.H, KT - 1, d.B, p.Cb, p.Wi, p.Hi) != PB_OK) return false;
    static unsigned long long once = 0;
    set_smem_once(dw_fwd3d_tma_kernel<KT, S, WS>, once);
    (void)launch_pdl(dw_fwd3d_tma_kernel<KT, S, WS>, dim3(persistent_grid(p)), dim3(DWT_THREADS),
                     (size_t)p.stages * p.stage_bytes + 128 + TAPS_BYTES, st, tm, ts, w_tc, y, p);
    return true;
}

static bool movinet3d_class(const DwDims& d) {
    return (d.kT == 3 || d.kT == 5) && d.kH == 3 && d.kW == 3 && d.sT == 1 && d.sH == d.sW && (d.sH == 1 || d.sH == 2) &&
           d.pH == 1 && d.pW == 1 && d.C % 8 == 0;
}

template <int KT>
static bool dispatch_fwd3d(const __nv_bfloat16* x, const __nv_bfloat16* sbuf, const float* w_tc, __nv_bfloat16* y,
                           const DwDims& d, int padT, cudaStream_t st) {
    const bool s7 = d.Wo % 7 == 0;
    if (d.sH == 1) return s7 ? launch_fwd3d<KT, 1, 7>(x, sbuf, w_tc, y, d, padT, st) : launch_fwd3d<KT, 1, 4>(x, sbuf, w_tc, y, d, padT, st);
    return s7 ? launch_fwd3d<KT, 2, 7>(x, sbuf, w_tc, y, d, padT, st) : launch_fwd3d<KT, 2, 4>(x, sbuf, w_tc, y, d, padT, st);
}

// causal streaming forward (pb_stream_dwconv3d_fwd): d.pT == kT - 1 frames of history from `sbuf`
bool dw_stream_fwd_tiled(const __nv_bfloat16* x, const __nv_bfloat16* sbuf, const float* w_tc, __nv_bfloat16* y,
                         const DwDims& d, cudaStream_t st) {
    if (!movinet3d_class(d) || !sbuf) return false;
    if (((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(sbuf)) & 15) != 0) return false;
    return d.kT == 3 ? dispatch_fwd3d<3>(x, sbuf, w_tc, y, d, 2, st) : dispatch_fwd3d<5>(x, sbuf, w_tc, y, d, 4, st);
}

static bool mobilenet_class(const DwDims& d) {
    return d.kT == 1 && d.kH == d.kW && (d.kH == 3 || d.kH == 5) && d.sH == d.sW && (d.sH == 1 || d.sH == 2) &&
           d.pH == d.pW && d.pH == d.kH / 2 && d.C % 8 == 0;
}
static bool aligned16(const void* a, const void* b) {
    return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0;
}
// strips of 7 fit the 112/56/28/14/7 widths of the 224x224 models exactly; 4 otherwise
static bool strip7(int wo) { return wo % 7 == 0; }

// pool != nullptr: also accumulate the global average pool of y into pool[B][C] (zeroed by the caller); only the
// strip kernels do that, so the (kT,3,3) and tensor-core kernels decline and the caller pools separately.
template <> bool dw_fwd_tiled<__nv_bfloat16>(const __nv_bfloat16* x, const float* w_tc, __nv_bfloat16* y,
                                            const DwDims& d, cudaStream_t st, float* pool) {
    if (movinet3d_class(d) && aligned16(x, y) && d.pT == d.kT / 2) {    // MoViNet's symmetric temporal padding
        if (pool) return false;
        return d.kT == 3 ? dispatch_fwd3d<3>(x, nullptr, w_tc, y, d, 1, st) : dispatch_fwd3d<5>(x, nullptr, w_tc, y, d, 2, st);
    }
    if (!mobilenet_class(d) || !aligned16(x, y)) return false;
    if (!pool && dw_fwd_mma(x, w_tc, y, d, st)) return true;   // stride 1: tensor-core kernel (dwconv_mma.cu)
    if (d.kH == 3 && d.sH == 1)
        return strip7(d.Wo) ? launch_fwd<3, 1, 7>(x, w_tc, y, d, d.pT, d.sT, 0, st, pool) : launch_fwd<3, 1, 4>(x, w_tc, y, d, d.pT, d.sT, 0, st, pool);
    if (d.kH == 3 && d.sH == 2)
        return strip7(d.Wo) ? launch_fwd<3, 2, 7>(x, w_tc, y, d, d.pT, d.sT, 0, st, pool) : launch_fwd<3, 2, 4>(x, w_tc, y, d, d.pT, d.sT, 0, st, pool);
    if (d.kH == 5 && d.sH == 1)
        return strip7(d.Wo) ? launch_fwd<5, 1, 7>(x, w_tc, y, d, d.pT, d.sT, 0, st, pool) : launch_fwd<5, 1, 4>(x, w_tc, y, d, d.pT, d.sT, 0, st, pool);
    if (d.kH == 5 && d.sH == 2) return launch_fwd<5, 2, 4>(x, w_tc, y, d, d.pT, d.sT, 0, st, pool);
    return false;
}

// stride-1 input gradient == forward correlation of dy with the flipped filter:
//   dx[t][h][w] = sum_{i,j} dy[t + pT][h + p - i][w + p - j] * w[i][j]
template <> bool dw_dgrad_tiled<__nv_bfloat16>(const __nv_bfloat16* dy, const float* w_tc, __nv_bfloat16* dx,
                                              const DwDims& d, cudaStream_t st) {
    if (!mobilenet_class(d) || !aligned16(dy, dx)) return false;
    if (d.sH == 2) {
        if ((d.H | d.W) < 2) return false;
        return d.kH == 3 ? launch_dgrad_s2<3>(dy, w_tc, dx, d, st) : launch_dgrad_s2<5>(dy, w_tc, dx, d, st);
    }
    if (d.sT != 1) return false;
    if (dw_dgrad_mma(dy, w_tc, dx, d, st)) return true;        // tensor-core kernel (dwconv_mma.cu)
    DwDims r = d;                      // roles swapped: "input" = dy (To,Ho,Wo), "output" = dx (T,H,W)
    r.T = d.To; r.H = d.Ho; r.W = d.Wo;
    r.To = d.T; r.Ho = d.H; r.Wo = d.W;
    if (d.kH == 3)
        return strip7(r.Wo) ? launch_fwd<3, 1, 7>(dy, w_tc, dx, r, -d.pT, 1, 1, st) : launch_fwd<3, 1, 4>(dy, w_tc, dx, r, -d.pT, 1, 1, st);
    return strip7(r.Wo) ? launch_fwd<5, 1, 7>(dy, w_tc, dx, r, -d.pT, 1, 1, st) : launch_fwd<5, 1, 4>(dy, w_tc, dx, r, -d.pT, 1, 1, st);
}

template <> bool dw_wgrad_tiled<__nv_bfloat16>(const __nv_bfloat16* x, const __nv_bfloat16* dy, float* dw_tc,
                                              const DwDims& d, cudaStream_t st) {
    if (!mobilenet_class(d) || !aligned16(x, dy)) return false;
    if (d.kH == 3 && d.sH == 1)
        return strip7(d.Wo) ? launch_wgrad<3, 1, 7>(x, dy, dw_tc, d, st) : launch_wgrad<3, 1, 4>(x, dy, dw_tc, d, st);
    if (d.kH == 3 && d.sH == 2)
        return strip7(d.Wo) ? launch_wgrad<3, 2, 7>(x, dy, dw_tc, d, st) : launch_wgrad<3, 2, 4>(x, dy, dw_tc, d, st);
    if (d.kH == 5 && d.sH == 1)
        return strip7(d.Wo) ? launch_wgrad<5, 1, 7>(x, dy, dw_tc, d, st) : launch_wgrad<5, 1, 4>(x, dy, dw_tc, d, st);
    if (d.kH == 5 && d.sH == 2) return launch_wgrad<5, 2, 4>(x, dy, dw_tc, d, st);
    return false;
}

template <> bool dw_fwd_tiled<float>(const float*, const float*, float*, const DwDims&, cudaStream_t, float*) { return false; }
template <> bool dw_dgrad_tiled<float>(const float*, const float*, float*, const DwDims&, cudaStream_t) { return false; }
template <> bool dw_wgrad_tiled<float>(const float*, const float*, float*, const DwDims&, cudaStream_t) { return false; }

}  // namespace pb
