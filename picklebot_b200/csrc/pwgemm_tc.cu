// Pointwise-conv GEMMs on the 5th-generation tensor cores: TMA -> 128B-swizzled shared memory ->
// tcgen05.mma (one elected thread) -> fp32 accumulators in TMEM -> tcgen05.ld epilogue -> bf16, staged per warp through swizzled shared memory so
// that the global stores are row-contiguous.
//
//   C[b][r][n] = (sum_k A[b][r][k] * W[bw][n][k] + bias[n]) * colscale[b][n] + coladd[b][n]
//
// Replaces nn.Conv3d(kernel_size=1) forward and input-gradient (mobilenet.py:64,79; movinet.py:47,63) for
// bf16 activations.  These layers are HBM-bound (4-70 MAC/B): the design goal is to
// stream A once at full bandwidth, so the kernel is persistent (one CTA per SM), keeps a multi-stage TMA
// ring in flight and double-buffers the accumulator in TMEM so the epilogue of tile i overlaps the MMAs
// of tile i+1.  Skinny layers (K <= 64 and N <= 128: 16->16, 16->64, 24->72 ...) move only a few KB per
// 128-row tile, so the fixed per-tile handshakes would dominate; there a tile spans `mt` (up to 4) stacked
// 128-row sub-tiles that share one pipeline step and one accumulator hand-over.  Tiles never straddle
// samples (3-D tensor maps), which lets the squeeze-excite gate be folded into per-sample weights
// (Bw == Bt) and makes per-sample epilogue vectors trivial.
//
// Warp roles (384 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator,
// warps 4-7 and 8-11 = two epilogue groups, one per accumulator stage (warp w owns TMEM lanes 32*(w%4)..+31).
#include <algorithm>
#include <cstdlib>
#include <mutex>

#include "tc_common.cuh"

namespace pb {
namespace tc {

static EncodeTiledFn g_encode = nullptr;
static std::once_flag g_encode_once;

EncodeTiledFn encode_fn() {
    std::call_once(g_encode_once, [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    });
    return g_encode;
}

constexpr int BM = 128;
constexpr int MAX_STAGES = 8;
constexpr int TMEM_COLS = 512;
constexpr int EPI_STAGE_BYTES = 32 * 128;     // per epilogue warp: 32 rows x one 64-column bf16 panel
constexpr int RING_BYTES = 190 * 1024;         // TMA ring budget (+ 8 staging panels + alignment <= 226 KB)

struct GemmParams {
    int Bt, K, N, Bw;
    long long R;
    int block_n, n_tiles, m_tiles, k_chunks, stages, mt;
    int n_res;         // N > 256 with shared weights: a CTA keeps ONE n-tile for its lifetime (weights resident) and strides over row tiles
    int w_resident;    // the whole weight matrix stays in shared memory for the CTA's lifetime (see pb_pw_gemm_tc)
    int bk;            // K elements per chunk = swizzle span / 2: 64 (128B), 32 (64B) or 16 (32B rows) for tiny K
    long long total_tiles;
    const float* bias;
    const float* colscale;
    const float* coladd;
    const float* tinit;  // additive vector applied by PRE-LOADING the accumulator (see the epilogue warps):
    int tinit_bstride;   //   tinit[b * tinit_bstride + n]; stride N = per sample, 0 = shared (a bias)
    int act;             // activation applied to the fp32 accumulators by the plain epilogue (PB_ACT_*)
    float slope;
    __nv_bfloat16* C;
    double* stats;     // optional BatchNorm statistics of C: [PB_STAT_REPLICAS][2][stat_mod] sums of x and x^2
    int stat_mod;      // real channel count (column c of a row-folded problem is channel c % stat_mod)
};

// t = blockIdx.x + i * gridDim.x is the CTA's i-th work item.  Default: item t is tile (t % n_tiles, t / n_tiles).
// n_res: the CTA owns n-tile blockIdx.x % n_tiles and walks the row tiles with the stride of its group, so the
// 57-92 KB weight tile of the 112->672 / 160->960 layers is fetched once per CTA instead of once per row tile
// (it was two thirds of the bytes TMA moved for those layers).
__device__ __forceinline__ bool tile_coords(const GemmParams& p, long long t, int& n_tile, long long& mtile) {
    if (!p.n_res) {
        if (t >= p.total_tiles) return false;
        n_tile = (int)(t % p.n_tiles);
        mtile = t / p.n_tiles;
        return true;
    }
    const long long i = t / gridDim.x;
    n_tile = (int)(blockIdx.x % p.n_tiles);
    const int rank = (int)(blockIdx.x / p.n_tiles);
    const int G = ((int)gridDim.x - n_tile - 1) / p.n_tiles + 1;
    mtile = rank + i * G;
    return mtile < (long long)p.Bt * p.m_tiles;
}

// EPI: any of bias / colscale / coladd present; STATS: column sums of C; ACT: activation in the plain epilogue
// (a template flag so that the training kernels keep their register allocation: with a run-time test the plain
// kernel lost 12 % on the whole step)
template <bool EPI, bool STATS, bool ACT = false>
__global__ void __launch_bounds__(384, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_bar[MAX_STAGES], empty_bar[MAX_STAGES], tfull_bar[2], tempty_bar[2], wres_bar;
    __shared__ __align__(16) float tin_s[2][256];      // per epilogue group: the additive vector of its next tile (tinit)
    __shared__ uint32_t tmem_base_s;
    pdl_trigger();

    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* tiles = smem_raw + (((raw + 1023u) & ~1023u) - raw);      // 1024-byte aligned (SWIZZLE_128B)
    const int BK = p.bk;
    const int pitch = BK * 2;                                           // bytes per smem row = swizzle span
    const int A_SUB_BYTES = BM * pitch;                                 // one 128-row sub-tile
    const uint32_t layout = pitch == 128 ? 2u : pitch == 64 ? 4u : 6u;  // UMMA layout type of the swizzle mode
    const int a_stage_bytes = p.mt * A_SUB_BYTES;
    const int w_bytes = p.block_n * pitch;                              // one K-chunk of the weight tile
    const int stage_bytes = a_stage_bytes + (p.w_resident ? 0 : w_bytes);
    uint8_t* wres = tiles + (size_t)p.stages * stage_bytes;             // resident weights: [k_chunks][block_n][pitch]
    const int wres_bytes = p.w_resident ? p.k_chunks * w_bytes : 0;
    const int acc_cols = p.mt * p.block_n;                              // TMEM columns per accumulator stage
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmW);
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 128); }
        mbar_init(&wres_bar, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(&tmem_base_s, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    pdl_wait();      // everything above overlapped the previous kernel's tail

    if (warp == 0) {
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            if (p.w_resident) {                 // one weight tile for every row tile: fetch it once
                mbar_expect_tx(&wres_bar, (uint32_t)wres_bytes);
                const int n0 = p.n_res ? (int)(blockIdx.x % p.n_tiles) * p.block_n : 0;
                for (int kc = 0; kc < p.k_chunks; ++kc)
                    tma_load_3d(wres + (size_t)kc * w_bytes, &tmW, &wres_bar, kc * BK, n0, 0);
            }
            int n_tile; long long mtile;
            for (long long t = blockIdx.x; tile_coords(p, t, n_tile, mtile); t += gridDim.x) {
                const int b = (int)(mtile / p.m_tiles), m_tile = (int)(mtile % p.m_tiles);
                for (int kc = 0; kc < p.k_chunks; ++kc) {
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    mbar_expect_tx(&full_bar[s], (uint32_t)stage_bytes);
                    uint8_t* st = tiles + (size_t)s * stage_bytes;
                    for (int sub = 0; sub < p.mt; ++sub)
                        tma_load_3d(st + sub * A_SUB_BYTES, &tmA, &full_bar[s], kc * BK, (m_tile * p.mt + sub) * BM, b);
                    if (!p.w_resident)
                        tma_load_3d(st + a_stage_bytes, &tmW, &full_bar[s], kc * BK, n_tile * p.block_n, p.Bw == 1 ? 0 : b);
                    if (++s == p.stages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc(BM, p.block_n, 0, 0);
            int s = 0; uint32_t ph = 0;
            long long it = 0;
            if (p.w_resident) {
                mbar_wait(&wres_bar, 0);
                tc_fence_after();
            }
            int n_tile_u; long long mtile_u;
            for (long long t = blockIdx.x; tile_coords(p, t, n_tile_u, mtile_u); t += gridDim.x, ++it) {
                const int a = (int)(it & 1);
                const uint32_t aph = (uint32_t)((it >> 1) & 1);
                // tinit: the epilogue group pre-loads the accumulator and arrives once more up front, so the n-th use
                // of a stage waits for its n-th completion instead of finding the first one free
                mbar_wait(&tempty_bar[a], p.tinit ? aph : aph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(a * acc_cols);
                for (int kc = 0; kc < p.k_chunks; ++kc) {
                    mbar_wait(&full_bar[s], ph);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(tiles + (size_t)s * stage_bytes);
                    const uint64_t bdesc = make_desc(p.w_resident ? smem_u32(wres) + (uint32_t)(kc * w_bytes) : sa + a_stage_bytes,
                                                     16, 8 * pitch, layout);
                    for (int sub = 0; sub < p.mt; ++sub) {
                        const uint64_t adesc = make_desc(sa + sub * A_SUB_BYTES, 16, 8 * pitch, layout);
                        for (int k = 0; k < BK / 16; ++k)  // UMMA_K = 16 bf16 = 32 bytes = 2 descriptor units
                            umma_bf16(d_tmem + (uint32_t)(sub * p.block_n), adesc + (uint64_t)(2 * k),
                                      bdesc + (uint64_t)(2 * k), idesc, ((kc | k) != 0) || p.tinit != nullptr);
                    }
                    umma_commit(&empty_bar[s]);             // frees the smem slot when these MMAs retire
                    if (++s == p.stages) { s = 0; ph ^= 1; }
                }
                umma_commit(&tfull_bar[a]);                 // accumulator complete -> epilogue
            }
        }
    } else if (warp >= 4) {
        // two epilogue warpgroups (warps 4-7 and 8-11): group g drains accumulator stage g, i.e. every other
        // tile, so the TMEM loads / conversions / stores of consecutive tiles overlap
        const int q = warp & 3;
        const int group = (warp - 4) >> 2;
        uint8_t* stage_w = wres + wres_bytes + (size_t)(warp - 4) * EPI_STAGE_BYTES;
        // STATS: after the staging transpose a lane owns 8 fixed columns of each 64-column panel, so the
        // BatchNorm sums of the rounded outputs accumulate in registers over every tile of this CTA
        float st_sum[STATS ? 4 : 1][8], st_sq[STATS ? 4 : 1][8];
#pragma unroll
        for (int pi = 0; pi < (STATS ? 4 : 1); ++pi)
#pragma unroll
            for (int i = 0; i < 8; ++i) { st_sum[pi][i] = 0.f; st_sq[pi][i] = 0.f; }
        // tinit: the additive vector of a tile is written into every row of its accumulator stage (each warp its 32 TMEM
        // lanes) BEFORE the MMAs, which then accumulate on top of it, so the plain bf16 epilogue serves the call.  Two
        // steps: `tinit_fetch` copies the tile's <= 256 values into a shared row of this epilogue group -- issued early,
        // its global-load latency hides behind the drain of the current tile -- and `tinit_store` broadcasts them into
        // TMEM (LDS + tcgen05.st).  (Reading the values straight from global memory in the store loop serialised 14-16
        // L2 round trips per tile: 166 us instead of 92 us on the 112 -> 672 squeeze-excite layer.)
        auto tinit_fetch = [&](long long t) {
            int n_tile; long long mt_;
            (void)tile_coords(p, t, n_tile, mt_);
            const int b = (int)(mt_ / p.m_tiles);
            const float* src = p.tinit + (long long)b * p.tinit_bstride + n_tile * p.block_n;
            const int ncols = min(p.block_n, p.N - n_tile * p.block_n);
            const int idx = q * 32 + lane;
            const float v0 = idx < ncols ? __ldg(src + idx) : 0.f;
            const float v1 = idx + 128 < ncols ? __ldg(src + idx + 128) : 0.f;
            asm volatile("bar.sync %0, 128;" ::"r"(2 + group) : "memory");      // the previous tile's stores have read the row
            tin_s[group][idx] = v0;
            tin_s[group][idx + 128] = v1;
        };
        auto tinit_store = [&](int a) {
            asm volatile("bar.sync %0, 128;" ::"r"(2 + group) : "memory");      // the row is complete
            for (int sub = 0; sub < p.mt; ++sub) {
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * acc_cols + sub * p.block_n);
                for (int c0 = 0; c0 < p.block_n; c0 += 16) {                     // block_n is a multiple of 16
                    uint32_t v[16];
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        const float4 f = *reinterpret_cast<const float4*>(&tin_s[group][c0 + j]);
                        v[j] = __float_as_uint(f.x); v[j + 1] = __float_as_uint(f.y);
                        v[j + 2] = __float_as_uint(f.z); v[j + 3] = __float_as_uint(f.w);
                    }
                    tmem_st16(taddr + (uint32_t)c0, v);
                }
            }
            tmem_st_wait();
        };
        if (!STATS && !EPI && p.tinit) {                     // (the pre-load only exists for the plain epilogue)
            const long long t0 = (long long)blockIdx.x + (long long)group * gridDim.x;
            int nt_; long long mt_;
            if (tile_coords(p, t0, nt_, mt_)) {
                tinit_fetch(t0);
                tinit_store(group);
                tc_fence_before();
                mbar_arrive(&tempty_bar[group]);
            }
        }
        long long it = 0;
        int n_tile; long long mtile;
        for (long long t = blockIdx.x; tile_coords(p, t, n_tile, mtile); t += gridDim.x, ++it) {
            const int a = (int)(it & 1);
            if (a != group) continue;
            const uint32_t aph = (uint32_t)((it >> 1) & 1);
            const int b = (int)(mtile / p.m_tiles), m_tile = (int)(mtile % p.m_tiles);
            const int n_base = n_tile * p.block_n;
            const float* cs = p.colscale ? p.colscale + (long long)b * p.N : nullptr;
            const float* ca = p.coladd ? p.coladd + (long long)b * p.N : nullptr;
            mbar_wait(&tfull_bar[a], aph);
            tc_fence_after();
            bool tinit_next = false;
            if (!STATS && !EPI && p.tinit) {
                int nt_; long long mt_;
                tinit_next = tile_coords(p, t + 2LL * gridDim.x, nt_, mt_);
                if (tinit_next) tinit_fetch(t + 2LL * gridDim.x);
            }
            for (int sub = 0; sub < p.mt; ++sub) {
                const long long row0 = ((long long)m_tile * p.mt + sub) * BM + q * 32;     // first row of this warp
                __nv_bfloat16* cbase = p.C + ((long long)b * p.R + row0) * p.N;
                const int rows_ok = (int)max(0LL, min(32LL, p.R - row0));
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * acc_cols + sub * p.block_n);
                if (EPI) {
                    // Epilogue vectors present: 32-column fp32 panels.  After the transpose through the staging
                    // rows a lane owns 4 fixed columns, so bias / colscale / coladd are 3 vector loads per panel
                    // instead of 12 per row, and the arithmetic still runs on the unrounded fp32 accumulators.
                    for (int c0 = 0; c0 < p.block_n; c0 += 32) {
                        const int pw = min(32, p.block_n - c0);                             // 16 or 32
                        uint32_t r[2][16];
#pragma unroll
                        for (int g = 0; g < 2; ++g)
                            if (g * 16 < pw) tmem_ld16(taddr + (uint32_t)(c0 + g * 16), r[g]);
                        tmem_ld_wait();
                        uint8_t* rowp = stage_w + lane * 128;
#pragma unroll
                        for (int g = 0; g < 2; ++g)
                            if (g * 16 < pw) {
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                    *reinterpret_cast<uint4*>(rowp + (((4 * g + j) ^ (lane & 7)) << 4)) =
                                        make_uint4(r[g][4 * j], r[g][4 * j + 1], r[g][4 * j + 2], r[g][4 * j + 3]);
                            }
                        __syncwarp();
                        const int sh = pw == 32 ? 3 : 2;                                   // 8 or 4 chunks per row
                        const int ch = lane & ((1 << sh) - 1), r0 = lane >> sh, rstep = 32 >> sh;
                        const int col = n_base + c0 + ch * 4;
                        if (col < p.N) {
                            float4 vb = make_float4(0.f, 0.f, 0.f, 0.f), vs = make_float4(1.f, 1.f, 1.f, 1.f), va = vb;
                            if (p.bias) vb = __ldg(reinterpret_cast<const float4*>(p.bias + col));
                            if (cs) vs = __ldg(reinterpret_cast<const float4*>(cs + col));
                            if (ca) va = __ldg(reinterpret_cast<const float4*>(ca + col));
                            for (int rr = r0; rr < rows_ok; rr += rstep) {
                                const float4 x = *reinterpret_cast<const float4*>(stage_w + rr * 128 + ((ch ^ (rr & 7)) << 4));
                                uint2 o;
                                o.x = pack_bf16x2(fmaf(x.x + vb.x, vs.x, va.x), fmaf(x.y + vb.y, vs.y, va.y));
                                o.y = pack_bf16x2(fmaf(x.z + vb.z, vs.z, va.z), fmaf(x.w + vb.w, vs.w, va.w));
                                *reinterpret_cast<uint2*>(cbase + (long long)rr * p.N + col) = o;
                            }
                        }
                        __syncwarp();
                    }
                    continue;
                }
                if (STATS) {
#pragma unroll
                    for (int pi = 0; pi < 4; ++pi) {
                        const int c0 = pi * 64;
                        if (c0 < p.block_n) {
                            const int pw = min(64, p.block_n - c0);                                 // 16, 32, 48 or 64
                            uint32_t r[4][16];
#pragma unroll
                            for (int g = 0; g < 4; ++g)
                                if (g * 16 < pw) tmem_ld16(taddr + (uint32_t)(c0 + g * 16), r[g]);
                            tmem_ld_wait();
#pragma unroll
                            for (int g = 0; g < 4; ++g) {
                                const int n0 = n_base + c0 + g * 16;
                                if (g * 16 < pw && n0 < p.N) {
                                    float v[16];
#pragma unroll
                                    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[g][j]);
                                    uint4 o0, o1;
                                    o0.x = pack_bf16x2(v[0], v[1]);   o0.y = pack_bf16x2(v[2], v[3]);
                                    o0.z = pack_bf16x2(v[4], v[5]);   o0.w = pack_bf16x2(v[6], v[7]);
                                    o1.x = pack_bf16x2(v[8], v[9]);   o1.y = pack_bf16x2(v[10], v[11]);
                                    o1.z = pack_bf16x2(v[12], v[13]); o1.w = pack_bf16x2(v[14], v[15]);
                                    uint8_t* rowp = stage_w + lane * 128;
                                    *reinterpret_cast<uint4*>(rowp + (((2 * g) ^ (lane & 7)) << 4)) = o0;
                                    *reinterpret_cast<uint4*>(rowp + (((2 * g + 1) ^ (lane & 7)) << 4)) = o1;
                                }
                            }
                            __syncwarp();
                            // fixed column slot per lane: cprs lanes per row (2, 4 or 8), pw/8 of them active
                            const int sh = pw == 16 ? 1 : pw == 32 ? 2 : 3;
                            const int ch = lane & ((1 << sh) - 1), rstep = 32 >> sh;
                            const int col = n_base + c0 + ch * 8;
                            if (ch * 8 < pw && col < p.N) {
                                for (int rr = lane >> sh; rr < rows_ok; rr += rstep) {
                                    const uint4 v4 = *reinterpret_cast<const uint4*>(stage_w + rr * 128 + ((ch ^ (rr & 7)) << 4));
                                    *reinterpret_cast<uint4*>(cbase + (long long)rr * p.N + col) = v4;
                                    const uint32_t u[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
                                    for (int i = 0; i < 4; ++i) {
                                        const float lo = bf16_lo(u[i]), hi = bf16_hi(u[i]);
                                        st_sum[pi][2 * i] += lo;     st_sq[pi][2 * i] = fmaf(lo, lo, st_sq[pi][2 * i]);
                                        st_sum[pi][2 * i + 1] += hi; st_sq[pi][2 * i + 1] = fmaf(hi, hi, st_sq[pi][2 * i + 1]);
                                    }
                                }
                            }
                            __syncwarp();
                        }
                    }
                    continue;
                }
                // 64-column panels: TMEM -> registers -> this warp's swizzled staging rows -> coalesced stores
                // (a lane owns a row in TMEM; storing rows directly would touch 32 lines per instruction)
                for (int c0 = 0; c0 < p.block_n; c0 += 64) {
                    const int pw = min(64, p.block_n - c0);                                 // 16, 32, 48 or 64
                    uint32_t r[4][16];
#pragma unroll
                    for (int g = 0; g < 4; ++g)
                        if (g * 16 < pw) tmem_ld16(taddr + (uint32_t)(c0 + g * 16), r[g]);
                    tmem_ld_wait();
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const int n0 = n_base + c0 + g * 16;
                        if (g * 16 < pw && n0 < p.N) {
                            float v[16];
#pragma unroll
                            for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[g][j]);
                            if (ACT) act_fwd_vec(v, p.act, p.slope);
                            uint4 o0, o1;
                            o0.x = pack_bf16x2(v[0], v[1]);   o0.y = pack_bf16x2(v[2], v[3]);
                            o0.z = pack_bf16x2(v[4], v[5]);   o0.w = pack_bf16x2(v[6], v[7]);
                            o1.x = pack_bf16x2(v[8], v[9]);   o1.y = pack_bf16x2(v[10], v[11]);
                            o1.z = pack_bf16x2(v[12], v[13]); o1.w = pack_bf16x2(v[14], v[15]);
                            uint8_t* rowp = stage_w + lane * 128;
                            *reinterpret_cast<uint4*>(rowp + (((2 * g) ^ (lane & 7)) << 4)) = o0;
                            *reinterpret_cast<uint4*>(rowp + (((2 * g + 1) ^ (lane & 7)) << 4)) = o1;
                        }
                    }
                    __syncwarp();
                    if (pw == 64) {                                    // 8 chunks per row: 4 rows per instruction
                        const int ch = lane & 7, r4 = lane >> 3;
                        const int col = n_base + c0 + ch * 8;
                        uint4 v4[8];
#pragma unroll
                        for (int s8 = 0; s8 < 8; ++s8) {
                            const int rr = s8 * 4 + r4;
                            v4[s8] = *reinterpret_cast<const uint4*>(stage_w + rr * 128 + ((ch ^ (rr & 7)) << 4));
                        }
                        if (col < p.N) {
#pragma unroll
                            for (int s8 = 0; s8 < 8; ++s8) {
                                const int rr = s8 * 4 + r4;
                                if (rr < rows_ok) *reinterpret_cast<uint4*>(cbase + (long long)rr * p.N + col) = v4[s8];
                            }
                        }
                    } else {
                        const int cpr = pw >> 3;                       // 2, 4 or 6 sixteen-byte chunks per row
                        const uint32_t inv = 65536u / (uint32_t)cpr + 1u;   // exact floor(L / cpr) for L < 256
                        for (int L = lane; L < 32 * cpr; L += 32) {
                            const int rr = (int)(((uint32_t)L * inv) >> 16), ch = L - rr * cpr;
                            const int col = n_base + c0 + ch * 8;
                            if (rr < rows_ok && col < p.N) {
                                const uint4 v4 = *reinterpret_cast<const uint4*>(stage_w + rr * 128 + ((ch ^ (rr & 7)) << 4));
                                *reinterpret_cast<uint4*>(cbase + (long long)rr * p.N + col) = v4;
                            }
                        }
                    }
                    __syncwarp();
                }
            }
            if (tinit_next) tinit_store(a);
            tc_fence_before();
            mbar_arrive(&tempty_bar[a]);
        }
        if (STATS) {
            double* dst = p.stats + (size_t)(blockIdx.x % PB_STAT_REPLICAS) * 2 * p.stat_mod;
#pragma unroll
            for (int pi = 0; pi < 4; ++pi) {
                const int c0 = pi * 64;
                if (c0 < p.block_n) {
                    const int pw = min(64, p.block_n - c0);
                    const int sh = pw == 16 ? 1 : pw == 32 ? 2 : 3;
                    const int ch = lane & ((1 << sh) - 1);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float a0 = st_sum[pi][i], a1 = st_sq[pi][i];
                        for (int off = 16; off >= (1 << sh); off >>= 1) {     // lanes that share the column slot
                            a0 += __shfl_xor_sync(0xffffffffu, a0, off);
                            a1 += __shfl_xor_sync(0xffffffffu, a1, off);
                        }
                        const int col = c0 + ch * 8 + i;
                        if (lane < (1 << sh) && ch * 8 < pw && col < p.N) {
                            atomicAdd(&dst[col % p.stat_mod], (double)a0);
                            atomicAdd(&dst[p.stat_mod + col % p.stat_mod], (double)a1);
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

}  // namespace tc
}  // namespace pb

using namespace pb;
using namespace pb::tc;

extern "C" int pb_pw_gemm_tc_act(const void* A, const void* W_bf16, int Bw, const float* bias, const float* colscale,
                                 const float* coladd, void* C, double* stats, int stat_mod, int Bt, long long R, int K,
                                 int N, int act, float slope, pb_stream_t stream);

extern "C" int pb_pw_gemm_tc(const void* A, const void* W_bf16, int Bw, const float* bias, const float* colscale,
                             const float* coladd, void* C, double* stats, int stat_mod, int Bt, long long R, int K,
                             int N, pb_stream_t stream) {
    return pb_pw_gemm_tc_act(A, W_bf16, Bw, bias, colscale, coladd, C, stats, stat_mod, Bt, R, K, N, PB_ACT_NONE, 0.f, stream);
}

// act != PB_ACT_NONE: C = act(A W^T + bias) -- the inference form of conv -> BatchNorm(eval) -> activation once the
// caller has folded the BatchNorm scale into W and its shift into bias (blocks.py bottleneck_eval).  Only with a
// shared bias or a per-sample coladd alone (both pre-load the accumulator) or no vector at all.
extern "C" int pb_pw_gemm_tc_act(const void* A, const void* W_bf16, int Bw, const float* bias, const float* colscale,
                                 const float* coladd, void* C, double* stats, int stat_mod, int Bt, long long R, int K,
                                 int N, int act, float slope, pb_stream_t stream) {
    PB_REQUIRE(A && W_bf16 && C, "pw_gemm_tc: null pointer");
    PB_REQUIRE(Bt > 0 && R > 0 && K > 0 && N > 0, "pw_gemm_tc: empty problem");
    PB_REQUIRE(K % 8 == 0 && N % 8 == 0, "pw_gemm_tc: K=%d and N=%d must be multiples of 8", K, N);
    PB_REQUIRE(Bw == 1 || Bw == Bt, "pw_gemm_tc: Bw=%d must be 1 or Bt=%d", Bw, Bt);
    PB_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W_bf16) & 15) == 0 &&
               (reinterpret_cast<uintptr_t>(C) & 15) == 0, "pw_gemm_tc: pointers must be 16-byte aligned");
    PB_REQUIRE((!bias || (reinterpret_cast<uintptr_t>(bias) & 15) == 0) &&
               (!colscale || (reinterpret_cast<uintptr_t>(colscale) & 15) == 0) &&
               (!coladd || (reinterpret_cast<uintptr_t>(coladd) & 15) == 0), "pw_gemm_tc: epilogue vectors must be 16-byte aligned");
    PB_REQUIRE((!colscale && !coladd) || N % 4 == 0, "pw_gemm_tc: per-sample epilogue vectors need N %% 4 == 0");
    GemmParams p;
    p.Bt = Bt; p.K = K; p.N = N; p.Bw = Bw; p.R = R;
    p.n_tiles = ceil_div(N, 256);
    p.block_n = (ceil_div(N, p.n_tiles) + 15) / 16 * 16;
    p.bk = K <= 16 ? 16 : (K <= 32 ? 32 : 64);
    const int BK = p.bk;
    p.k_chunks = ceil_div(K, BK);
    // skinny layers: stack up to 4 sub-tiles of 128 rows per pipeline step
    p.mt = 1;
    if (p.n_tiles == 1 && p.k_chunks <= 4) {
        p.mt = std::min(4, 256 / p.block_n);
        p.mt = (int)std::max<long long>(1, std::min<long long>(p.mt, (R + BM - 1) / BM));
        while (p.mt > 1 && RING_BYTES / ((p.mt * BM + p.block_n) * BK * 2) < 3) --p.mt;   // keep >= 3 stages
        // wave quantisation: 458 three-high tiles on 148 CTAs are 4 rounds of 3 (12 units) where 686 two-high tiles
        // are 5 rounds of 2 (10 units) -- the 14x14 layers (100-230 k rows) sit exactly there.  Cost model: rounds x
        // (sub-tiles + a fixed per-tile handshake worth ~0.35 sub-tiles).
        int dev0 = 0, sms0 = 148;
        cudaGetDevice(&dev0);
        cudaDeviceGetAttribute(&sms0, cudaDevAttrMultiProcessorCount, dev0);
        const long long sub_tiles = (R + BM - 1) / BM;
        double best_cost = 1e30;
        int best_mt = p.mt;
        for (int m = p.mt; m >= 1; --m) {
            const long long tiles = (long long)Bt * ((sub_tiles + m - 1) / m);
            const long long rounds = (tiles + sms0 - 1) / sms0;
            const double cost = (double)rounds * (m + 0.35);
            if (cost < best_cost - 1e-9) { best_cost = cost; best_mt = m; }
        }
        if (!getenv("PB_GEMM_NO_WAVE_FIT")) p.mt = best_mt;
    }
    p.m_tiles = ceil_div(R, (long long)BM * p.mt);
    p.total_tiles = (long long)Bt * p.m_tiles * p.n_tiles;
    // One weight tile serves every row tile when N fits one tile and the weights are shared by all samples: keep it
    // resident instead of re-fetching block_n TMA rows per row tile (for 40 -> 240 channels that was two thirds of
    // all box rows the TMA unit processed).
    const int wres_bytes = p.k_chunks * p.block_n * BK * 2;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    p.n_res = (Bw == 1 && p.n_tiles > 1 && wres_bytes <= 96 * 1024 && (long long)Bt * p.m_tiles >= sms &&
               !getenv("PB_GEMM_NO_NRES")) ? 1 : 0;
    p.w_resident = ((Bw == 1 && p.n_tiles == 1 && wres_bytes <= 64 * 1024) || p.n_res) ? 1 : 0;
    const int stage_bytes = p.w_resident ? p.mt * BM * BK * 2 : (p.mt * BM + p.block_n) * BK * 2;
    p.stages = std::min(MAX_STAGES, (RING_BYTES - (p.w_resident ? wres_bytes : 0)) / stage_bytes);
    PB_REQUIRE(p.stages >= 2, "pw_gemm_tc: internal tiling error");
    // A per-sample additive vector alone does not need the fp32 epilogue: the accumulator is pre-loaded with it
    // (tcgen05.st by the epilogue warps) and the bf16 panel path stores the result.  PB_GEMM_NO_TINIT=1 keeps the
    // old route for A/B tests.
    p.tinit = nullptr; p.tinit_bstride = 0;
    if (!stats && !colscale && !getenv("PB_GEMM_NO_TINIT")) {
        if (coladd && !bias) { p.tinit = coladd; p.tinit_bstride = N; coladd = nullptr; }
        else if (bias && !coladd) { p.tinit = bias; p.tinit_bstride = 0; bias = nullptr; }
    }
    p.act = act; p.slope = slope;
    PB_REQUIRE(act == PB_ACT_NONE || (!bias && !colscale && !coladd && !stats),
               "pw_gemm_tc: an activation needs the plain epilogue (no statistics, at most one additive vector)");
    p.bias = bias; p.colscale = colscale; p.coladd = coladd; p.C = (__nv_bfloat16*)C;
    p.stats = stats; p.stat_mod = stat_mod;
    if (stats) {
        PB_REQUIRE(!bias && !colscale && !coladd && p.n_tiles == 1, "pw_gemm_tc: fused statistics need N <= 256 and no epilogue vectors");
        PB_REQUIRE(stat_mod > 0 && N % stat_mod == 0, "pw_gemm_tc: stat_mod=%d must divide N=%d", stat_mod, N);
        PB_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * PB_STAT_REPLICAS * 2 * stat_mod, (cudaStream_t)stream));
    }
    const size_t smem = (size_t)p.stages * stage_bytes + (p.w_resident ? wres_bytes : 0) + 8 * EPI_STAGE_BYTES + 1024;

    CUtensorMap tmA, tmW;
    {
        uint64_t dims[3] = {(uint64_t)K, (uint64_t)R, (uint64_t)Bt};
        uint64_t str[3] = {2, (uint64_t)K * 2, (uint64_t)R * K * 2};
        uint32_t box[3] = {(uint32_t)BK, BM, 1};
        if (int e = make_tmap_bf16(&tmA, A, 3, dims, str, box, BK * 2)) return e;
    }
    {
        uint64_t dims[3] = {(uint64_t)K, (uint64_t)N, (uint64_t)Bw};
        uint64_t str[3] = {2, (uint64_t)K * 2, (uint64_t)N * K * 2};
        uint32_t box[3] = {(uint32_t)BK, (uint32_t)p.block_n, 1};
        if (int e = make_tmap_bf16(&tmW, W_bf16, 3, dims, str, box, BK * 2)) return e;
    }
    static unsigned long long attr_done[4] = {0, 0, 0, 0};
    cudaError_t attr_err = ensure_dyn_smem(gemm_tc_kernel<false, false>, 224 * 1024, &attr_done[0]);
    if (attr_err == cudaSuccess) attr_err = ensure_dyn_smem(gemm_tc_kernel<false, false, true>, 224 * 1024, &attr_done[3]);
    if (attr_err == cudaSuccess) attr_err = ensure_dyn_smem(gemm_tc_kernel<true, false>, 224 * 1024, &attr_done[1]);
    if (attr_err == cudaSuccess) attr_err = ensure_dyn_smem(gemm_tc_kernel<false, true>, 224 * 1024, &attr_done[2]);
    if (attr_err != cudaSuccess) return cuda_fail(attr_err, "cudaFuncSetAttribute(gemm_tc_kernel)");
    int grid = (int)std::min<long long>(p.total_tiles, sms);
    if (bias || colscale || coladd)
        PB_CUDA(launch_pdl(gemm_tc_kernel<true, false>, dim3(grid), dim3(384), smem, (cudaStream_t)stream, tmA, tmW,
                           p));
    else if (stats)
        PB_CUDA(launch_pdl(gemm_tc_kernel<false, true>, dim3(grid), dim3(384), smem, (cudaStream_t)stream, tmA, tmW,
                           p));
    else if (p.act != PB_ACT_NONE)
        PB_CUDA(launch_pdl(gemm_tc_kernel<false, false, true>, dim3(grid), dim3(384), smem, (cudaStream_t)stream, tmA, tmW,
                           p));
    else
        PB_CUDA(launch_pdl(gemm_tc_kernel<false, false>, dim3(grid), dim3(384), smem, (cudaStream_t)stream, tmA, tmW,
                           p));
    PB_CHECK_LAUNCH("gemm_tc_kernel");
    count_path(PB_PATH_GEMM_TC);
    return PB_OK;
}
