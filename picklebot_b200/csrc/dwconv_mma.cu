// Depthwise (1,k,k) convolution, stride 1, bf16 NDHWC: TMA-staged halo tiles in, warp-level tensor-core FMAs,
// TMA-stored output tiles out.
//
// Why tensor cores for an HBM-bound stencil: the CUDA-core kernels of dwconv_tiled.cu need ~20 (3x3) to ~30 (5x5)
// issued instructions per output element -- bf16->fp32 unpacking, address arithmetic and 25 FMAs per element of a
// 5x5 filter -- which caps them at 2-3.5 TB/s with the issue slots half busy (ncu, profiles/r01_ncu_dw_*); a 5x5
// stride-1 layer at 70 % of the HBM rate would need 77 % of the chip's fp32 FMA peak.  tcgen05.mma cannot take
// this problem: its operands must be canonical shared-memory / TMEM tiles, and a depthwise filter shares no operand
// between channels, so with channels-last data every K-step would be strided by the channel count.  The legacy
// warp-level mma.sync takes its operands from REGISTERS, which lets each lane assemble exactly the fragment it
// needs from the staged tile:
//
//   per channel c and filter row i:   D[m][n] += sum_k A[m][k] * B_i[k][n]            (m16n8k16, bf16 -> fp32)
//     A[m][k]   = tile[row(m) + i][wbase(m) + k][c]        16 contexts (row, 16-pixel window) x 16 input columns
//     B_i[k][n] = w[i][k - n]  (0 <= k-n < K, else 0)      the filter row as a banded (Toeplitz) matrix
//     D[m][n]   = output pixel wbase(m) + n of context m
//
// A window is 16 input columns and yields 16 - (K-1) outputs; output columns 8..15 reuse the same B registers
// shifted by 8 ({0,b0}), so a lane holds 2 weight registers per (channel, i).  A lane owns contexts g and g+8
// (g = lane/4).  Fragments are built from 16-byte shared-memory loads (8 channels of one pixel) with one PRMT per
// register (pixel pair of one channel): ~0.5 ALU instructions per staged element and filter row instead of ~1 per
// tap.  Products are exact and accumulate in fp32, like the CUDA-core kernels.  Instruction budget: ~4 (3x3) /
// ~7 (5x5) per output element.
//
// Shared-memory banks decide the layout (ncu on the first version: 83 % of the shared-memory wavefront peak,
// 7.3 wavefronts per 8-byte load): every lane address is a multiple of the pixel pitch, so 8-byte loads can at best
// reach a 2-way conflict.  16-byte loads are conflict free when, per quarter warp (lanes g in {2k, 2k+1}, t = 0..3),
// the eight 16-byte pieces fall into different bank groups: columns 2t give t*(Cb/4) mod 8 = {0,2,4,6} when Cb/8 is
// odd, and consecutive contexts g, g+1 are consecutive ROWS of one window, one odd multiple of 16 bytes apart when
// the staged width Wi is odd as well.
//
// Output: lanes hold (pixel, 4 channels) fragments, i.e. 8-byte pieces 2*C bytes apart -- written straight to
// global memory that would be one 32-byte sector per lane.  They are staged in a shared-memory output tile
// [NF][Ht][Wt][Cb] instead and one thread issues a cp.async.bulk.tensor store (the TMA unit clips ragged edges),
// double buffered so the store of tile n overlaps the math of tile n+1.
//
// Used for: forward of the stride-1 Bottleneck3D.depthwise_conv layers (mobilenet.py:67-75) and their input
// gradient (same kernel, flipped taps, roles of x and y swapped).
#include <algorithm>
#include <cstdlib>

#include "dwconv.cuh"
#include "tc_common.cuh"

namespace pb {

using namespace tc;

constexpr int DWM_CWARPS = 8;                       // compute warps
constexpr int DWM_COMPUTE = DWM_CWARPS * 32;
constexpr int DWM_MAX_STAGES = 4;
constexpr int DWM_THREADS = DWM_COMPUTE + 32 * DWM_MAX_STAGES;   // + one warp per ring stage that drives the TMA unit
constexpr int DWM_MAX_ZF = 64;
constexpr int DWM_SMEM_BUDGET = 216 * 1024;

struct MmaPlan {
    int B, C;
    int To, Ho, Wo;             // destination tensor
    int p;                      // spatial padding
    int Cb, NCG, nblk;          // channel block, 4-channel groups per block, blocks
    int NW, VW, RP;             // windows per tile row, outputs per window, destination rows per pass
    int Ht, Hi, Wi, NF;         // tile: NF frames x Ht x Wo outputs, staged NF x Hi x Wi (full width)
    int NP;                     // passes per tile = ceil(NF*Ht / RP)
    int tiles_h, tiles_f;
    int f_first, f_count, src_first;   // destination frames with a source frame (contiguous), and the first source frame
    int nzf;
    unsigned char zf[DWM_MAX_ZF];
    long long ntiles;
    int in_bytes, stages, tab_bytes;
    int flip;
};

struct MmaCtx {
    uint64_t full[DWM_MAX_STAGES];      // TMA load of the stage has landed
    uint64_t done[DWM_MAX_STAGES];      // all compute warps have finished the tile (outputs written in place)
};

__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t a0, const uint32_t a1, const uint32_t a2,
                                         const uint32_t a3, const uint32_t b0, const uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// Plain C++ accesses to the dynamic shared-memory array: the compiler is free to schedule them around the MMAs
// (software pipelining of the next row's loads), while the "memory" clobbers of the mbarrier waits and the block
// barrier keep them on the right side of the synchronisation points.
__device__ __forceinline__ uint2 lds64m(const uint8_t* p) { return *reinterpret_cast<const uint2*>(p); }
__device__ __forceinline__ uint4 lds128m(const uint8_t* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ void sts128m(uint8_t* p, uint4 v) { *reinterpret_cast<uint4*>(p) = v; }
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* tmap, const void* smem_src, int c0, int c1, int c2, int c3,
                                             int c4) {
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// A-fragment halves of one tile row for 8 channels: lo = pixels (2t, 2t+1), hi = pixels (2t+8, 2t+9)
struct RowFrag {
    uint32_t lo[8];
    uint32_t hi[8];
};

// `a` = start of this lane's context row, off[j] = byte offsets of its four columns (window columns 2t, 2t+1, 2t+8,
// 2t+9 shifted by the padding).  Columns outside the image read the staged zero column Wo instead (the TMA unit
// filled it): no predicates in the inner loop.
__device__ __forceinline__ void load_row(RowFrag& f, const uint8_t* a, const uint32_t (&off)[4]) {
    const uint4 l0 = lds128m(a + off[0]), l1 = lds128m(a + off[1]), l8 = lds128m(a + off[2]), l9 = lds128m(a + off[3]);
    f.lo[0] = prmt(l0.x, l1.x, 0x5410); f.lo[1] = prmt(l0.x, l1.x, 0x7632);
    f.lo[2] = prmt(l0.y, l1.y, 0x5410); f.lo[3] = prmt(l0.y, l1.y, 0x7632);
    f.lo[4] = prmt(l0.z, l1.z, 0x5410); f.lo[5] = prmt(l0.z, l1.z, 0x7632);
    f.lo[6] = prmt(l0.w, l1.w, 0x5410); f.lo[7] = prmt(l0.w, l1.w, 0x7632);
    f.hi[0] = prmt(l8.x, l9.x, 0x5410); f.hi[1] = prmt(l8.x, l9.x, 0x7632);
    f.hi[2] = prmt(l8.y, l9.y, 0x5410); f.hi[3] = prmt(l8.y, l9.y, 0x7632);
    f.hi[4] = prmt(l8.z, l9.z, 0x5410); f.hi[5] = prmt(l8.z, l9.z, 0x7632);
    f.hi[6] = prmt(l8.w, l9.w, 0x5410); f.hi[7] = prmt(l8.w, l9.w, 0x7632);
}

template <int K>
__global__ void __launch_bounds__(DWM_THREADS, 1)
dw_s1_mma_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY,
                 const float* __restrict__ w_tc, __nv_bfloat16* __restrict__ y, const MmaPlan p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ MmaCtx cx;
    constexpr int NT = 2;                                // 8-wide output column groups per window
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* ring = smem_raw + (((raw + 127u) & ~127u) - raw);
    uint8_t* tab = ring + (size_t)p.stages * p.in_bytes;
    const int tid = threadIdx.x;
    const int c_base = blockIdx.y * p.Cb;
    const uint32_t pb = (uint32_t)p.Cb * 2;              // bytes per staged pixel
    pdl_trigger();
    if (tid == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(&cx.full[s], 1); mbar_init(&cx.done[s], DWM_CWARPS); }
        fence_barrier_init();
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmY);
    }
    pdl_wait();                 // w_tc (a cast kernel's output) and the activations are valid from here on
    // weight table: tab[cg][i][slot][8 ch] = packed (w[i][slot-1], w[i][slot]) as bf16 pairs, zero outside the filter
    {
        uint32_t* t32 = reinterpret_cast<uint32_t*>(tab);
        const int n = p.NCG * K * 8 * 8;
        for (int e = tid; e < n; e += DWM_THREADS) {
            const int ch = e & 7, slot = (e >> 3) & 7, i = (e >> 6) % K, cg = (e >> 6) / K;
            const int c = c_base + cg * 8 + ch;
            float lo = 0.f, hi = 0.f;
            if (c < p.C) {
                const int d0 = slot - 1, d1 = slot;
                if (d0 >= 0 && d0 < K) { const int tap = i * K + d0; lo = w_tc[(long long)(p.flip ? K * K - 1 - tap : tap) * p.C + c]; }
                if (d1 >= 0 && d1 < K) { const int tap = i * K + d1; hi = w_tc[(long long)(p.flip ? K * K - 1 - tap : tap) * p.C + c]; }
            }
            t32[e] = pack_bf16x2(lo, hi);
        }
    }
    __syncthreads();
    const int my_tiles = p.ntiles > blockIdx.x ? (int)((p.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;

    if (tid >= DWM_COMPUTE) {
        // ---- TMA warps, one lane each, one per ring stage: stage s serves tiles s, s + stages, ...  When the compute
        // warps are done with a tile (its outputs now sit in the stage, at the positions of the top-left input
        // pixels) it is stored from there and, once the store has READ the stage, the stage is refilled.  A stage
        // per warp keeps that wait off the other stages' critical path (bulk async-groups are per thread).
        const int s = (tid - DWM_COMPUTE) >> 5;
        if ((tid & 31) == 0 && s < p.stages) {
            auto tile_coord = [&](int n) {
                unsigned t = (unsigned)blockIdx.x + (unsigned)n * gridDim.x;      // ntiles < 2^31 (checked on the host)
                int3 c;
                c.z = (int)(t % p.tiles_h) * p.Ht; t /= p.tiles_h;
                c.y = (int)(t % p.tiles_f);
                c.x = (int)(t / p.tiles_f);
                return c;
            };
            uint8_t* stage = ring + (size_t)s * p.in_bytes;
            auto issue_load = [&](int n) {
                const int3 c = tile_coord(n);
                mbar_expect_tx(&cx.full[s], (uint32_t)(p.NF * p.Hi * p.Wi * p.Cb * 2));
                tma_load_5d(stage, &tmX, &cx.full[s], c_base, 0, c.z - p.p, p.src_first + c.y * p.NF, c.x);
            };
            if (s < my_tiles) issue_load(s);
            uint32_t phase = 0;
            for (int n = s; n < my_tiles; n += p.stages, phase ^= 1) {
                mbar_wait_parked(&cx.done[s], phase);
                const int3 c = tile_coord(n);
                tma_store_5d(&tmY, stage, c_base, 0, c.z, p.f_first + c.y * p.NF, c.x);
                bulk_commit();
                if (n + p.stages < my_tiles) {
                    bulk_wait_read0();                      // the stores have read the stage: it may be overwritten
                    issue_load(n + p.stages);
                }
            }
            bulk_wait0();
        }
        return;
    }

    // destination frames without a source frame are all zero: whole frames, round-robin over the grid
    if (p.nzf > 0) {
        const int frame16 = (int)((long long)p.Ho * p.Wo * p.C / 8);
        const int nframes = p.B * p.nzf;
        const int ncta = gridDim.x * gridDim.y;
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        for (int q = blockIdx.y * gridDim.x + blockIdx.x; q < nframes; q += ncta) {
            const int f = p.zf[q % p.nzf], b = q / p.nzf;
            uint4* dst = reinterpret_cast<uint4*>(y) + ((long long)b * p.To + f) * frame16;
            for (int e = tid; e < frame16; e += DWM_COMPUTE) dst[e] = z;
        }
    }

    const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    // contexts g (h = 0) and g + 8 (h = 1): context m is row m % RP of window m / RP of the pass, so that the lane
    // pairs (g, g+1) of a quarter warp read consecutive rows (bank-conflict free, see the file header)
    int ctx_row[2];
    uint32_t off[2][4];                                  // byte offsets of the four input columns this lane reads
    uint32_t ocol[2];                                    // byte offset of output column 2t of the lane's window
    uint32_t omask[2] = {0, 0};                          // bit (2a + e): output column 8a + 2t + e exists
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int m = g + 8 * h;
        ctx_row[h] = m % p.RP;
        const int win = m / p.RP;
        const int c0 = win * p.VW + 2 * t - p.p;         // first input column (may be left of the image)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c0 + (j & 1) + (j >> 1) * 8;
            off[h][j] = (uint32_t)((c >= 0 && c < p.Wo) ? c : p.Wo) * pb;      // column Wo is staged as zeros
        }
        ocol[h] = (uint32_t)(win * p.VW + 2 * t) * pb;
#pragma unroll
        for (int a = 0; a < NT; ++a)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int wl = 8 * a + 2 * t + e;
                if (wl < p.VW && win * p.VW + wl < p.Wo) omask[h] |= 1u << (2 * a + e);
            }
    }
    // weight-table slots of this lane: b0 = P[e], b1 = P[e+8] with e = 2t - g (slot 7 holds zeros)
    const int e0 = 2 * t - g;
    const uint32_t slot0 = (uint32_t)((e0 >= -1 && e0 <= K - 1) ? e0 + 1 : 7) * 32;
    const uint32_t slot1 = (uint32_t)((e0 + 8 >= -1 && e0 + 8 <= K - 1) ? e0 + 9 : 7) * 32;
    const int nfht = p.NF * p.Ht;
    const uint32_t row_bytes = (uint32_t)p.Wi * pb;

    for (int n = 0; n < my_tiles; ++n) {
        const int s = n % p.stages;
        mbar_wait_parked(&cx.full[s], (uint32_t)((n / p.stages) & 1));
        uint8_t* stage = ring + (size_t)s * p.in_bytes;

        // A warp owns whole channel groups (8 channels = 16 bytes of every staged pixel) and walks their passes top
        // to bottom.  Outputs replace the inputs in place: output (r, c) goes where input (r, c) was.  Nobody else
        // touches this channel group's bytes, a pass has read everything it needs before it writes, and the rows it
        // overwrites (its own destination rows) are not read by the passes below it.
        for (int cg = warp; cg < p.NCG; cg += DWM_CWARPS) {
            const uint8_t* wrow = tab + (uint32_t)(cg * K) * 256;
            for (int pass = 0; pass < p.NP; ++pass) {
                // destination rows (frame-major) of this lane's two contexts, and where they sit in the staged tile
                bool valid[2];
                uint8_t* orow[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int vr = pass * p.RP + ctx_row[h];
                    valid[h] = vr < nfht;
                    const int vrc = valid[h] ? vr : 0;
                    const int f = p.NF == 1 ? 0 : vrc / p.Ht;
                    orow[h] = stage + (uint32_t)(vrc + f * (K - 1)) * row_bytes + (uint32_t)cg * 16;   // frame f starts at row f*Hi
                }

                float acc[NT][8][4];
#pragma unroll
                for (int a = 0; a < NT; ++a)
#pragma unroll
                    for (int c = 0; c < 8; ++c)
#pragma unroll
                        for (int r = 0; r < 4; ++r) acc[a][c][r] = 0.f;

#pragma unroll 1
                for (int i = 0; i < K; ++i) {
                    RowFrag r0, r1;
                    load_row(r0, orow[0] + (uint32_t)i * row_bytes, off[0]);
                    load_row(r1, orow[1] + (uint32_t)i * row_bytes, off[1]);
                    const uint4 b0a = lds128m(wrow + (uint32_t)i * 256 + slot0), b0b = lds128m(wrow + (uint32_t)i * 256 + slot0 + 16);
                    const uint4 b1a = lds128m(wrow + (uint32_t)i * 256 + slot1), b1b = lds128m(wrow + (uint32_t)i * 256 + slot1 + 16);
                    const uint32_t b0c[8] = {b0a.x, b0a.y, b0a.z, b0a.w, b0b.x, b0b.y, b0b.z, b0b.w};
                    const uint32_t b1c[8] = {b1a.x, b1a.y, b1a.z, b1a.w, b1b.x, b1b.y, b1b.z, b1b.w};
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        mma_bf16(acc[0][c], r0.lo[c], r1.lo[c], r0.hi[c], r1.hi[c], b0c[c], b1c[c]);   // outputs 0..7
                        mma_bf16(acc[1][c], r0.lo[c], r1.lo[c], r0.hi[c], r1.hi[c], 0u, b0c[c]);       // outputs 8..15 (k >= 8)
                    }
                }
                __syncwarp();        // every lane has consumed its inputs (the MMAs are warp-wide): writes may start

                // epilogue: (context h, column 8a + 2t + e) x 8 channels -> 16-byte pieces, in place
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (!valid[h]) continue;
                    uint8_t* o0 = orow[h] + ocol[h];
#pragma unroll
                    for (int a = 0; a < NT; ++a) {
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            if ((omask[h] >> (2 * a + e)) & 1u) {
                                uint4 o;
                                o.x = pack_bf16x2(acc[a][0][2 * h + e], acc[a][1][2 * h + e]);
                                o.y = pack_bf16x2(acc[a][2][2 * h + e], acc[a][3][2 * h + e]);
                                o.z = pack_bf16x2(acc[a][4][2 * h + e], acc[a][5][2 * h + e]);
                                o.w = pack_bf16x2(acc[a][6][2 * h + e], acc[a][7][2 * h + e]);
                                sts128m(o0 + (uint32_t)(8 * a + e) * pb, o);
                            }
                        }
                    }
                }
                __syncwarp();        // ... and are complete before the next pass reads rows below
            }
        }
        fence_proxy_async();                                   // in-place outputs -> visible to the TMA unit
        __syncwarp();
        if (lane == 0) mbar_arrive(&cx.done[s]);
    }
}

// ------------------------------------------------------------------------------------------------
// host-side planning
// ------------------------------------------------------------------------------------------------
static int make_map5_mma(CUtensorMap* tm, const void* base, int C, int W, int H, int T, int B, int bc, int bw, int bh, int bt) {
    uint64_t dims[5] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)T, (uint64_t)B};
    uint64_t str[5] = {2, (uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2, (uint64_t)T * H * W * C * 2};
    uint32_t box[5] = {(uint32_t)bc, (uint32_t)bw, (uint32_t)bh, (uint32_t)bt, 1};
    return make_tmap_bf16(tm, base, 5, dims, str, box, 0);
}

// Tile planner.  Constraints that come from the hardware (measured, tools/tma_bench.cu and the bank structure):
//  * shared-memory banks: lanes of a warp read 8-byte pieces of pixels (row g, column 2t + ...).  All strides are
//    multiples of the pixel pitch, so the best case is a 2-way conflict, reached when Cb/8 and the staged width Wi
//    are both odd; Cb = 96 with Wi = 16 is a 32-way conflict (measured: 3x slower than Cb = 120).
//  * TMA unit: a box row is one staged pixel (Cb*2 bytes, rows are 2*C bytes apart in memory); one CTA gets a row
//    every ~5 cycles up to ~140 bytes and ~28 bytes per cycle beyond, zero-filled rows included.
//  * one tile spans the full width (in-place output; columns beyond the image are clipped by the store).
// The score is useful bytes per estimated tile time = max(HBM time, TMA time, issue time).
static bool plan_mma(MmaPlan& p, int B, int C, int To, int Ho, int Wo, int K, int f_count) {
    p.B = B; p.C = C; p.To = To; p.Ho = Ho; p.Wo = Wo; p.p = K / 2;
    p.VW = 16 - (K - 1);
    const int nw = ceil_div(Wo, p.VW);
    if (nw > 8) return false;
    p.NW = nw <= 1 ? 1 : nw <= 2 ? 2 : nw <= 4 ? 4 : 8;
    p.RP = 16 / p.NW;
    p.Wi = Wo + 1 + (Wo & 1);                         // odd staged width with at least one zero column right of the image
    if (p.Wi > 256) return false;
    if (C < 64) return false;                         // box rows under 128 bytes: the CUDA-core kernels are faster
    double best = -1.0;
    int bCb = 0, bHt = 0, bNF = 0, bSt = 0;
    for (int Cb = 72; Cb <= 120; Cb += 16) {          // Cb/8 odd
        const int nblk = ceil_div(C, Cb);
        const int ncg = Cb / 8;
        const int tab = ncg * K * 256;
        const double ch_eff = (double)C / (nblk * Cb);                   // real channels per staged channel
        const double warp_eff = (double)ncg / (ceil_div(ncg, DWM_CWARPS) * DWM_CWARPS);
        for (int NF = 1; NF <= 4; ++NF) {
            if (NF > 1 && NF > f_count) break;
            for (int Ht = 1; Ht <= std::min(Ho, 64); ++Ht) {
                if (NF > 1 && Ht < Ho) continue;                         // several frames per tile only for whole planes
                const int Hi = Ht + K - 1;
                const long long inb = ((long long)NF * Hi * p.Wi * Cb * 2 + 127) / 128 * 128;
                const long long left = DWM_SMEM_BUDGET - tab;
                if (left < 2 * inb) continue;
                const int st = (int)std::min<long long>(DWM_MAX_STAGES, left / inb);
                const int NP = ceil_div(NF * Ht, p.RP);
                const int th = ceil_div(Ho, Ht);
                const double f_eff = (double)f_count / (ceil_div(f_count, NF) * NF);
                const double ctx_eff = ((double)Ho / th * NF * f_eff) / (NP * p.RP) * ((double)Wo / (p.NW * p.VW));
                const double halo_eff = ((double)Ho / th * 2.0) / (Hi + (NF > 1 ? Hi : Ht));   // useful share of the loaded + stored rows
                // the math has headroom over the memory system (ncu: issue slots ~40 % busy at 4 TB/s), so its
                // efficiencies enter damped; a third ring stage hides the store -> refill turnaround
                const double row_pref = 0.7 + 0.3 * std::min(120.0, (double)C / nblk) / 120.0;   // long box rows
                const double score = (0.2 + 0.8 * ch_eff) * (0.4 + 0.6 * ctx_eff * warp_eff) * halo_eff * row_pref *
                                     (st >= 3 ? 1.0 : 0.85);
                if (score > best + 1e-9) { best = score; bCb = Cb; bHt = Ht; bNF = NF; bSt = st; }
            }
        }
    }
    if (best < 0) return false;
    p.Cb = bCb; p.NCG = bCb / 8; p.nblk = ceil_div(C, bCb);
    p.Ht = bHt; p.NF = bNF; p.Hi = bHt + K - 1; p.stages = bSt;
    p.NP = ceil_div(p.NF * p.Ht, p.RP);
    p.tiles_h = ceil_div(Ho, p.Ht);
    p.in_bytes = (int)(((long long)p.NF * p.Hi * p.Wi * p.Cb * 2 + 127) / 128 * 128);
    p.tab_bytes = p.NCG * K * 256;
    return true;
}


// dst[b][to][ho][wo][c] = sum_{i,j} src[b][to - pT][ho + i - p][wo + j - p][c] * w[(flip) i*K + j][c]
// with src of size (Ts, Hs, Ws) == spatially (Ho, Wo); destination frames without a source frame are zeroed.
template <int K>
static bool launch_mma(const __nv_bfloat16* src, const float* w_tc, __nv_bfloat16* dst, int B, int C, int Ts, int To, int Ho,
                       int Wo, int pT, int flip, cudaStream_t st) {
    MmaPlan p{};
    // destination frame `to` reads source frame to - pT
    p.nzf = 0;
    int first = -1, count = 0;
    for (int f = 0; f < To; ++f) {
        const int v = f - pT;
        if (v >= 0 && v < Ts) { if (first < 0) first = f; ++count; }
        else { if (p.nzf >= DWM_MAX_ZF || f > 255) return false; p.zf[p.nzf++] = (unsigned char)f; }
    }
    if (count == 0) return false;
    if (!plan_mma(p, B, C, To, Ho, Wo, K, count)) return false;
    p.f_first = first; p.f_count = count; p.src_first = first - pT;
    p.tiles_f = ceil_div(count, p.NF);
    p.ntiles = (long long)B * p.tiles_f * p.tiles_h;
    if (p.ntiles >= (1LL << 31)) return false;
    p.flip = flip;
    CUtensorMap tmx, tmy;
    if (make_map5_mma(&tmx, src, C, Wo, Ho, Ts, B, p.Cb, p.Wi, p.Hi, p.NF) != PB_OK) return false;
    // The store reads the stage itself: full staged width (columns >= Wo are clipped by the TMA unit) and, with
    // several frames per tile, the staged frame pitch Hi (rows >= Ho clipped: such tiles cover whole planes; a
    // store per frame would need frame offsets that are multiples of 128 bytes).
    if (make_map5_mma(&tmy, dst, C, Wo, Ho, To, B, p.Cb, p.Wi, p.NF > 1 ? p.Hi : p.Ht, p.NF) != PB_OK) return false;
    static unsigned long long once = 0;
    if (ensure_dyn_smem(dw_s1_mma_kernel<K>, 226 * 1024, &once) != cudaSuccess) return false;
    int ctas = std::max(1, 148 / p.nblk);
    ctas = (int)std::min<long long>(ctas, p.ntiles);
    if (p.stages > DWM_MAX_STAGES) p.stages = DWM_MAX_STAGES;
    const size_t smem = (size_t)p.stages * p.in_bytes + p.tab_bytes + 128;
    (void)launch_pdl(dw_s1_mma_kernel<K>, dim3(ctas, p.nblk), dim3(DWM_THREADS), smem, st, tmx, tmy, w_tc, dst, p);
    return true;
}

// PB_DW_MMA=1 forces this kernel for every shape it supports (tests, experiments), =0 disables it; by default it
// serves the shapes where it measured faster than the CUDA-core kernels of dwconv_tiled.cu on B200
// (profiles/r02_dw_microbench.txt): 3x3 layers on 14x14 planes with >= 400 channels (MobileNetLarge3D block4.4 /
// block4.5: 4.1 vs 3.7 TB/s forward, 4.2 vs 3.9 TB/s input gradient).
static int dw_mma_mode() {           // read per call: tests switch it with os.environ
    const char* e = getenv("PB_DW_MMA");
    return !e || !e[0] ? -1 : (e[0] == '0' ? 0 : 1);
}
static bool mma_class(const DwDims& d) {
    const bool shape_ok = d.kT == 1 && d.kH == d.kW && (d.kH == 3 || d.kH == 5) && d.sH == 1 && d.sW == 1 && d.sT == 1 &&
                          d.pH == d.pW && d.pH == d.kH / 2 && d.C % 8 == 0 && d.Ho == d.H && d.Wo == d.W;
    if (!shape_ok || dw_mma_mode() == 0) return false;
    if (dw_mma_mode() == 1) return true;
    return d.kH == 3 && d.W <= 14 && d.H >= 12 && d.C >= 400;
}

bool dw_fwd_mma(const __nv_bfloat16* x, const float* w_tc, __nv_bfloat16* y, const DwDims& d, cudaStream_t st) {
    if (!mma_class(d)) return false;
    if (((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) != 0) return false;
    return d.kH == 3 ? launch_mma<3>(x, w_tc, y, d.B, d.C, d.T, d.To, d.H, d.W, d.pT, 0, st)
                     : launch_mma<5>(x, w_tc, y, d.B, d.C, d.T, d.To, d.H, d.W, d.pT, 0, st);
}

// stride-1 input gradient == correlation of dy with the flipped filter: dx[t] reads dy[t + pT]
bool dw_dgrad_mma(const __nv_bfloat16* dy, const float* w_tc, __nv_bfloat16* dx, const DwDims& d, cudaStream_t st) {
    if (!mma_class(d)) return false;
    if (((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx)) & 15) != 0) return false;
    return d.kH == 3 ? launch_mma<3>(dy, w_tc, dx, d.B, d.C, d.To, d.T, d.H, d.W, -d.pT, 1, st)
                     : launch_mma<5>(dy, w_tc, dx, d.B, d.C, d.To, d.T, d.H, d.W, -d.pT, 1, st);
}

}  // namespace pb
