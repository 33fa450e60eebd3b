// Pointwise-conv weight gradient on tcgen05:  P[n][k] = sum_m dC[m][n] * A[m][k]  (fp32 in TMEM).
//
// The reduction runs over the rows m (up to 6.4 M), the output is tiny (N x K <= 960 x 160), so both MMA
// operands are "MN-major": the row-major activations dC[m][n] and A[m][k] are exactly the transposes the MMA
// needs, fetched by TMA as rows x 128-byte boxes (no transposes in HBM).  A work item is one group of output
// tiles (up to 512 TMEM columns = 128 x 512 accumulators) times one chunk of rows; persistent CTAs (one per
// SM) walk over the items, stream each item's rows once through a TMA ring, then dump the fp32 accumulators
// to a workspace slice.  A second, small kernel sums the slices and, for squeeze-excite blocks, applies
// the gate per sample and produces the gate gradient from the same per-sample products (so the backward
// pass never re-reads the expanded activations for it):
//     dW[n][k]    = sum_b gate[b][k] * P_b[n][k]
//     dgate[b][k] = sum_n W[n][k]    * P_b[n][k]
// Replaces the weight-gradient half of nn.Conv3d(kernel_size=1) autograd (train.py:269).
#include <algorithm>
#include <cstdlib>
#include <mutex>

#include "tc_common.cuh"

namespace pb {
namespace tc {

// Operand boxes: one TMA box = rows x ew channels, ew in {16, 32, 64} (swizzle 32/64/128 bytes) picked per
// operand so that narrow layers (16..112 channels) do not fill shared memory with out-of-bounds zeros: the
// bytes in flight per SM are what bounds a streaming reduction, so every staged byte should be a useful one.
// rows (reduction) per pipeline stage: 64, 128 or 256 -- the fewer columns a stage holds the more rows.
constexpr int WG_STAGES_MAX = 8;

struct WgradPlan {
    int fold;        // F consecutive rows are viewed as one row of F*C channels (see make_plan)
    int NT;          // 128-row output tiles along n
    int KW;          // k-slice width per item (multiple of 16, <= 256)
    int k_groups;    // slices along k
    int per_group;   // n-tiles per item (per_group * KW <= 512)
    int n_groups;
    int ew_d, ew_a;  // box width (channels) of the dC / A operand
    int nb_d;        // dC boxes per 128-wide n-tile
    int kw_boxes;    // A boxes per item: ceil(KW / ew_a)
    int chunks;      // row chunks per batch entry
    long long chunk_rows;   // multiple of rows
    int rows;               // reduction rows per stage
    int box_bytes_d, box_bytes_a;
    int stages;
    int stage_bytes;
    int slack_bytes;        // the 128-wide MMA operand of a narrow dC tile reads past the boxes that were loaded
    long long items;        // groups * chunks * Bt
};

// box width for an operand of W channels: least padding, fewer boxes on ties
static int pick_box_width(int W) {
    if (W > 128) return 64;
    int best = 64, best_cost = 1 << 30;
    for (int ew : {64, 32, 16}) {
        const int boxes = ceil_div(W, ew);
        // a 32-byte box row still costs a 64-byte transfer from L2 (ncu: 424 MB moved for 231 MB on 24 -> 72 channels
        // with 16-wide boxes), so boxes narrower than 32 channels are charged as 32
        static const bool old_cost = getenv("PB_WGRAD_OLD_BOXCOST") != nullptr;
        const int cost = boxes * (old_cost ? ew : std::max(ew, 32)) + 4 * boxes;
        if (cost < best_cost) { best_cost = cost; best = ew; }
    }
    return best;
}

// Row folding.  The TMA unit spends about as long on a 32-byte box row as on a 128-byte one, so a 16- or
// 32-channel activation matrix cannot be streamed at HBM speed row by row.  Since X[M][C] is contiguous it is
// also X'[M/F][F*C]; the product of the folded operands, P'[(i,n)][(j,k)] = sum_r' dC[F r'+i][n] A[F r'+j][k],
// holds the wanted sum in its F diagonal blocks (i == j) -- the off-diagonal MMA work is free, the kernel
// is nowhere near tensor-bound -- and the epilogue writes block i as one more partial slice.
static int pick_fold(long long R, int K, int N) {
    if (std::min(K, N) > 32) return 1;
    for (int F : {4, 2})
        if (R % F == 0 && K * F <= 128 && N * F <= 128) return F;
    // 24 <-> 72 channels: unfolded, the 48-byte rows of one operand and the 32-byte boxes of the other make TMA move
    // 424 MB for 231 MB of activations (ncu: the L2 -> SM fabric is the limit at 7 TB/s); folded by two the rows are
    // 96 / 288 bytes in 128-byte boxes (1.33x), at the price of a second 128-row output tile
    if (R % 2 == 0 && K * 2 <= 256 && N * 2 <= 256 && !getenv("PB_WGRAD_NO_WIDE_FOLD")) return 2;
    return 1;
}

static WgradPlan make_plan(int Bt, long long R, int K, int N) {
    WgradPlan p;
    p.fold = pick_fold(R, K, N);
    R /= p.fold; K *= p.fold; N *= p.fold;      // from here on: the folded problem
    p.NT = ceil_div(N, 128);
    int kg = ceil_div(K, 256);
    p.KW = (ceil_div(K, kg) + 15) / 16 * 16;
    p.k_groups = ceil_div(K, p.KW);
    p.ew_d = pick_box_width(N);
    p.ew_a = pick_box_width(p.k_groups > 1 ? 256 : K);
    p.nb_d = p.NT > 1 ? 128 / p.ew_d : ceil_div(N, p.ew_d);
    p.kw_boxes = ceil_div(p.KW, p.ew_a);
    const int tile_cols = p.nb_d * p.ew_d, a_cols = p.kw_boxes * p.ew_a;
    p.per_group = std::max(1, std::min(p.NT, 512 / p.KW));
    // two 64-row stages must fit in 216 KB
    p.per_group = std::max(1, std::min(p.per_group, (216 * 1024 / (2 * 64 * 2) - a_cols) / tile_cols));
    p.n_groups = ceil_div(p.NT, p.per_group);
    const int cols = p.per_group * tile_cols + a_cols;
    p.rows = cols <= 64 ? 256 : cols <= 256 ? 128 : 64;
    p.box_bytes_d = p.rows * p.ew_d * 2;
    p.box_bytes_a = p.rows * p.ew_a * 2;
    p.stage_bytes = cols * p.rows * 2;
    p.slack_bytes = (128 / p.ew_d - p.nb_d) * p.box_bytes_d;
    p.stages = std::max(2, std::min(WG_STAGES_MAX, (216 * 1024 - p.slack_bytes) / p.stage_bytes));
    // Row chunks: items are equal-sized and statically assigned, so what matters is that their count fills
    // whole waves of 148 CTAs; fewer, longer items also mean fewer fp32 partials to write and re-read.
    // Keep >= 1024 rows per item.
    const long long groups = (long long)p.k_groups * p.n_groups;
    const long long unit = Bt * groups;
    const long long max_chunks = std::max<long long>(1, R / 1024);
    long long best = std::min(max_chunks, std::max<long long>(1, 148 / unit));
    if (unit * max_chunks > 148) {
        double best_eff = 0;
        for (int w = 1; w <= 4; ++w) {
            const long long c = std::min(max_chunks, (148LL * w) / unit);
            if (c < 1) continue;
            const long long it = unit * c;
            const double eff = (double)it / (148.0 * (double)((it + 147) / 148));
            if (eff > best_eff + 0.03) { best_eff = eff; best = c; }
        }
    }
    p.chunks = (int)best;
    p.chunk_rows = ((R + p.chunks - 1) / p.chunks + p.rows - 1) / p.rows * p.rows;
    p.chunks = (int)((R + p.chunk_rows - 1) / p.chunk_rows);
    p.items = groups * p.chunks * Bt;
    return p;
}

struct WgradParams {
    int Bt, K, N;       // K, N: the real (unfolded) channel counts
    long long R;        // rows per batch entry of the folded problem
    WgradPlan plan;
    float* partial;     // [Bt][chunks][N][K]
};

struct WgItem { int kg, ng, chunk, b; };

__device__ __forceinline__ WgItem decode_item(const WgradPlan& pl, long long it) {
    WgItem w;
    const int groups = pl.k_groups * pl.n_groups;
    const int gidx = (int)(it % groups); it /= groups;
    w.kg = gidx % pl.k_groups; w.ng = gidx / pl.k_groups;
    w.chunk = (int)(it % pl.chunks);
    w.b = (int)(it / pl.chunks);
    return w;
}

__global__ void __launch_bounds__(256, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmA, WgradParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_bar[WG_STAGES_MAX], empty_bar[WG_STAGES_MAX], done_bar, tfree_bar;
    __shared__ uint32_t tmem_base_s;
    pdl_trigger();
    const WgradPlan& pl = p.plan;
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* tiles = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile_bytes_d = pl.nb_d * pl.box_bytes_d;
    const int d_bytes = pl.per_group * tile_bytes_d;       // dC part of a stage

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmD);
        tma_prefetch_desc(&tmA);
        for (int s = 0; s < pl.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&done_bar, 1);
        mbar_init(&tfree_bar, 128);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(&tmem_base_s, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    pdl_wait();

    if (warp == 0) {
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (long long item = blockIdx.x; item < pl.items; item += gridDim.x) {
                const WgItem w = decode_item(pl, item);
                const int nt0 = w.ng * pl.per_group;
                const int ntiles = min(pl.per_group, pl.NT - nt0);
                const int k0 = w.kg * pl.KW;
                const long long r_begin = (long long)w.chunk * pl.chunk_rows;
                const long long r_end = min(p.R, r_begin + pl.chunk_rows);
                const int iters = (int)((r_end - r_begin + pl.rows - 1) / pl.rows);
                const uint32_t tx = (uint32_t)(ntiles * tile_bytes_d + pl.kw_boxes * pl.box_bytes_a);
                for (int it = 0; it < iters; ++it) {
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    mbar_expect_tx(&full_bar[s], tx);
                    uint8_t* st = tiles + (size_t)s * pl.stage_bytes;
                    const int r = (int)(r_begin + (long long)it * pl.rows);
                    for (int j = 0; j < ntiles * pl.nb_d; ++j)
                        tma_load_3d(st + j * pl.box_bytes_d, &tmD, &full_bar[s], nt0 * 128 + j * pl.ew_d, r, w.b);
                    for (int j = 0; j < pl.kw_boxes; ++j)
                        tma_load_3d(st + d_bytes + j * pl.box_bytes_a, &tmA, &full_bar[s], k0 + j * pl.ew_a, r, w.b);
                    if (++s == pl.stages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc(128, pl.KW, 1, 1);
            // MN-major canonical layouts: an atom is 8 rows x ew channels (16*ew bytes); SBO = atom stride along
            // the rows, LBO = stride between ew-wide channel groups (= one box); one MMA eats 16 rows = 2 atoms
            const uint32_t atom_d = 16u * pl.ew_d, atom_a = 16u * pl.ew_a;
            const uint32_t type_d = pl.ew_d == 64 ? 2u : pl.ew_d == 32 ? 4u : 6u;
            const uint32_t type_a = pl.ew_a == 64 ? 2u : pl.ew_a == 32 ? 4u : 6u;
            int s = 0; uint32_t ph = 0;
            uint32_t n_item = 0;
            for (long long item = blockIdx.x; item < pl.items; item += gridDim.x, ++n_item) {
                const WgItem w = decode_item(pl, item);
                const int nt0 = w.ng * pl.per_group;
                const int ntiles = min(pl.per_group, pl.NT - nt0);
                const long long r_begin = (long long)w.chunk * pl.chunk_rows;
                const long long r_end = min(p.R, r_begin + pl.chunk_rows);
                const int iters = (int)((r_end - r_begin + pl.rows - 1) / pl.rows);
                if (n_item > 0) {                       // the epilogue must have drained TMEM of the previous item
                    mbar_wait(&tfree_bar, (n_item - 1) & 1);
                    tc_fence_after();
                }
                for (int it = 0; it < iters; ++it) {
                    mbar_wait(&full_bar[s], ph);
                    tc_fence_after();
                    const uint32_t sb = smem_u32(tiles + (size_t)s * pl.stage_bytes);
                    for (int j = 0; j < ntiles; ++j) {
                        for (int ks = 0; ks < pl.rows / 16; ++ks) {
                            const uint64_t adesc = make_desc(sb + j * tile_bytes_d + ks * 2 * atom_d, pl.box_bytes_d,
                                                             atom_d, type_d);
                            const uint64_t bdesc = make_desc(sb + d_bytes + ks * 2 * atom_a, pl.box_bytes_a, atom_a,
                                                             type_a);
                            umma_bf16(tmem_base + (uint32_t)(j * pl.KW), adesc, bdesc, idesc, (it | ks) != 0);
                        }
                    }
                    umma_commit(&empty_bar[s]);
                    if (++s == pl.stages) { s = 0; ph ^= 1; }
                }
                umma_commit(&done_bar);
            }
        }
    } else if (warp >= 4) {
        const int q = warp & 3;
        uint32_t n_item = 0;
        for (long long item = blockIdx.x; item < pl.items; item += gridDim.x, ++n_item) {
            const WgItem w = decode_item(pl, item);
            const int nt0 = w.ng * pl.per_group;
            const int ntiles = min(pl.per_group, pl.NT - nt0);
            const int k0 = w.kg * pl.KW;
            mbar_wait(&done_bar, n_item & 1);
            tc_fence_after();
            const long long slice = (long long)w.b * pl.chunks + w.chunk;
            if (pl.fold == 1) {
                float* out = p.partial + slice * p.N * p.K;
                for (int j = 0; j < ntiles; ++j) {
                    const int n = (nt0 + j) * 128 + q * 32 + lane;
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * pl.KW);
                    for (int c0 = 0; c0 < pl.KW; c0 += 16) {
                        uint32_t r[16];
                        tmem_ld16(taddr + (uint32_t)c0, r);
                        tmem_ld_wait();
                        const int k = k0 + c0;
                        if (n < p.N && k < p.K) {
                            float* dst = out + (long long)n * p.K + k;
                            const int nv = min(16, p.K - k);            // 8 or 16
                            *reinterpret_cast<uint4*>(dst) = make_uint4(r[0], r[1], r[2], r[3]);
                            *reinterpret_cast<uint4*>(dst + 4) = make_uint4(r[4], r[5], r[6], r[7]);
                            if (nv > 8) {
                                *reinterpret_cast<uint4*>(dst + 8) = make_uint4(r[8], r[9], r[10], r[11]);
                                *reinterpret_cast<uint4*>(dst + 12) = make_uint4(r[12], r[13], r[14], r[15]);
                            }
                        }
                    }
                }
            } else {
                // folded: TMEM lane (i, n) of tile j (folded row j*128 + lane) keeps the columns of diagonal block i
                // -> partial slice i
                for (int j = 0; j < ntiles; ++j) {
                    const int np = (nt0 + j) * 128 + q * 32 + lane;
                    const int i = np / p.N, n = np - i * p.N;
                    const bool live = i < pl.fold;
                    float* dst = p.partial + ((slice * pl.fold + i) * p.N + n) * p.K;
                    const int lo = i * p.K, hi = lo + p.K;
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * pl.KW);
                    for (int c0 = 0; c0 < pl.KW; c0 += 16) {
                        uint32_t r[16];
                        tmem_ld16(taddr + (uint32_t)c0, r);
                        tmem_ld_wait();
#pragma unroll
                        for (int e = 0; e < 16; ++e) {
                            const int col = c0 + e;
                            if (live && col >= lo && col < hi) dst[col - lo] = __uint_as_float(r[e]);
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&tfree_bar);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// dW[n][k] = sum_b gate[b][k] * sum_c partial[b][c][n][k].
// blockDim = (32 outputs, 8 slices of the Bt*chunks partial list): a skinny layer can have ~600 partials of a
// few hundred floats, so the sum over partials is parallel too; shared-memory reduction over the slices.
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ partial, const float* __restrict__ gate,
                    float* __restrict__ dW, int Bt, int chunks, int N, int K) {
    __shared__ float red[8][33];
    pdl_trigger();
    pdl_wait();
    const int ol = threadIdx.x & 31, ps = threadIdx.x >> 5;
    const long long NK = (long long)N * K;
    const long long idx = (long long)blockIdx.x * 32 + ol;
    const int P = Bt * chunks;
    float acc = 0.f;
    if (idx < NK) {
        const int k = (int)(idx % K);
        int q = ps;
        for (; q + 24 < P; q += 32) {                         // four partials in flight per thread
            const float v0 = partial[(long long)q * NK + idx], v1 = partial[(long long)(q + 8) * NK + idx];
            const float v2 = partial[(long long)(q + 16) * NK + idx], v3 = partial[(long long)(q + 24) * NK + idx];
            if (gate) {
                acc = fmaf(__ldg(gate + (long long)(q / chunks) * K + k), v0, acc);
                acc = fmaf(__ldg(gate + (long long)((q + 8) / chunks) * K + k), v1, acc);
                acc = fmaf(__ldg(gate + (long long)((q + 16) / chunks) * K + k), v2, acc);
                acc = fmaf(__ldg(gate + (long long)((q + 24) / chunks) * K + k), v3, acc);
            } else {
                acc += (v0 + v1) + (v2 + v3);
            }
        }
        for (; q < P; q += 8) {
            const float v = partial[(long long)q * NK + idx];
            acc = gate ? fmaf(gate[(long long)(q / chunks) * K + k], v, acc) : acc + v;
        }
    }
    red[ps][ol] = acc;
    __syncthreads();
    if (ps == 0 && idx < NK) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += red[i][ol];
        dW[idx] = s;
    }
}

// dgate[b][k] = sum_n W[n][k] * P_b[n][k],  P_b = sum_c partial[b][c].
// blockDim = (32 k, 8 slices of the chunks*N rows); shared-memory reduction over the slices.
__global__ void __launch_bounds__(256)
wgrad_dgate_kernel(const float* __restrict__ partial, const float* __restrict__ W,
                   float* __restrict__ dgate, int Bt, int chunks, int N, int K) {
    __shared__ float red[8][33];
    pdl_trigger();
    pdl_wait();
    const int kl = threadIdx.x & 31, sl = threadIdx.x >> 5;
    const int k = blockIdx.x * 32 + kl;
    const int b = blockIdx.y;
    const long long NK = (long long)N * K;
    float acc = 0.f;
    if (k < K) {
        for (int c = 0; c < chunks; ++c) {
            const float* pb = partial + ((long long)b * chunks + c) * NK + k;
            const float* wk = W + k;
            int n = sl;
            for (; n + 24 < N; n += 32) {                   // four rows in flight per thread
                const float p0 = pb[(long long)n * K], p1 = pb[(long long)(n + 8) * K];
                const float p2 = pb[(long long)(n + 16) * K], p3 = pb[(long long)(n + 24) * K];
                acc = fmaf(__ldg(wk + (long long)n * K), p0, acc);
                acc = fmaf(__ldg(wk + (long long)(n + 8) * K), p1, acc);
                acc = fmaf(__ldg(wk + (long long)(n + 16) * K), p2, acc);
                acc = fmaf(__ldg(wk + (long long)(n + 24) * K), p3, acc);
            }
            for (; n < N; n += 8) acc = fmaf(__ldg(wk + (long long)n * K), pb[(long long)n * K], acc);
        }
    }
    red[sl][kl] = acc;
    __syncthreads();
    if (sl == 0 && k < K) {
        float v = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) v += red[i][kl];
        dgate[(long long)b * K + k] = v;
    }
}

}  // namespace tc
}  // namespace pb

using namespace pb;
using namespace pb::tc;

extern "C" long long pb_pw_wgrad_tc_workspace_bytes(int Bt, long long R, int K, int N) {
    if (Bt <= 0 || R <= 0 || K <= 0 || N <= 0) return 0;
    WgradPlan pl = make_plan(Bt, R, K, N);
    return (long long)Bt * pl.chunks * pl.fold * N * K * (long long)sizeof(float);
}

extern "C" int pb_pw_wgrad_tc(const void* A, const void* dC, const float* gate, const float* Wf32, void* workspace,
                              float* dW, float* dgate, int Bt, long long R, int K, int N, pb_stream_t stream) {
    PB_REQUIRE(A && dC && workspace && dW, "pw_wgrad_tc: null pointer");
    PB_REQUIRE(Bt > 0 && Bt <= 65535 && R > 0 && K > 0 && N > 0, "pw_wgrad_tc: bad problem size");
    PB_REQUIRE(K % 8 == 0 && N % 8 == 0, "pw_wgrad_tc: K=%d and N=%d must be multiples of 8", K, N);
    PB_REQUIRE(!dgate || (Wf32 != nullptr), "pw_wgrad_tc: dgate needs the fp32 weights");
    WgradParams p;
    p.Bt = Bt; p.K = K; p.N = N;
    p.plan = make_plan(Bt, R, K, N);
    p.R = R / p.plan.fold;
    p.partial = (float*)workspace;
    const WgradPlan& pl = p.plan;
    const uint64_t F = (uint64_t)pl.fold;
    CUtensorMap tmD, tmA;
    {
        uint64_t dims[3] = {(uint64_t)N * F, (uint64_t)R / F, (uint64_t)Bt};
        uint64_t str[3] = {2, (uint64_t)N * F * 2, (uint64_t)R * N * 2};
        uint32_t box[3] = {(uint32_t)pl.ew_d, (uint32_t)pl.rows, 1};
        if (int e = make_tmap_bf16(&tmD, dC, 3, dims, str, box, pl.ew_d * 2)) return e;
    }
    {
        uint64_t dims[3] = {(uint64_t)K * F, (uint64_t)R / F, (uint64_t)Bt};
        uint64_t str[3] = {2, (uint64_t)K * F * 2, (uint64_t)R * K * 2};
        uint32_t box[3] = {(uint32_t)pl.ew_a, (uint32_t)pl.rows, 1};
        if (int e = make_tmap_bf16(&tmA, A, 3, dims, str, box, pl.ew_a * 2)) return e;
    }
    static unsigned long long attr_done = 0;
    if (cudaError_t attr_err = ensure_dyn_smem(wgrad_tc_kernel, 226 * 1024, &attr_done))
        return cuda_fail(attr_err, "cudaFuncSetAttribute(wgrad_tc_kernel)");
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = (int)std::min<long long>(pl.items, sms);
    const size_t smem = (size_t)pl.stages * pl.stage_bytes + pl.slack_bytes + 1024;
    PB_CUDA(launch_pdl(wgrad_tc_kernel, dim3(grid), dim3(256), smem, st, tmD, tmA, p));
    PB_CHECK_LAUNCH("wgrad_tc_kernel");
    const long long NK = (long long)N * K;
    PB_CUDA(launch_pdl(wgrad_reduce_kernel, dim3(ceil_div(NK, 32)), dim3(256), 0, st, (const float*)p.partial, gate, dW, Bt,
                       pl.chunks * pl.fold, N, K));
    PB_CHECK_LAUNCH("wgrad_reduce_kernel");
    if (dgate) {
        // (a fused reduce + dgate kernel, one pass over the partials, was measured slower than these two: 43-60 us
        // against 12 + 20 us, its 128-byte pieces per (sample, row) do not stream)
        dim3 g2(ceil_div(K, 32), Bt);
        PB_CUDA(launch_pdl(wgrad_dgate_kernel, g2, dim3(256), 0, st, (const float*)p.partial, Wf32, dgate, Bt,
                           pl.chunks * pl.fold, N, K));
        PB_CHECK_LAUNCH("wgrad_dgate_kernel");
    }
    count_path(PB_PATH_WGRAD_TC);
    return PB_OK;
}
