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
#include <mutex>

#include "tc_common.cuh"

namespace pb {
namespace tc {

// rows (reduction) per pipeline stage: 64, or 128 for skinny layers (few boxes per stage) where the fixed
// per-stage cost would otherwise dominate; one TMA box = rows x 64 bf16 elements (rows x 128 bytes)
constexpr int WG_STAGES_MAX = 4;

struct WgradPlan {
    int NT;          // 128-row output tiles along n
    int KW;          // k-slice width per item (multiple of 16, <= 256)
    int k_groups;    // slices along k
    int per_group;   // n-tiles per item (per_group * KW <= 512)
    int n_groups;
    int kw_boxes;    // ceil(KW / 64)
    int chunks;      // row chunks per batch entry
    long long chunk_rows;   // multiple of rows
    int rows;               // reduction rows per stage (64 or 128)
    int box_bytes;          // rows * 128
    int stages;
    int stage_bytes;
    long long items;        // groups * chunks * Bt
};

static WgradPlan make_plan(int Bt, long long R, int K, int N) {
    WgradPlan p;
    p.NT = ceil_div(N, 128);
    int kg = ceil_div(K, 256);
    p.KW = (ceil_div(K, kg) + 15) / 16 * 16;
    p.k_groups = ceil_div(K, p.KW);
    p.kw_boxes = ceil_div(p.KW, 64);
    p.per_group = std::max(1, std::min(p.NT, 512 / p.KW));
    p.per_group = std::max(1, std::min(p.per_group, (13 - p.kw_boxes) / 2));   // two stages must fit in 216 KB
    p.n_groups = ceil_div(p.NT, p.per_group);
    p.rows = (p.per_group * 2 + p.kw_boxes) <= 4 ? 128 : 64;
    p.box_bytes = p.rows * 128;
    p.stage_bytes = (p.per_group * 2 + p.kw_boxes) * p.box_bytes;
    p.stages = std::max(2, std::min(WG_STAGES_MAX, (216 * 1024) / p.stage_bytes));
    // aim at ~4 items per SM so the persistent CTAs stay balanced, but keep >= 1024 rows per item
    long long groups = (long long)p.k_groups * p.n_groups;
    long long want = std::max<long long>(1, (148LL * 4 + Bt * groups - 1) / (Bt * groups));
    long long max_chunks = std::max<long long>(1, R / 1024);
    p.chunks = (int)std::min(want, max_chunks);
    p.chunk_rows = ((R + p.chunks - 1) / p.chunks + p.rows - 1) / p.rows * p.rows;
    p.chunks = (int)((R + p.chunk_rows - 1) / p.chunk_rows);
    p.items = groups * p.chunks * Bt;
    return p;
}

struct WgradParams {
    int Bt, K, N;
    long long R;
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
    const WgradPlan& pl = p.plan;
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* tiles = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int d_bytes = pl.per_group * 2 * pl.box_bytes;   // dC part of a stage

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
                const uint32_t tx = (uint32_t)((ntiles * 2 + pl.kw_boxes) * pl.box_bytes);
                for (int it = 0; it < iters; ++it) {
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    mbar_expect_tx(&full_bar[s], tx);
                    uint8_t* st = tiles + (size_t)s * pl.stage_bytes;
                    const int r = (int)(r_begin + (long long)it * pl.rows);
                    for (int j = 0; j < ntiles * 2; ++j)
                        tma_load_3d(st + j * pl.box_bytes, &tmD, &full_bar[s], (nt0 * 2 + j) * 64, r, w.b);
                    for (int j = 0; j < pl.kw_boxes; ++j)
                        tma_load_3d(st + d_bytes + j * pl.box_bytes, &tmA, &full_bar[s], k0 + j * 64, r, w.b);
                    if (++s == pl.stages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc(128, pl.KW, 1, 1);
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
                        for (int ks = 0; ks < pl.rows / 16; ++ks) {     // 16 rows = 2 swizzle atoms = 2 KB
                            const uint64_t adesc = make_desc(sb + j * 2 * pl.box_bytes + ks * 2048, pl.box_bytes, 1024);
                            const uint64_t bdesc = make_desc(sb + d_bytes + ks * 2048, pl.box_bytes, 1024);
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
            float* out = p.partial + ((long long)w.b * pl.chunks + w.chunk) * p.N * p.K;
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
    const int ol = threadIdx.x & 31, ps = threadIdx.x >> 5;
    const long long NK = (long long)N * K;
    const long long idx = (long long)blockIdx.x * 32 + ol;
    const int P = Bt * chunks;
    float acc = 0.f;
    if (idx < NK) {
        const int k = (int)(idx % K);
        for (int q = ps; q < P; q += 8) {
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

// dgate[b][k] = sum_n W[n][k] * P_b[n][k],  P_b = sum_c partial[b][c]
__global__ void wgrad_dgate_kernel(const float* __restrict__ partial, const float* __restrict__ W,
                                   float* __restrict__ dgate, int Bt, int chunks, int N, int K) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (k >= K) return;
    const long long NK = (long long)N * K;
    float acc = 0.f;
    for (int c = 0; c < chunks; ++c) {
        const float* pb = partial + ((long long)b * chunks + c) * NK;
        for (int n = 0; n < N; ++n) acc = fmaf(W[(long long)n * K + k], pb[(long long)n * K + k], acc);
    }
    dgate[(long long)b * K + k] = acc;
}

}  // namespace tc
}  // namespace pb

using namespace pb;
using namespace pb::tc;

extern "C" long long pb_pw_wgrad_tc_workspace_bytes(int Bt, long long R, int K, int N) {
    if (Bt <= 0 || R <= 0 || K <= 0 || N <= 0) return 0;
    WgradPlan pl = make_plan(Bt, R, K, N);
    return (long long)Bt * pl.chunks * N * K * (long long)sizeof(float);
}

extern "C" int pb_pw_wgrad_tc(const void* A, const void* dC, const float* gate, const float* Wf32, void* workspace,
                              float* dW, float* dgate, int Bt, long long R, int K, int N, pb_stream_t stream) {
    PB_REQUIRE(A && dC && workspace && dW, "pw_wgrad_tc: null pointer");
    PB_REQUIRE(Bt > 0 && Bt <= 65535 && R > 0 && K > 0 && N > 0, "pw_wgrad_tc: bad problem size");
    PB_REQUIRE(K % 8 == 0 && N % 8 == 0, "pw_wgrad_tc: K=%d and N=%d must be multiples of 8", K, N);
    PB_REQUIRE(!dgate || (Wf32 != nullptr), "pw_wgrad_tc: dgate needs the fp32 weights");
    WgradParams p;
    p.Bt = Bt; p.K = K; p.N = N; p.R = R;
    p.plan = make_plan(Bt, R, K, N);
    p.partial = (float*)workspace;
    const WgradPlan& pl = p.plan;
    CUtensorMap tmD, tmA;
    {
        uint64_t dims[3] = {(uint64_t)N, (uint64_t)R, (uint64_t)Bt};
        uint64_t str[3] = {2, (uint64_t)N * 2, (uint64_t)R * N * 2};
        uint32_t box[3] = {64, (uint32_t)pl.rows, 1};
        if (int e = make_tmap_bf16(&tmD, dC, 3, dims, str, box)) return e;
    }
    {
        uint64_t dims[3] = {(uint64_t)K, (uint64_t)R, (uint64_t)Bt};
        uint64_t str[3] = {2, (uint64_t)K * 2, (uint64_t)R * K * 2};
        uint32_t box[3] = {64, (uint32_t)pl.rows, 1};
        if (int e = make_tmap_bf16(&tmA, A, 3, dims, str, box)) return e;
    }
    static std::once_flag attr_once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(attr_once, [] {
        attr_err = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    });
    if (attr_err != cudaSuccess) return cuda_fail(attr_err, "cudaFuncSetAttribute(wgrad_tc_kernel)");
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = (int)std::min<long long>(pl.items, sms);
    const size_t smem = (size_t)pl.stages * pl.stage_bytes + 1024;
    wgrad_tc_kernel<<<grid, 256, smem, st>>>(tmD, tmA, p);
    PB_CHECK_LAUNCH("wgrad_tc_kernel");
    const long long NK = (long long)N * K;
    wgrad_reduce_kernel<<<ceil_div(NK, 32), 256, 0, st>>>(p.partial, gate, dW, Bt, pl.chunks, N, K);
    PB_CHECK_LAUNCH("wgrad_reduce_kernel");
    if (dgate) {
        dim3 g2(ceil_div(K, 128), Bt);
        wgrad_dgate_kernel<<<g2, 128, 0, st>>>(p.partial, Wf32, dgate, Bt, pl.chunks, N, K);
        PB_CHECK_LAUNCH("wgrad_dgate_kernel");
    }
    return PB_OK;
}
