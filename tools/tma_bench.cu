// TMA halo-tile load micro-benchmark: how fast can persistent CTAs stream [Hi][Wi][Cb] boxes of an NDHWC bf16
// tensor into a shared-memory ring, as a function of the box row size (Cb*2 bytes), swizzle, ring depth and
// CTAs per SM?  No compute: the consumer warp only waits for the data and releases the stage.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/tma_bench tools/tma_bench.cu
//   tools/_build/tma_bench            (runs the built-in table; results feed the depthwise tile planner)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0, spins = 0;
    while (!ok) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (++spins > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
                 : "memory");
}

struct P {
    int B, T, H, W, C, Cb, Hi, Wi, Ht, Wt, pad, stages, stage_bytes, nblk, tiles_h, tiles_w;
    long long ntiles;
    int nprod;     // TMA-issuing threads per CTA (one per warp), tiles dealt round-robin
};

__global__ void __launch_bounds__(160) tma_stream_kernel(const __grid_constant__ CUtensorMap tm, const P p, unsigned* sink) {
    extern __shared__ uint8_t raw[];
    __shared__ uint64_t full[8], empty[8];
    const uint32_t a = smem_u32(raw);
    uint8_t* ring = raw + (((a + 1023u) & ~1023u) - a);
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long mine = p.ntiles > blockIdx.x ? (p.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int c_base = blockIdx.y * p.Cb;
    if (threadIdx.x >= 32 && (threadIdx.x & 31) == 0 && (int)(threadIdx.x >> 5) - 1 < p.nprod) {
        for (long long n = (threadIdx.x >> 5) - 1; n < mine; n += p.nprod) {
            const int s = (int)(n % p.stages);
            if (n >= p.stages) mbar_wait(&empty[s], (uint32_t)(((n / p.stages) - 1) & 1));
            long long t = blockIdx.x + n * (long long)gridDim.x;
            const int tw = (int)(t % p.tiles_w); t /= p.tiles_w;
            const int th = (int)(t % p.tiles_h); t /= p.tiles_h;
            const int f = (int)(t % p.T);
            const int b = (int)(t / p.T);
            mbar_expect_tx(&full[s], (uint32_t)(p.Hi * p.Wi * p.Cb * 2));
            tma_load_5d(ring + (size_t)s * p.stage_bytes, &tm, &full[s], c_base, tw * p.Wt - p.pad, th * p.Ht - p.pad, f, b);
        }
    } else if (threadIdx.x < 32) {
        unsigned acc = 0;
        for (long long n = 0; n < mine; ++n) {
            const int s = (int)(n % p.stages);
            mbar_wait(&full[s], (uint32_t)((n / p.stages) & 1));
            acc += *reinterpret_cast<const unsigned*>(ring + (size_t)s * p.stage_bytes + threadIdx.x * 4);
            __syncwarp();
            if (threadIdx.x == 0) mbar_arrive(&empty[s]);
        }
        if (acc == 0x12345678u) sink[0] = acc;
    }
}

static EncodeTiledFn enc() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    return (EncodeTiledFn)fn;
}

static void run(const char* name, int B, int T, int H, int W, int C, int Cb, int Ht, int Wt, int K, int swz, int ctas_per_sm,
                int stages, void* buf, unsigned* sink, int nprod = 1) {
    P p{};
    p.nprod = nprod;
    p.B = B; p.T = T; p.H = H; p.W = W; p.C = C; p.Cb = Cb; p.Ht = Ht; p.Wt = Wt; p.pad = K / 2;
    p.Hi = Ht + K - 1; p.Wi = Wt + K - 1;
    p.stages = stages;
    p.stage_bytes = (p.Hi * p.Wi * Cb * 2 + 1023) / 1024 * 1024;
    p.nblk = (C + Cb - 1) / Cb;
    p.tiles_h = (H + Ht - 1) / Ht; p.tiles_w = (W + Wt - 1) / Wt;
    p.ntiles = (long long)B * T * p.tiles_h * p.tiles_w;
    CUtensorMap tm;
    cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)T, (cuuint64_t)B};
    cuuint64_t str[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2, (cuuint64_t)T * H * W * C * 2};
    cuuint32_t box[5] = {(cuuint32_t)Cb, (cuuint32_t)p.Wi, (cuuint32_t)p.Hi, 1, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUtensorMapSwizzle sw = swz == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : swz == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swz == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUresult r = enc()(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, buf, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("%-34s encode failed (%d)\n", name, (int)r); return; }
    const size_t smem = (size_t)stages * p.stage_bytes + 1024;
    if (smem > 227 * 1024 / ctas_per_sm) { printf("%-34s does not fit (%zu B x %d)\n", name, smem, ctas_per_sm); return; }
    CK(cudaFuncSetAttribute(tma_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int ctas = 148 * ctas_per_sm / p.nblk;
    if (ctas < 1) ctas = 1;
    dim3 grid(ctas, p.nblk);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int it = 0; it < 5; ++it) {
        CK(cudaEventRecord(e0));
        tma_stream_kernel<<<grid, 32 + 32 * nprod, smem>>>(tm, p, sink);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (it > 0 && ms < best) best = ms;
    }
    CK(cudaGetLastError());
    const double uniq = (double)B * T * H * W * C * 2;
    const double boxb = (double)p.ntiles * p.nblk * p.Hi * p.Wi * Cb * 2;
    const double rows = (double)p.ntiles * p.nblk * p.Hi * p.Wi;
    const double clk = best * 1e-3 * 1.9e9;   // ~SM cycles
    printf("%-34s Cb=%3d box %2dx%2d rowB=%3d swz=%3d %dcta/SM st=%d np=%d | %7.1f us | unique %6.0f GB/s | box %6.0f GB/s | %5.2f clk/row/SM\n",
           name, Cb, p.Hi, p.Wi, Cb * 2, swz, ctas_per_sm, stages, nprod, best * 1000, uniq / best / 1e6, boxb / best / 1e6,
           clk / (rows / 148.0));
}

int main() {
    const size_t bytes = (size_t)1 << 30;
    void* buf; unsigned* sink;
    CK(cudaMalloc(&buf, bytes)); CK(cudaMemset(buf, 1, bytes)); CK(cudaMalloc(&sink, 4));
    if (getenv("TMA_BENCH_NPROD")) {   // does a CTA move more box rows per clock when several of its threads issue the loads?
        for (int np : {1, 2, 4})
            for (int cb : {32, 64, 128})
                run("producers per CTA C=256", 64, 8, 28, 28, 256, cb, 7, 14, 3, 0, 1, 4, buf, sink, np);
        for (int np : {1, 2, 4})
            run("producers per CTA swz128", 64, 8, 28, 28, 256, 64, 7, 14, 3, 128, 1, 4, buf, sink, np);
        return 0;
    }
    // row-size sweep on a 256-channel tensor, 3x3 halo, 7x14 tiles; then ring depth / CTAs per SM at fixed row sizes
    for (int cb : {16, 32, 64, 128, 256})
        run("sweep C=256 28x28 T=8 B=64", 64, 8, 28, 28, 256, cb, 7, 14, 3, 0, 1, 2, buf, sink);
    for (int st : {2, 3, 4, 6})
        for (int cb : {32, 64, 128})
            run("depth sweep C=256", 64, 8, 28, 28, 256, cb, 7, 14, 3, 0, 1, st, buf, sink);
    for (int cb : {32, 64, 128})
        run("2 CTAs/SM C=256", 64, 8, 28, 28, 256, cb, 7, 14, 3, 0, 2, 3, buf, sink);
    // MobileNetLarge3D stride-1 layers (input tensors), candidate tilings of the mma kernel (full-width tiles)
    run("b2.0 C=16 112x112 T=8", 64, 8, 112, 112, 16, 16, 8, 112, 3, 0, 1, 4, buf, sink);
    run("b2.0 C=16 112x112 T=8", 64, 8, 112, 112, 16, 16, 16, 56, 3, 0, 2, 3, buf, sink);
    run("b2.2 C=72 56x56 T=6", 64, 6, 56, 56, 72, 72, 8, 60, 3, 0, 1, 2, buf, sink);
    run("b2.2 C=72 56x56 T=6", 64, 6, 56, 56, 72, 72, 14, 28, 3, 0, 2, 2, buf, sink);
    run("b3.2 C=120 28x28 T=10 k5", 64, 10, 28, 28, 120, 64, 14, 28, 5, 0, 1, 2, buf, sink);
    run("b3.2 C=120 28x28 T=10 k5", 64, 10, 28, 28, 120, 120, 7, 28, 5, 0, 1, 2, buf, sink);
    run("b3.2 C=120 28x28 T=10 k5", 64, 10, 28, 28, 120, 120, 14, 14, 5, 0, 1, 2, buf, sink);
    run("b3.2 C=120 k5 Cb=40", 64, 10, 28, 28, 120, 40, 28, 28, 5, 0, 1, 2, buf, sink);
    run("b4.5 C=672 14x14 T=16", 64, 16, 14, 14, 672, 112, 14, 14, 3, 0, 1, 3, buf, sink);
    run("b4.5 C=672 14x14 T=16", 64, 16, 14, 14, 672, 168, 14, 14, 3, 0, 1, 2, buf, sink);
    run("b4.5 C=672 14x14 T=16", 64, 16, 14, 14, 672, 96, 14, 14, 3, 0, 2, 2, buf, sink);
    run("b4.5 C=672 14x14 T=16", 64, 16, 14, 14, 672, 224, 14, 14, 3, 0, 1, 2, buf, sink);
    run("b5.2 C=960 7x7 T=15 k5", 64, 15, 7, 7, 960, 120, 7, 7, 5, 0, 2, 3, buf, sink);
    run("b5.2 C=960 7x7 T=15 k5", 64, 15, 7, 7, 960, 240, 7, 7, 5, 0, 2, 2, buf, sink);
    run("b5.2 C=960 7x7 T=15 k5 12x16", 64, 15, 7, 7, 960, 120, 8, 12, 5, 0, 1, 4, buf, sink);
    run("swz128 Cb=64 C=256", 64, 8, 28, 28, 256, 64, 7, 14, 3, 128, 1, 4, buf, sink);
    return 0;
}
