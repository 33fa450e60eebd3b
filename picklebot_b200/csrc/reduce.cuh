// Column reductions over NDHWC matrices X[nbatch][R][C]: every thread owns 8 channels and strides over
// rows (4 rows in flight per iteration to cover HBM latency); fp32 per-thread partials -> shared-memory
// atomics per CTA -> one global atomic per (value, channel) per CTA (fp64 for BatchNorm statistics, fp32
// for pooling).
#pragma once
#include "common.cuh"

namespace pb {

constexpr int COLRED_UNROLL = 4;

// Fold the per-thread partials of one CTA over the row-slot index rr (threads tid and tid+h*G own the same
// channels): a halving tree through shared memory (channel-major, so every access is conflict-free).  On
// return the rr==0 threads hold the CTA totals.  Replaces per-thread shared-memory atomics, which serialise
// RPI-way on the same address when C is small.
template <int NACC>
__device__ __forceinline__ void cta_fold_rows(float (&acc)[NACC], int G, int RPI, int rr, float* sm /*[NACC][256]*/) {
    const int tid = threadIdx.x;
    for (int n = RPI; n > 1;) {
        const int half = (n + 1) >> 1;
        if (rr >= half && rr < n) {
#pragma unroll
            for (int j = 0; j < NACC; ++j) sm[j * 256 + tid] = acc[j];
        }
        __syncthreads();
        if (rr + half < n) {
#pragma unroll
            for (int j = 0; j < NACC; ++j) acc[j] += sm[j * 256 + tid + half * G];
        }
        __syncthreads();
        n = half;
    }
}

// F: struct with  __device__ void operator()(int batch, long long row_in_batch, int c0, float (&out)[NV][8]) const
template <typename F, int NV, typename OUT>
__global__ void __launch_bounds__(256)
colreduce_kernel(F f, long long R, int C, OUT* __restrict__ out, int nbatch, float out_scale) {
    pdl_trigger();
    pdl_wait();
    __shared__ float sm_fold[NV * 8 * 256];
    const int G = C >> 3;
    const int RPI = blockDim.x / G;
    const int g = threadIdx.x % G, rr = threadIdx.x / G;
    const int b = blockIdx.y;
    float acc[NV * 8];
#pragma unroll
    for (int i = 0; i < NV * 8; ++i) acc[i] = 0.f;
    const int c0 = g << 3;
    if (rr < RPI) {
        const long long stride = (long long)gridDim.x * RPI;
        long long r = (long long)blockIdx.x * RPI + rr;
        for (; r + (COLRED_UNROLL - 1) * stride < R; r += COLRED_UNROLL * stride) {
            float t[COLRED_UNROLL][NV][8];
#pragma unroll
            for (int u = 0; u < COLRED_UNROLL; ++u) f(b, r + u * stride, c0, t[u]);
#pragma unroll
            for (int u = 0; u < COLRED_UNROLL; ++u)
#pragma unroll
                for (int v = 0; v < NV; ++v)
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc[v * 8 + i] += t[u][v][i];
        }
        for (; r < R; r += stride) {
            float t[NV][8];
            f(b, r, c0, t);
#pragma unroll
            for (int v = 0; v < NV; ++v)
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[v * 8 + i] += t[v][i];
        }
    }
    cta_fold_rows<NV * 8>(acc, G, RPI, rr, sm_fold);
    if (rr == 0) {
#pragma unroll
        for (int v = 0; v < NV; ++v)
#pragma unroll
            for (int i = 0; i < 8; ++i)
                atomicAdd(&out[((long long)v * nbatch + b) * C + c0 + i], (OUT)(acc[v * 8 + i] * out_scale));
    }
}

// grid sizing: enough CTAs to fill the machine, never more than one CTA per RPI rows
static inline dim3 colreduce_grid(long long R, int C, int nbatch) {
    int G = C >> 3;
    int RPI = 256 / G;
    long long max_ctas = (R + RPI - 1) / RPI;
    long long want = (148LL * 8 + nbatch - 1) / nbatch;
    int gx = (int)(max_ctas < want ? max_ctas : want);
    if (gx < 1) gx = 1;
    return dim3(gx, nbatch);
}

}  // namespace pb
