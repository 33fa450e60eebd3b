// BatchNorm (train/eval) + activation + Dropout3d mask, forward and backward, on NDHWC matrices.
// Replaces nn.BatchNorm3d / BatchNorm1d, nn.Hardswish / ReLU / LeakyReLU and nn.Dropout3d call sites
// (mobilenet.py:80-82,90-92,142-143,180-181,247-248; movinet.py:65,75-76,93,141-143,150-152).
//
// All row-streaming kernels share one thread layout: a thread owns 8 fixed channels (so the per-channel
// constants live in registers for the whole kernel) and walks over rows with a CTA-wide stride; the
// activation is a template parameter, index math is 32-bit.
#include <algorithm>
#include <cstdlib>

#include "reduce.cuh"
#include "tc_common.cuh"

namespace pb {

__global__ void bn_finalize_kernel(const double* __restrict__ sums, double M, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ rmean,
                                   float* __restrict__ rvar, int training, float momentum, float eps,
                                   float* __restrict__ scale, float* __restrict__ shift,
                                   float* __restrict__ mean_o, float* __restrict__ invstd_o,
                                   long long* __restrict__ nbt, int C) {
    pdl_trigger();
    pdl_wait();
    // momentum < 0: nn.BatchNorm(momentum=None), the cumulative moving average with factor 1/num_batches_tracked
    // (after the increment).  That mode is launched as ONE block so that every thread reads the old counter
    // before thread 0 advances it.
    if (momentum < 0.f) {
        const long long seen = nbt ? *nbt : 0;
        momentum = 1.f / (float)(seen + 1);
        __syncthreads();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && nbt && training) *nbt += 1;           // nn.BatchNorm's num_batches_tracked
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gridDim.x * blockDim.x) {
    float mean, var;
    if (training) {
        double s0 = 0, s1 = 0;
        for (int r = 0; r < PB_STAT_REPLICAS; ++r) { s0 += sums[r * 2 * C + c]; s1 += sums[r * 2 * C + C + c]; }
        double mu = s0 / M;
        double v = s1 / M - mu * mu;
        if (v < 0) v = 0;
        mean = (float)mu; var = (float)v;
        if (rmean) {
            double unbiased = M > 1 ? v * (M / (M - 1.0)) : v;
            rmean[c] = (1.f - momentum) * rmean[c] + momentum * mean;
            rvar[c]  = (1.f - momentum) * rvar[c] + momentum * (float)unbiased;
        }
    } else {
        mean = rmean[c]; var = rvar[c];
    }
    float invstd = rsqrtf(var + eps);
    // rsqrtf is 2 ulp; one Newton step brings it to fp32 round-off (matters for the 1e-4 parity mode)
    invstd = invstd * (1.5f - 0.5f * (var + eps) * invstd * invstd);
    float g = gamma ? gamma[c] : 1.f;
    float bta = beta ? beta[c] : 0.f;
    float sc = g * invstd;
    scale[c] = sc;
    shift[c] = bta - mean * sc;
    if (mean_o) mean_o[c] = mean;
    if (invstd_o) invstd_o[c] = invstd;
    }
}

template <int ACT> __device__ __forceinline__ float actf(float u, float slope) { return act_fwd(u, ACT, slope); }
template <int ACT> __device__ __forceinline__ float actg(float u, float slope) { return act_grad(u, ACT, slope); }

struct RowLayout {
    int G, RPI, g, rr, c0;
    bool active;
    __device__ __forceinline__ RowLayout(int C) {
        G = C >> 3; RPI = blockDim.x / G;
        g = threadIdx.x % G; rr = threadIdx.x / G; c0 = g << 3;
        active = rr < RPI;
    }
};

__device__ __forceinline__ void load_vec8(const float* p, float (&v)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// out = act(z*scale + shift) * mask[b][c]
template <typename T, int ACT>
__global__ void __launch_bounds__(256)
bn_act_fwd_kernel(const T* __restrict__ z, const float* __restrict__ scale, const float* __restrict__ shift,
                  const float* __restrict__ mask, T* __restrict__ out, unsigned M, unsigned R, int C, float slope) {
    pdl_trigger();
    pdl_wait();
    const RowLayout L(C);
    if (!L.active) return;
    float sc[8], sh[8];
    load_vec8(scale + L.c0, sc);
    load_vec8(shift + L.c0, sh);
    const unsigned stride = gridDim.x * L.RPI;
    auto finish = [&](unsigned m, F8 v) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v.v[i] = actf<ACT>(fmaf(v.v[i], sc[i], sh[i]), slope);
        if (mask) {
            float mk[8];
            load_vec8(mask + (size_t)(m / R) * C + L.c0, mk);
#pragma unroll
            for (int i = 0; i < 8; ++i) v.v[i] *= mk[i];
        }
        store8(out + (size_t)m * C + L.c0, v);
    };
    unsigned m = blockIdx.x * L.RPI + L.rr;
    for (; m + stride < M; m += 2 * stride) {               // two rows in flight per thread
        const F8 v0 = load8(z + (size_t)m * C + L.c0);
        const F8 v1 = load8(z + (size_t)(m + stride) * C + L.c0);
        finish(m, v0);
        finish(m + stride, v1);
    }
    if (m < M) finish(m, load8(z + (size_t)m * C + L.c0));
}

// du = dout * mask * act'(u), u = z*scale + shift, xhat = z*a + bb with a = invstd, bb = -mean*invstd
template <typename T, int ACT, bool BCAST>
__device__ __forceinline__ void bn_bwd_elem(const void* dout, const T* z, const float* mask, unsigned m, unsigned R,
                                            int C, int c0, const float (&sc)[8], const float (&sh)[8],
                                            const float (&a)[8], const float (&bb)[8], float slope,
                                            float (&du)[8], float (&xh)[8]) {
    const size_t o = (size_t)m * C + c0;
    const F8 zv = load8(z + o);
    F8 dv;
    const unsigned b = (BCAST || mask) ? m / R : 0;
    if (BCAST) dv = load8(reinterpret_cast<const float*>(dout) + (size_t)b * C + c0);
    else       dv = load8(reinterpret_cast<const T*>(dout) + o);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float u = fmaf(zv.v[i], sc[i], sh[i]);
        du[i] = dv.v[i] * actg<ACT>(u, slope);
        xh[i] = fmaf(zv.v[i], a[i], bb[i]);
    }
    if (mask) {
        float mk[8];
        load_vec8(mask + (size_t)b * C + c0, mk);
#pragma unroll
        for (int i = 0; i < 8; ++i) du[i] *= mk[i];
    }
}

// pass 1: sums[0][c] = sum du, sums[1][c] = sum du*xhat
template <typename T, int ACT, bool BCAST>
__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(const void* __restrict__ dout, const T* __restrict__ z, const float* __restrict__ scale,
                     const float* __restrict__ shift, const float* __restrict__ mean,
                     const float* __restrict__ invstd, const float* __restrict__ mask, double* __restrict__ sums,
                     unsigned M, unsigned R, int C, float slope) {
    pdl_trigger();
    pdl_wait();
    __shared__ float sm_fold[16 * 256];
    const RowLayout L(C);
    float s[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) s[i] = 0.f;
    if (L.active) {
        float sc[8], sh[8], a[8], bb[8];
        load_vec8(scale + L.c0, sc);
        load_vec8(shift + L.c0, sh);
        load_vec8(invstd + L.c0, a);
        load_vec8(mean + L.c0, bb);
#pragma unroll
        for (int i = 0; i < 8; ++i) bb[i] = -bb[i] * a[i];
        const unsigned stride = gridDim.x * L.RPI;
        unsigned m = blockIdx.x * L.RPI + L.rr;
        for (; m + stride < M; m += 2 * stride) {           // two rows in flight
            float d0[8], x0[8], d1[8], x1[8];
            bn_bwd_elem<T, ACT, BCAST>(dout, z, mask, m, R, C, L.c0, sc, sh, a, bb, slope, d0, x0);
            bn_bwd_elem<T, ACT, BCAST>(dout, z, mask, m + stride, R, C, L.c0, sc, sh, a, bb, slope, d1, x1);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                s[i] += d0[i] + d1[i];
                s[8 + i] = fmaf(d0[i], x0[i], fmaf(d1[i], x1[i], s[8 + i]));
            }
        }
        if (m < M) {
            float d0[8], x0[8];
            bn_bwd_elem<T, ACT, BCAST>(dout, z, mask, m, R, C, L.c0, sc, sh, a, bb, slope, d0, x0);
#pragma unroll
            for (int i = 0; i < 8; ++i) { s[i] += d0[i]; s[8 + i] = fmaf(d0[i], x0[i], s[8 + i]); }
        }
    }
    cta_fold_rows<16>(s, L.G, L.RPI, L.rr, sm_fold);
    if (L.rr == 0) {
        double* dst = sums + (size_t)(blockIdx.x % PB_STAT_REPLICAS) * 2 * C;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            atomicAdd(&dst[L.c0 + i], (double)s[i]);
            atomicAdd(&dst[C + L.c0 + i], (double)s[8 + i]);
        }
    }
}

// batch statistics: sums[0][c] = sum z, sums[1][c] = sum z^2 (four rows in flight per thread)
template <typename T>
__global__ void __launch_bounds__(256)
colstats_kernel(const T* __restrict__ z, double* __restrict__ sums, unsigned M, int C) {
    pdl_trigger();
    pdl_wait();
    __shared__ float sm_fold[16 * 256];
    const RowLayout L(C);
    float s[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) s[i] = 0.f;
    if (L.active) {
        const unsigned stride = gridDim.x * L.RPI;
        unsigned m = blockIdx.x * L.RPI + L.rr;
        for (; m + 3 * stride < M; m += 4 * stride) {
            F8 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = load8(z + (size_t)(m + u * stride) * C + L.c0);
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i) { s[i] += v[u].v[i]; s[8 + i] = fmaf(v[u].v[i], v[u].v[i], s[8 + i]); }
        }
        for (; m < M; m += stride) {
            const F8 v = load8(z + (size_t)m * C + L.c0);
#pragma unroll
            for (int i = 0; i < 8; ++i) { s[i] += v.v[i]; s[8 + i] = fmaf(v.v[i], v.v[i], s[8 + i]); }
        }
    }
    cta_fold_rows<16>(s, L.G, L.RPI, L.rr, sm_fold);
    if (L.rr == 0) {
        double* dst = sums + (size_t)(blockIdx.x % PB_STAT_REPLICAS) * 2 * C;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            atomicAdd(&dst[L.c0 + i], (double)s[i]);
            atomicAdd(&dst[C + L.c0 + i], (double)s[8 + i]);
        }
    }
}

__global__ void bn_bwd_finalize_kernel(const double* __restrict__ sums, double M, int training,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta,
                                       float* __restrict__ coef, int C) {
    pdl_trigger();
    pdl_wait();
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double s0 = 0, s1 = 0;
    for (int r = 0; r < PB_STAT_REPLICAS; ++r) { s0 += sums[r * 2 * C + c]; s1 += sums[r * 2 * C + C + c]; }
    if (dbeta) dbeta[c] = (float)s0;
    if (dgamma) dgamma[c] = (float)s1;
    coef[c] = training ? (float)(s0 / M) : 0.f;
    coef[C + c] = training ? (float)(s1 / M) : 0.f;
}

// pass 2: dz = scale * (du - coef0 - xhat*coef1)
template <typename T, int ACT, bool BCAST>
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const void* __restrict__ dout, const T* __restrict__ z, const float* __restrict__ scale,
                    const float* __restrict__ shift, const float* __restrict__ mean,
                    const float* __restrict__ invstd, const float* __restrict__ mask,
                    const float* __restrict__ coef, T* __restrict__ dz, unsigned M, unsigned R, int C, float slope) {
    pdl_trigger();
    pdl_wait();
    const RowLayout L(C);
    if (!L.active) return;
    float sc[8], sh[8], a[8], bb[8], k0[8], k1[8];
    load_vec8(scale + L.c0, sc);
    load_vec8(shift + L.c0, sh);
    load_vec8(invstd + L.c0, a);
    load_vec8(mean + L.c0, bb);
    load_vec8(coef + L.c0, k0);
    load_vec8(coef + C + L.c0, k1);
#pragma unroll
    for (int i = 0; i < 8; ++i) bb[i] = -bb[i] * a[i];
    const unsigned stride = gridDim.x * L.RPI;
    for (unsigned m = blockIdx.x * L.RPI + L.rr; m < M; m += stride) {
        float du[8], xh[8];
        bn_bwd_elem<T, ACT, BCAST>(dout, z, mask, m, R, C, L.c0, sc, sh, a, bb, slope, du, xh);
        F8 o;
#pragma unroll
        for (int i = 0; i < 8; ++i) o.v[i] = sc[i] * (du[i] - k0[i] - xh[i] * k1[i]);
        store8(dz + (size_t)m * C + L.c0, o);
    }
}

// ------------------------------------------------------------------------------------------------
// bf16 row streams through a bulk-copy ring (round 2)
//
// The register kernels above keep 32..64 bytes per thread in flight (two rows x 16 bytes per tensor) at ~100
// registers per thread: 32 KB per SM, which buys 2.4 TB/s for the backward reduction (two input tensors) and
// ~3.8-4.1 TB/s for the elementwise passes.  Here a producer thread streams contiguous row chunks (~8 KB per
// tensor) into a 4-stage shared-memory ring with cp.async.bulk, so 64..128 KB per SM are in flight no matter how
// many registers the arithmetic needs; the 256 consumer threads keep the same layout (8 fixed channels per thread,
// row slots strided over the chunk) and read 16-byte pieces from shared memory.  Persistent CTAs, two per SM.
// ------------------------------------------------------------------------------------------------
constexpr int BNB_CONSUMERS = 256;
constexpr int BNB_THREADS = BNB_CONSUMERS + 32;
constexpr int BNB_STAGES = 4;

struct BnBulk {
    unsigned M, R;
    int C;
    int rc;                  // rows per chunk (multiple of the row slots per pass)
    unsigned nchunks;
    unsigned tensor_pitch;   // bytes between the two tensors of a stage (chunk bytes rounded to 128)
    unsigned stage_pitch;
    int stages;              // ring depth (<= BNB_STAGES)
};

struct BnbCtx { uint64_t full[BNB_STAGES], empty[BNB_STAGES]; };

__device__ __forceinline__ void bnb_init(BnbCtx& cx) {
    if (threadIdx.x == 0) {
        for (int s = 0; s < BNB_STAGES; ++s) { tc::mbar_init(&cx.full[s], 1); tc::mbar_init(&cx.empty[s], BNB_CONSUMERS / 32); }
        tc::fence_barrier_init();
    }
    __syncthreads();
}

// producer lane: chunk n of this CTA = rows [n*rc, ...) of up to two tensors (src1 may be null)
__device__ __forceinline__ void bnb_produce(BnbCtx& cx, const BnBulk& p, uint32_t ring, const __nv_bfloat16* src0,
                                            const __nv_bfloat16* src1) {
    int s = 0; uint32_t ph = 1;
    for (unsigned n = blockIdx.x; n < p.nchunks; n += gridDim.x) {
        tc::mbar_wait_parked(&cx.empty[s], ph);
        const unsigned row0 = n * (unsigned)p.rc;
        const unsigned rows = min((unsigned)p.rc, p.M - row0);
        const uint32_t bytes = rows * (uint32_t)p.C * 2u;
        const uint32_t dst = ring + (uint32_t)s * p.stage_pitch;
        tc::mbar_expect_tx(&cx.full[s], src1 ? 2u * bytes : bytes);
        tc::bulk_load_1d(dst, src0 + (size_t)row0 * p.C, bytes, &cx.full[s]);
        if (src1) tc::bulk_load_1d(dst + p.tensor_pitch, src1 + (size_t)row0 * p.C, bytes, &cx.full[s]);
        if (++s == p.stages) { s = 0; ph ^= 1u; }
    }
}

__device__ __forceinline__ F8 lds8_bf16(uint32_t addr) {
    uint4 u;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(addr));
    F8 r;
    r.v[0] = bf16_lo(u.x); r.v[1] = bf16_hi(u.x); r.v[2] = bf16_lo(u.y); r.v[3] = bf16_hi(u.y);
    r.v[4] = bf16_lo(u.z); r.v[5] = bf16_hi(u.z); r.v[6] = bf16_lo(u.w); r.v[7] = bf16_hi(u.w);
    return r;
}

// consumer-side row layout (256 consumer threads, whatever blockDim is)
struct BnbLayout {
    int G, RPI, g, rr, c0;
    bool active;
    __device__ __forceinline__ BnbLayout(int C) {
        G = C >> 3; RPI = BNB_CONSUMERS / G;
        g = threadIdx.x % G; rr = threadIdx.x / G; c0 = g << 3;
        active = rr < RPI;
    }
};

// halving tree over the row slots like cta_fold_rows, synchronising only the consumer warps
template <int NACC>
__device__ __forceinline__ void bnb_fold_rows(float (&acc)[NACC], int G, int RPI, int rr, float* sm) {
    const int tid = threadIdx.x;
    for (int n = RPI; n > 1;) {
        const int half = (n + 1) >> 1;
        if (rr >= half && rr < n) {
#pragma unroll
            for (int j = 0; j < NACC; ++j) sm[j * 256 + tid] = acc[j];
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (rr + half < n) {
#pragma unroll
            for (int j = 0; j < NACC; ++j) acc[j] += sm[j * 256 + tid + half * G];
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        n = half;
    }
}

// du / xhat of one row piece already in registers
template <int ACT>
__device__ __forceinline__ void bn_bwd_math(const F8& zv, const F8& dv, const float* mk, const float (&sc)[8],
                                            const float (&sh)[8], const float (&a)[8], const float (&bb)[8], float slope,
                                            float (&du)[8], float (&xh)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float u = fmaf(zv.v[i], sc[i], sh[i]);
        du[i] = dv.v[i] * actg<ACT>(u, slope);
        xh[i] = fmaf(zv.v[i], a[i], bb[i]);
    }
    if (mk) {
        float m8[8];
        load_vec8(mk, m8);
#pragma unroll
        for (int i = 0; i < 8; ++i) du[i] *= m8[i];
    }
}

// MODE 0: backward reduction (sums of du and du*xhat); MODE 1: backward apply (dz); BCAST: dout is fp32 [B][C]
template <int ACT, bool BCAST, int MODE>
__global__ void __launch_bounds__(BNB_THREADS, 2)
bn_bwd_bulk_kernel(const void* __restrict__ dout, const __nv_bfloat16* __restrict__ z, const float* __restrict__ scale,
                   const float* __restrict__ shift, const float* __restrict__ mean, const float* __restrict__ invstd,
                   const float* __restrict__ mask, const float* __restrict__ coef, double* __restrict__ sums,
                   __nv_bfloat16* __restrict__ dz, const BnBulk p, float slope) {
    extern __shared__ uint8_t bnb_smem[];
    __shared__ BnbCtx cx;
    pdl_trigger();
    const uint32_t raw = tc::smem_u32(bnb_smem);
    const uint32_t ring = (raw + 127u) & ~127u;
    float* sm_fold = reinterpret_cast<float*>(bnb_smem + (ring - raw));   // MODE 0: the fold scratch reuses the drained ring
    bnb_init(cx);
    pdl_wait();
    if (threadIdx.x >= BNB_CONSUMERS) {
        if (threadIdx.x == BNB_CONSUMERS)
            bnb_produce(cx, p, ring, z, BCAST ? nullptr : reinterpret_cast<const __nv_bfloat16*>(dout));
        return;
    }
    const BnbLayout L(p.C);
    const int lane = threadIdx.x & 31;
    float s[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) s[i] = 0.f;
    float sc[8], sh[8], a[8], bb[8], k0[8], k1[8];
    if (L.active) {
        load_vec8(scale + L.c0, sc);
        load_vec8(shift + L.c0, sh);
        load_vec8(invstd + L.c0, a);
        load_vec8(mean + L.c0, bb);
#pragma unroll
        for (int i = 0; i < 8; ++i) bb[i] = -bb[i] * a[i];
        if (MODE == 1) { load_vec8(coef + L.c0, k0); load_vec8(coef + p.C + L.c0, k1); }
    }
    int st = 0; uint32_t ph = 0;
    unsigned b_lo = 1, b_hi = 0;                 // rows of the sample whose vectors are cached (empty range at first)
    F8 dvb;
    float mkv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { dvb.v[i] = 0.f; mkv[i] = 1.f; }
    for (unsigned n = blockIdx.x; n < p.nchunks; n += gridDim.x) {
        tc::mbar_wait_parked(&cx.full[st], ph);
        const unsigned row0 = n * (unsigned)p.rc;
        const int rows = (int)min((unsigned)p.rc, p.M - row0);
        const uint32_t zb = ring + (uint32_t)st * p.stage_pitch + (uint32_t)L.c0 * 2u;
        if (L.active) {
            for (int r = L.rr; r < rows; r += L.RPI) {
                const unsigned m = row0 + (unsigned)r;
                const uint32_t off = (uint32_t)r * (uint32_t)p.C * 2u;
                const F8 zv = lds8_bf16(zb + off);
                if ((BCAST || mask) && (m < b_lo || m >= b_hi)) {      // per-sample vectors: reload when the sample changes,
                    const unsigned b = m / p.R;                        // not per row (a division and 2-4 loads each)
                    b_lo = b * p.R; b_hi = b_lo + p.R;
                    if (BCAST) dvb = load8(reinterpret_cast<const float*>(dout) + (size_t)b * p.C + L.c0);
                    if (mask) load_vec8(mask + (size_t)b * p.C + L.c0, mkv);
                }
                F8 dv;
                if (BCAST) dv = dvb;
                else       dv = lds8_bf16(zb + p.tensor_pitch + off);
                float du[8], xh[8];
                bn_bwd_math<ACT>(zv, dv, nullptr, sc, sh, a, bb, slope, du, xh);
                if (mask) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) du[i] *= mkv[i];
                }
                if (MODE == 0) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) { s[i] += du[i]; s[8 + i] = fmaf(du[i], xh[i], s[8 + i]); }
                } else {
                    F8 o;
#pragma unroll
                    for (int i = 0; i < 8; ++i) o.v[i] = sc[i] * (du[i] - k0[i] - xh[i] * k1[i]);
                    store8(dz + (size_t)m * p.C + L.c0, o);
                }
            }
        }
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&cx.empty[st]);
        if (++st == p.stages) { st = 0; ph ^= 1u; }
    }
    if (MODE == 0) {
        asm volatile("bar.sync 1, 256;" ::: "memory");     // every consumer is done with the ring: it becomes the scratch
        bnb_fold_rows<16>(s, L.G, L.RPI, L.rr, sm_fold);
        if (L.active && L.rr == 0) {
            double* dst = sums + (size_t)(blockIdx.x % PB_STAT_REPLICAS) * 2 * p.C;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                atomicAdd(&dst[L.c0 + i], (double)s[i]);
                atomicAdd(&dst[p.C + L.c0 + i], (double)s[8 + i]);
            }
        }
    }
}

// out = act(z*scale + shift) * mask through the same ring
template <int ACT>
__global__ void __launch_bounds__(BNB_THREADS, 2)
bn_act_fwd_bulk_kernel(const __nv_bfloat16* __restrict__ z, const float* __restrict__ scale,
                       const float* __restrict__ shift, const float* __restrict__ mask,
                       __nv_bfloat16* __restrict__ out, const BnBulk p, float slope) {
    extern __shared__ uint8_t bnb_smem[];
    __shared__ BnbCtx cx;
    pdl_trigger();
    const uint32_t raw = tc::smem_u32(bnb_smem);
    const uint32_t ring = (raw + 127u) & ~127u;
    bnb_init(cx);
    pdl_wait();
    if (threadIdx.x >= BNB_CONSUMERS) {
        if (threadIdx.x == BNB_CONSUMERS) bnb_produce(cx, p, ring, z, nullptr);
        return;
    }
    const BnbLayout L(p.C);
    const int lane = threadIdx.x & 31;
    float sc[8], sh[8];
    if (L.active) { load_vec8(scale + L.c0, sc); load_vec8(shift + L.c0, sh); }
    unsigned b_lo = 1, b_hi = 0;
    float mkv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) mkv[i] = 1.f;
    int st = 0; uint32_t ph = 0;
    for (unsigned n = blockIdx.x; n < p.nchunks; n += gridDim.x) {
        tc::mbar_wait_parked(&cx.full[st], ph);
        const unsigned row0 = n * (unsigned)p.rc;
        const int rows = (int)min((unsigned)p.rc, p.M - row0);
        const uint32_t zb = ring + (uint32_t)st * p.stage_pitch + (uint32_t)L.c0 * 2u;
        if (L.active) {
            for (int r = L.rr; r < rows; r += L.RPI) {
                const unsigned m = row0 + (unsigned)r;
                F8 v = lds8_bf16(zb + (uint32_t)r * (uint32_t)p.C * 2u);
#pragma unroll
                for (int i = 0; i < 8; ++i) v.v[i] = actf<ACT>(fmaf(v.v[i], sc[i], sh[i]), slope);
                if (mask) {
                    if (m < b_lo || m >= b_hi) {                       // reload the sample's mask when the sample changes
                        const unsigned b = m / p.R;
                        b_lo = b * p.R; b_hi = b_lo + p.R;
                        load_vec8(mask + (size_t)b * p.C + L.c0, mkv);
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) v.v[i] *= mkv[i];
                }
                store8(out + (size_t)m * p.C + L.c0, v);
            }
        }
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&cx.empty[st]);
        if (++st == p.stages) { st = 0; ph ^= 1u; }
    }
}

// batch statistics (sums of z and z^2) through the same ring
__global__ void __launch_bounds__(BNB_THREADS, 2)
colstats_bulk_kernel(const __nv_bfloat16* __restrict__ z, double* __restrict__ sums, const BnBulk p) {
    extern __shared__ uint8_t bnb_smem[];
    __shared__ BnbCtx cx;
    pdl_trigger();
    const uint32_t raw = tc::smem_u32(bnb_smem);
    const uint32_t ring = (raw + 127u) & ~127u;
    float* sm_fold = reinterpret_cast<float*>(bnb_smem + (ring - raw));
    bnb_init(cx);
    pdl_wait();
    if (threadIdx.x >= BNB_CONSUMERS) {
        if (threadIdx.x == BNB_CONSUMERS) bnb_produce(cx, p, ring, z, nullptr);
        return;
    }
    const BnbLayout L(p.C);
    const int lane = threadIdx.x & 31;
    float s[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) s[i] = 0.f;
    int st = 0; uint32_t ph = 0;
    for (unsigned n = blockIdx.x; n < p.nchunks; n += gridDim.x) {
        tc::mbar_wait_parked(&cx.full[st], ph);
        const int rows = (int)min((unsigned)p.rc, p.M - n * (unsigned)p.rc);
        const uint32_t zb = ring + (uint32_t)st * p.stage_pitch + (uint32_t)L.c0 * 2u;
        if (L.active) {
            for (int r = L.rr; r < rows; r += L.RPI) {
                const F8 v = lds8_bf16(zb + (uint32_t)r * (uint32_t)p.C * 2u);
#pragma unroll
                for (int i = 0; i < 8; ++i) { s[i] += v.v[i]; s[8 + i] = fmaf(v.v[i], v.v[i], s[8 + i]); }
            }
        }
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&cx.empty[st]);
        if (++st == p.stages) { st = 0; ph ^= 1u; }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    bnb_fold_rows<16>(s, L.G, L.RPI, L.rr, sm_fold);
    if (L.active && L.rr == 0) {
        double* dst = sums + (size_t)(blockIdx.x % PB_STAT_REPLICAS) * 2 * p.C;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            atomicAdd(&dst[L.c0 + i], (double)s[i]);
            atomicAdd(&dst[p.C + L.c0 + i], (double)s[8 + i]);
        }
    }
}

// plan: ~8 KB of rows per tensor and stage; false = use the register kernels (tiny or misaligned problems)
static inline bool bnb_plan(long long M, long long R, int C, int tensors, const void* p0, const void* p1, BnBulk* b,
                            int* grid, size_t* smem, bool fold_scratch) {
    if (getenv("PB_BN_REGISTER_KERNELS")) return false;
    if ((reinterpret_cast<uintptr_t>(p0) & 15) || (p1 && (reinterpret_cast<uintptr_t>(p1) & 15))) return false;
    const int G = C >> 3, RPI = BNB_CONSUMERS / G;
    if (RPI < 1 || M < 4LL * RPI) return false;
    // ~16 KB of rows per tensor and stage, at least 4 rows per thread and chunk (a chunk costs a barrier round trip:
    // with 2 rows per thread the 960-channel layers ran at 1.6 TB/s)
    // measured with tools/bn_bench.py (sum over MobileNetLarge3D's 13 BN shapes, L2 flushed): 16 KB chunks, three stages
    // for two tensors / four for one (fwd 466, reduce 571, apply 616 us against 500 / 718 / 791 us for the register
    // kernels; 8 KB x 4 stages: 498 / 613 / 623 us).  The reduction's 16 KB fold scratch reuses the drained ring.
    const int chunk_kb = 16, min_rows = 4;
    int rc = std::max(min_rows * RPI, chunk_kb * 1024 / (C * 2));
    rc = std::max(RPI, rc / RPI * RPI);
    while ((long long)rc * C * 2 > (chunk_kb + 4) * 1024 && rc > RPI) rc -= RPI;
    b->M = (unsigned)M; b->R = (unsigned)R; b->C = C; b->rc = rc;
    b->nchunks = (unsigned)((M + rc - 1) / rc);
    b->tensor_pitch = ((unsigned)rc * C * 2 + 127u) & ~127u;
    b->stage_pitch = b->tensor_pitch * tensors;
    b->stages = tensors == 2 ? 3 : BNB_STAGES;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    *grid = (int)std::min<long long>(b->nchunks, 2LL * sms);
    *smem = 128 + std::max((size_t)b->stages * b->stage_pitch, fold_scratch ? 16 * 256 * sizeof(float) : (size_t)0);
    return *smem <= 112 * 1024;
}

template <int ACT, bool BCAST, int MODE>
static cudaError_t bnb_launch_bwd(int grid, size_t smem, cudaStream_t st, const void* dout, const void* z, const float* scale,
                                  const float* shift, const float* mean, const float* invstd, const float* mask,
                                  const float* coef, double* sums, void* dz, const BnBulk& b, float slope) {
    static unsigned long long done = 0;
    cudaError_t e = ensure_dyn_smem(bn_bwd_bulk_kernel<ACT, BCAST, MODE>, 112 * 1024, &done);
    if (e != cudaSuccess) return e;
    return launch_pdl(bn_bwd_bulk_kernel<ACT, BCAST, MODE>, dim3(grid), dim3(BNB_THREADS), smem, st, dout,
                      (const __nv_bfloat16*)z, scale, shift, mean, invstd, mask, coef, sums, (__nv_bfloat16*)dz, b, slope);
}
template <int ACT>
static cudaError_t bnb_launch_fwd(int grid, size_t smem, cudaStream_t st, const void* z, const float* scale,
                                  const float* shift, const float* mask, void* out, const BnBulk& b, float slope) {
    static unsigned long long done = 0;
    cudaError_t e = ensure_dyn_smem(bn_act_fwd_bulk_kernel<ACT>, 112 * 1024, &done);
    if (e != cudaSuccess) return e;
    return launch_pdl(bn_act_fwd_bulk_kernel<ACT>, dim3(grid), dim3(BNB_THREADS), smem, st, (const __nv_bfloat16*)z, scale,
                      shift, mask, (__nv_bfloat16*)out, b, slope);
}

// rows_per_thread: 8 for the reductions (few atomics per CTA), 2 for the pure streaming kernels (small tensors
// then still fill the machine)
static inline int row_grid(long long M, int C, int rows_per_thread = 8) {
    int G = C >> 3, RPI = 256 / G;
    long long per_cta = (long long)RPI * rows_per_thread;
    long long max_ctas = (M + per_cta - 1) / per_cta;
    long long want = 148LL * 8;
    return (int)std::max<long long>(1, std::min(max_ctas, want));
}

}  // namespace pb

using namespace pb;

// Dispatch on the activation code; body sees the compile-time constant ACT.
#define PB_DISPATCH_ACT(act, ...)                                                      \
    do {                                                                               \
        switch (act) {                                                                 \
            case PB_ACT_NONE:     { constexpr int ACT = PB_ACT_NONE; __VA_ARGS__; } break;     \
            case PB_ACT_RELU:     { constexpr int ACT = PB_ACT_RELU; __VA_ARGS__; } break;     \
            case PB_ACT_HSWISH:   { constexpr int ACT = PB_ACT_HSWISH; __VA_ARGS__; } break;   \
            case PB_ACT_LRELU:    { constexpr int ACT = PB_ACT_LRELU; __VA_ARGS__; } break;    \
            case PB_ACT_HSIGMOID: { constexpr int ACT = PB_ACT_HSIGMOID; __VA_ARGS__; } break; \
            default: pb::set_error("unknown activation %d", (int)(act)); return PB_ERR_BAD_ARG; \
        }                                                                              \
    } while (0)

extern "C" int pb_colstats(const void* x, int dtype, long long M, int C, double* sums, pb_stream_t stream) {
    PB_REQUIRE(x && sums && M > 0 && C > 0 && C % 8 == 0 && C <= 2048, "colstats: bad args (M=%lld C=%d)", M, C);
    cudaStream_t st = (cudaStream_t)stream;
    PB_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * PB_STAT_REPLICAS * 2 * C, st));
    PB_REQUIRE(M < (1LL << 31), "colstats: too many rows");
    {
        BnBulk bb; int bgrid = 0; size_t bsmem = 0;
        if (dtype == PB_BF16 && bnb_plan(M, M, C, 1, x, nullptr, &bb, &bgrid, &bsmem, true)) {
            static unsigned long long done = 0;
            PB_CUDA(ensure_dyn_smem(colstats_bulk_kernel, 112 * 1024, &done));
            PB_CUDA(launch_pdl(colstats_bulk_kernel, dim3(bgrid), dim3(BNB_THREADS), bsmem, st, (const __nv_bfloat16*)x, sums, bb));
            PB_CHECK_LAUNCH("colstats_bulk");
            return PB_OK;
        }
    }
    const int grid = row_grid(M, C);
    PB_DISPATCH_DTYPE(dtype, {
        (void)launch_pdl(colstats_kernel<T>, dim3(grid), dim3(256), 0, st, (const T*)x, sums, (unsigned)M, C);
    });
    PB_CHECK_LAUNCH("colstats");
    return PB_OK;
}

extern "C" int pb_bn_finalize(const double* sums, long long M, const float* gamma, const float* beta,
                              float* running_mean, float* running_var, int training, float momentum, float eps,
                              float* scale, float* shift, float* mean, float* invstd,
                              long long* num_batches_tracked, int C, pb_stream_t stream) {
    PB_REQUIRE(scale && shift && C > 0, "bn_finalize: bad args");
    PB_REQUIRE(training ? (sums != nullptr && M > 0) : (running_mean && running_var), "bn_finalize: missing statistics");
    const bool cumulative = momentum < 0.f;
    PB_REQUIRE(!cumulative || !training || num_batches_tracked, "bn_finalize: momentum=None needs num_batches_tracked");
    (void)launch_pdl(bn_finalize_kernel, dim3(cumulative ? 1 : ceil_div(C, 128)), dim3(cumulative ? 1024 : 128), 0,
                     (cudaStream_t)stream, sums, (double)M,
                     gamma, beta, running_mean, running_var, training, momentum, eps, scale, shift, mean, invstd,
                     num_batches_tracked, C);
    PB_CHECK_LAUNCH("bn_finalize");
    return PB_OK;
}

extern "C" int pb_bn_act_fwd(const void* z, const float* scale, const float* shift, const float* mask, void* out,
                             int dtype, int B, long long R, int C, int act, float slope, pb_stream_t stream) {
    PB_REQUIRE(z && scale && shift && out && B > 0 && R > 0 && C > 0 && C % 8 == 0 && C <= 2048, "bn_act_fwd: bad args");
    const long long M = (long long)B * R;
    PB_REQUIRE(M < (1LL << 31) && R < (1LL << 31), "bn_act_fwd: too many rows");
    {
        BnBulk bb; int bgrid = 0; size_t bsmem = 0;
        if (dtype == PB_BF16 && bnb_plan(M, R, C, 1, z, out, &bb, &bgrid, &bsmem, false)) {
            PB_DISPATCH_ACT(act, { PB_CUDA(bnb_launch_fwd<ACT>(bgrid, bsmem, (cudaStream_t)stream, z, scale, shift, mask, out, bb, slope)); });
            PB_CHECK_LAUNCH("bn_act_fwd_bulk");
            return PB_OK;
        }
    }
    const int grid = row_grid(M, C, 2);
    PB_DISPATCH_DTYPE(dtype, PB_DISPATCH_ACT(act, {
        (void)launch_pdl(bn_act_fwd_kernel<T, ACT>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const T*)z,
                         scale, shift, mask, (T*)out, (unsigned)M, (unsigned)R, C, slope);
    }));
    PB_CHECK_LAUNCH("bn_act_fwd");
    return PB_OK;
}

extern "C" int pb_bn_act_bwd_reduce(const void* dout, int dout_bcast, const void* z, const float* scale,
                                    const float* shift, const float* mean, const float* invstd, const float* mask,
                                    double* sums, int dtype, int B, long long R, int C, int act, float slope,
                                    pb_stream_t stream) {
    PB_REQUIRE(dout && z && scale && shift && mean && invstd && sums, "bn_act_bwd_reduce: null pointer");
    PB_REQUIRE(B > 0 && R > 0 && C > 0 && C % 8 == 0 && C <= 2048, "bn_act_bwd_reduce: bad dims");
    const long long M = (long long)B * R;
    PB_REQUIRE(M < (1LL << 31) && R < (1LL << 31), "bn_act_bwd_reduce: too many rows");
    cudaStream_t st = (cudaStream_t)stream;
    PB_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * PB_STAT_REPLICAS * 2 * C, st));
    {
        BnBulk bb; int bgrid = 0; size_t bsmem = 0;
        if (dtype == PB_BF16 && bnb_plan(M, R, C, dout_bcast ? 1 : 2, z, dout_bcast ? nullptr : dout, &bb, &bgrid, &bsmem, true)) {
            PB_DISPATCH_ACT(act, {
                if (dout_bcast) PB_CUDA((bnb_launch_bwd<ACT, true, 0>(bgrid, bsmem, st, dout, z, scale, shift, mean, invstd, mask, nullptr, sums, nullptr, bb, slope)));
                else            PB_CUDA((bnb_launch_bwd<ACT, false, 0>(bgrid, bsmem, st, dout, z, scale, shift, mean, invstd, mask, nullptr, sums, nullptr, bb, slope)));
            });
            PB_CHECK_LAUNCH("bn_bwd_reduce_bulk");
            return PB_OK;
        }
    }
    const int grid = row_grid(M, C);
    const size_t smem = 0;
    PB_DISPATCH_DTYPE(dtype, PB_DISPATCH_ACT(act, {
        if (dout_bcast)
            (void)launch_pdl(bn_bwd_reduce_kernel<T, ACT, true>, dim3(grid), dim3(256), smem, st, dout, (const T*)z,
                             scale, shift, mean, invstd, mask, sums, (unsigned)M, (unsigned)R, C, slope);
        else
            (void)launch_pdl(bn_bwd_reduce_kernel<T, ACT, false>, dim3(grid), dim3(256), smem, st, dout, (const T*)z,
                             scale, shift, mean, invstd, mask, sums, (unsigned)M, (unsigned)R, C, slope);
    }));
    PB_CHECK_LAUNCH("bn_act_bwd_reduce");
    return PB_OK;
}

extern "C" int pb_bn_bwd_finalize(const double* sums, long long M, int training, float* dgamma, float* dbeta,
                                  float* coef, int C, pb_stream_t stream) {
    PB_REQUIRE(sums && coef && M > 0 && C > 0, "bn_bwd_finalize: bad args");
    (void)launch_pdl(bn_bwd_finalize_kernel, dim3(ceil_div(C, 128)), dim3(128), 0, (cudaStream_t)stream, sums,
                     (double)M, training, dgamma, dbeta, coef, C);
    PB_CHECK_LAUNCH("bn_bwd_finalize");
    return PB_OK;
}

extern "C" int pb_bn_act_bwd_apply(const void* dout, int dout_bcast, const void* z, const float* scale,
                                   const float* shift, const float* mean, const float* invstd, const float* mask,
                                   const float* coef, void* dz, int dtype, int B, long long R, int C, int act,
                                   float slope, pb_stream_t stream) {
    PB_REQUIRE(dout && z && scale && shift && mean && invstd && coef && dz, "bn_act_bwd_apply: null pointer");
    PB_REQUIRE(B > 0 && R > 0 && C > 0 && C % 8 == 0 && C <= 2048, "bn_act_bwd_apply: bad dims");
    const long long M = (long long)B * R;
    PB_REQUIRE(M < (1LL << 31) && R < (1LL << 31), "bn_act_bwd_apply: too many rows");
    const int grid = row_grid(M, C, 2);
    cudaStream_t st = (cudaStream_t)stream;
    {
        BnBulk bb; int bgrid = 0; size_t bsmem = 0;
        if (dtype == PB_BF16 && (reinterpret_cast<uintptr_t>(dz) & 15) == 0 &&
            bnb_plan(M, R, C, dout_bcast ? 1 : 2, z, dout_bcast ? nullptr : dout, &bb, &bgrid, &bsmem, false)) {
            PB_DISPATCH_ACT(act, {
                if (dout_bcast) PB_CUDA((bnb_launch_bwd<ACT, true, 1>(bgrid, bsmem, st, dout, z, scale, shift, mean, invstd, mask, coef, nullptr, dz, bb, slope)));
                else            PB_CUDA((bnb_launch_bwd<ACT, false, 1>(bgrid, bsmem, st, dout, z, scale, shift, mean, invstd, mask, coef, nullptr, dz, bb, slope)));
            });
            PB_CHECK_LAUNCH("bn_bwd_apply_bulk");
            return PB_OK;
        }
    }
    PB_DISPATCH_DTYPE(dtype, PB_DISPATCH_ACT(act, {
        if (dout_bcast)
            (void)launch_pdl(bn_bwd_apply_kernel<T, ACT, true>, dim3(grid), dim3(256), 0, st, dout, (const T*)z,
                             scale, shift, mean, invstd, mask, coef, (T*)dz, (unsigned)M, (unsigned)R, C, slope);
        else
            (void)launch_pdl(bn_bwd_apply_kernel<T, ACT, false>, dim3(grid), dim3(256), 0, st, dout, (const T*)z,
                             scale, shift, mean, invstd, mask, coef, (T*)dz, (unsigned)M, (unsigned)R, C, slope);
    }));
    PB_CHECK_LAUNCH("bn_act_bwd_apply");
    return PB_OK;
}
