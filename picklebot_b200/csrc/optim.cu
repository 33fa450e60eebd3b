// Multi-tensor AdamW: one launch updates every parameter of the model (~170 small tensors for MobileNetLarge3D).
// Replaces the optimiser step that follows backward in the reference's loop (train.py:208-212, 283-289; the
// reference uses bitsandbytes' AdamW8bit, which is not available here -- this is plain fp32-state AdamW with
// torch.optim.AdamW's arithmetic: decoupled weight decay, bias-corrected moments).
//
// The host passes device tables: for tensor i the addresses of parameter, gradient, exp_avg, exp_avg_sq and its
// element count; and a chunk list (tensor index, first element) so that every CTA owns one chunk of one tensor.
#include "common.cuh"

namespace pb {

constexpr int ADAMW_CHUNK = 4096;      // elements per CTA (256 threads x 4 float4)

__global__ void __launch_bounds__(256)
adamw_kernel(const long long* __restrict__ ptrs, const long long* __restrict__ sizes,
             const int* __restrict__ chunk_tensor, const long long* __restrict__ chunk_start, int n_tensors,
             float lr, float beta1, float beta2, float eps, float weight_decay, float bc1, float bc2_sqrt,
             float grad_scale) {
    pdl_trigger();
    pdl_wait();
    const int t = chunk_tensor[blockIdx.x];
    const long long start = chunk_start[blockIdx.x];
    const long long n = sizes[t];
    float* p = reinterpret_cast<float*>(ptrs[t]);
    const float* g = reinterpret_cast<const float*>(ptrs[n_tensors + t]);
    float* m = reinterpret_cast<float*>(ptrs[2 * n_tensors + t]);
    float* v = reinterpret_cast<float*>(ptrs[3 * n_tensors + t]);
    const float step_size = lr / bc1;
    const long long end = min(n, start + (long long)ADAMW_CHUNK);
    auto update = [&](float& pv, float gv, float& mv, float& vv) {
        gv *= grad_scale;
        pv *= 1.f - lr * weight_decay;
        mv = mv + (1.f - beta1) * (gv - mv);                 // torch: exp_avg.lerp_(grad, 1 - beta1)
        vv = beta2 * vv + (1.f - beta2) * gv * gv;
        const float denom = sqrtf(vv) / bc2_sqrt + eps;
        pv -= step_size * (mv / denom);
    };
    const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                       reinterpret_cast<uintptr_t>(v)) & 15) == 0;
    if (vec) {
        for (long long i = start + threadIdx.x * 4; i + 3 < end; i += 256 * 4) {
            float4 pv = *reinterpret_cast<float4*>(p + i), mv = *reinterpret_cast<float4*>(m + i);
            float4 vv = *reinterpret_cast<float4*>(v + i);
            const float4 gv = *reinterpret_cast<const float4*>(g + i);
            update(pv.x, gv.x, mv.x, vv.x); update(pv.y, gv.y, mv.y, vv.y);
            update(pv.z, gv.z, mv.z, vv.z); update(pv.w, gv.w, mv.w, vv.w);
            *reinterpret_cast<float4*>(p + i) = pv;
            *reinterpret_cast<float4*>(m + i) = mv;
            *reinterpret_cast<float4*>(v + i) = vv;
        }
        const long long tail = start + ((end - start) & ~3LL);
        for (long long i = tail + threadIdx.x; i < end; i += 256) update(p[i], g[i], m[i], v[i]);
    } else {
        for (long long i = start + threadIdx.x; i < end; i += 256) update(p[i], g[i], m[i], v[i]);
    }
}

}  // namespace pb

using namespace pb;

extern "C" int pb_adamw_chunk_elems(void) { return ADAMW_CHUNK; }

extern "C" int pb_adamw_step(const long long* ptrs, const long long* sizes, const int* chunk_tensor,
                             const long long* chunk_start, int n_tensors, int n_chunks, float lr, float beta1,
                             float beta2, float eps, float weight_decay, float bias_correction1,
                             float bias_correction2_sqrt, float grad_scale, pb_stream_t stream) {
    PB_REQUIRE(ptrs && sizes && chunk_tensor && chunk_start && n_tensors > 0 && n_chunks > 0, "adamw_step: bad args");
    PB_REQUIRE(bias_correction1 > 0.f && bias_correction2_sqrt > 0.f, "adamw_step: bias corrections must be positive");
    (void)launch_pdl(adamw_kernel, dim3(n_chunks), dim3(256), 0, (cudaStream_t)stream, ptrs, sizes, chunk_tensor,
                     chunk_start, n_tensors, lr, beta1, beta2, eps, weight_decay, bias_correction1,
                     bias_correction2_sqrt, grad_scale);
    PB_CHECK_LAUNCH("adamw_kernel");
    return PB_OK;
}
