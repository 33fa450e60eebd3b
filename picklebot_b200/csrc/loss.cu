// Cross-entropy loss, its gradient and the number of correct argmax calls in one launch
// (criterion = nn.CrossEntropyLoss(), train.py:214, 266-267; calculate_accuracy, train.py:110-114).
// One warp per sample: max, log-sum-exp, loss_b = lse - logit[label]; loss = scale * mean_b loss_b;
// dlogits[b][c] = scale / B * (softmax_c - [c == label]); correct += (argmax_c logits[b][c] == label).
// Labels outside [0, NC) (torch's ignore_index = -100 is one) are never used as an index: such samples add no
// loss, no gradient and no correct call, and the mean runs over the remaining samples like
// nn.CrossEntropyLoss(reduction="mean") does for ignored targets.
#include "common.cuh"

namespace pb {

__global__ void __launch_bounds__(256)
ce_loss_kernel(const float* __restrict__ logits, const long long* __restrict__ labels, float* __restrict__ loss,
               float* __restrict__ dlogits, int* __restrict__ correct, int B, int NC, float scale) {
    pdl_trigger();
    pdl_wait();
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= B) return;
    const float* row = logits + (long long)warp * NC;
    const long long label64 = labels[warp];
    const bool valid = label64 >= 0 && label64 < NC;
    const int label = valid ? (int)label64 : -1;
    int nvalid = 0;                                    // every warp counts the usable labels of the batch
    for (int b = lane; b < B; b += 32) { const long long l = labels[b]; nvalid += (l >= 0 && l < NC) ? 1 : 0; }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) nvalid += __shfl_xor_sync(0xffffffffu, nvalid, off);
    float mx = -INFINITY;
    int arg = 0;
    for (int c = lane; c < NC; c += 32) {
        const float v = row[c];
        if (v > mx) { mx = v; arg = c; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {           // max with the lowest index on ties (torch.argmax)
        const float om = __shfl_xor_sync(0xffffffffu, mx, off);
        const int oa = __shfl_xor_sync(0xffffffffu, arg, off);
        if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
    }
    float se = 0.f;
    for (int c = lane; c < NC; c += 32) se += expf(row[c] - mx);
    se = warp_sum(se);
    const float lse = mx + logf(se);
    const float inv = valid ? scale / (float)nvalid : 0.f;
    if (dlogits) {
        float* drow = dlogits + (long long)warp * NC;
        for (int c = lane; c < NC; c += 32) drow[c] = inv * (expf(row[c] - lse) - (c == label ? 1.f : 0.f));
    }
    if (lane == 0 && valid) {
        atomicAdd(loss, inv * (lse - row[label]));
        if (correct && arg == label) atomicAdd(correct, 1);
    }
}

}  // namespace pb

using namespace pb;

extern "C" int pb_ce_loss(const float* logits, const long long* labels, float* loss, float* dlogits, int* correct,
                          int B, int NC, float scale, pb_stream_t stream) {
    PB_REQUIRE(logits && labels && loss && B > 0 && NC > 0, "ce_loss: bad args");
    cudaStream_t st = (cudaStream_t)stream;
    PB_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), st));
    if (correct) PB_CUDA(cudaMemsetAsync(correct, 0, sizeof(int), st));
    (void)launch_pdl(ce_loss_kernel, dim3(ceil_div((long long)B * 32, 256)), dim3(256), 0, st, logits, labels, loss, dlogits,
                     correct, B, NC, scale);
    PB_CHECK_LAUNCH("ce_loss_kernel");
    return PB_OK;
}
