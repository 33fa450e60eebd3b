// Interface between stem.cu (entry points, direct kernels) and stem_tc.cu (tcgen05 implicit-im2col kernels).
#pragma once
#include <cuda_runtime.h>

namespace pb {

struct StemTc {
    int B, T, H, W, To, Ho, Wo;
    int sT, sH, sW, pT, pH, pW;
    long long xs_b, xs_t, xs_h;       // element strides of the clip (channel stride 1, pixel stride 3)
    long long P;                      // output pixels
    long long steps;                  // 256-pixel steps
    float inv_scale;                  // uint8 clips: 1 / in_scale (train.py:106's /255)
    int act;                          // activation applied by the forward epilogue (PB_ACT_*; inference with BN folded)
    float slope;
};

// Return true if they launched; false = not covered (caller uses the direct kernels).
bool stem_tc_fwd(const void* x, int x_dtype, const float* w, const float* bias, void* y, int kT, const StemTc& d,
                 cudaStream_t st);
bool stem_tc_wgrad(const void* x, int x_dtype, const void* dy, float* dw, float* dbias, int kT, const StemTc& d,
                   cudaStream_t st);

}  // namespace pb
