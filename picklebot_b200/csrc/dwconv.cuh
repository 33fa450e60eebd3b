// Shared declarations for the depthwise kernels (dwconv.cu = general shapes, dwconv_tiled.cu = fast paths).
#pragma once
#include "common.cuh"

namespace pb {

struct DwDims {
    int B, C, T, H, W;
    int kT, kH, kW, sT, sH, sW, pT, pH, pW;
    int To, Ho, Wo;
};

// Fast paths; each returns true if it handled (launched) the problem, false to fall through to the
// general-shape kernel of the same library.
template <typename T> bool dw_fwd_tiled(const T* x, const float* w_tc, T* y, const DwDims& d, cudaStream_t st);
template <typename T> bool dw_dgrad_tiled(const T* dy, const float* w_tc, T* dx, const DwDims& d, cudaStream_t st);
template <typename T> bool dw_wgrad_tiled(const T* x, const T* dy, float* dw_tc, const DwDims& d, cudaStream_t st);

}  // namespace pb
