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
template <typename T> bool dw_fwd_tiled(const T* x, const float* w_tc, T* y, const DwDims& d, cudaStream_t st, float* pool = nullptr);
template <typename T> bool dw_dgrad_tiled(const T* dy, const float* w_tc, T* dx, const DwDims& d, cudaStream_t st);
template <typename T> bool dw_wgrad_tiled(const T* x, const T* dy, float* dw_tc, const DwDims& d, cudaStream_t st);

// Tensor-core (mma.sync) stride-1 kernels of dwconv_mma.cu; same contract (false = not handled).
bool dw_fwd_mma(const __nv_bfloat16* x, const float* w_tc, __nv_bfloat16* y, const DwDims& d, cudaStream_t st);
bool dw_dgrad_mma(const __nv_bfloat16* dy, const float* w_tc, __nv_bfloat16* dx, const DwDims& d, cudaStream_t st);

// TMA-tiled causal streaming forward for the (kT,3,3) MoViNet layers (dwconv_tiled.cu); false = not handled.
bool dw_stream_fwd_tiled(const __nv_bfloat16* x, const __nv_bfloat16* sbuf, const float* w_tc, __nv_bfloat16* y,
                         const DwDims& d, cudaStream_t st);

}  // namespace pb
