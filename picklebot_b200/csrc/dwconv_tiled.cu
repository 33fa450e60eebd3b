// Tiled fast paths for the depthwise kernels (filled in after the general kernels are parity-green).
#include "dwconv.cuh"

namespace pb {

template <typename T> bool dw_fwd_tiled(const T*, const float*, T*, const DwDims&, cudaStream_t) { return false; }
template <typename T> bool dw_dgrad_tiled(const T*, const float*, T*, const DwDims&, cudaStream_t) { return false; }
template <typename T> bool dw_wgrad_tiled(const T*, const T*, float*, const DwDims&, cudaStream_t) { return false; }

#define INST(T)                                                                                   \
    template bool dw_fwd_tiled<T>(const T*, const float*, T*, const DwDims&, cudaStream_t);       \
    template bool dw_dgrad_tiled<T>(const T*, const float*, T*, const DwDims&, cudaStream_t);     \
    template bool dw_wgrad_tiled<T>(const T*, const T*, float*, const DwDims&, cudaStream_t);
INST(float)
INST(__nv_bfloat16)

}  // namespace pb
