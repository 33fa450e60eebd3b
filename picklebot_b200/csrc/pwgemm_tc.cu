// tcgen05 / TMEM / TMA bf16 GEMM for the pointwise convolutions (placeholder until the kernel lands).
#include "common.cuh"

using namespace pb;

extern "C" int pb_pw_gemm_tc(const void*, const void*, int, const float*, const float*, const float*, void*,
                             int, long long, int, int, pb_stream_t) {
    set_error("pw_gemm_tc: not built in this revision");
    return PB_ERR_UNSUPPORTED;
}

extern "C" int pb_pw_wgrad_tc(const void*, const void*, float*, float*, int, long long, int, int, int, pb_stream_t) {
    set_error("pw_wgrad_tc: not built in this revision");
    return PB_ERR_UNSUPPORTED;
}
