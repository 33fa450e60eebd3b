// Error strings, launch counter and device check for the C ABI (include/picklebot_b200.h).
#include <cstdlib>
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace pb {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};
static std::atomic<long long> g_paths[PB_PATH_COUNT];

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
    set_error("CUDA error in %s: %s", what, cudaGetErrorString(e));
    return PB_ERR_CUDA;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
void count_path(int path) { if (path >= 0 && path < PB_PATH_COUNT) g_paths[path].fetch_add(1, std::memory_order_relaxed); }
bool pdl_enabled() {
    static const bool on = [] { const char* e = getenv("PB_PDL"); return !(e && e[0] == '0'); }();
    return on;
}

}  // namespace pb

extern "C" {

int pb_abi_version(void) { return PB_ABI_VERSION; }

const char* pb_last_error_string(void) { return pb::g_err; }

long long pb_launch_count(void) { return pb::g_launches.load(std::memory_order_relaxed); }

long long pb_path_count(int path) {
    return (path >= 0 && path < PB_PATH_COUNT) ? pb::g_paths[path].load(std::memory_order_relaxed) : -1;
}
void pb_path_reset(void) {
    for (int i = 0; i < PB_PATH_COUNT; ++i) pb::g_paths[i].store(0, std::memory_order_relaxed);
}

int pb_device_check(void) {
    int dev = 0;
    PB_CUDA(cudaGetDevice(&dev));
    int major = 0, minor = 0;
    PB_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    PB_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    if (major != 10) {
        pb::set_error("picklebot_b200 kernels are built for sm_100a only; device is sm_%d%d", major, minor);
        return PB_ERR_UNSUPPORTED;
    }
    return PB_OK;
}

}  // extern "C"
