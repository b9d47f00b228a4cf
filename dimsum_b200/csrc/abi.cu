// Error string, launch accounting and ABI version of libdimsum_b200.so.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace dimsum {

static thread_local char g_error[512] = "";
static std::atomic<int64_t> g_launches{0};

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
    return code;
}

int check_launch(const char *what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(DIMSUM_ERR_CUDA, "%s: CUDA launch failed: %s", what, cudaGetErrorString(e));
    return DIMSUM_OK;
}

}  // namespace dimsum

extern "C" int dimsum_abi_version(void) { return DIMSUM_ABI_VERSION; }
extern "C" const char *dimsum_last_error(void) { return dimsum::g_error; }
extern "C" int64_t dimsum_launch_count(void) { return dimsum::g_launches.load(std::memory_order_relaxed); }
