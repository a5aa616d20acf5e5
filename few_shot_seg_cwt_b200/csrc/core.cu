// Error reporting, launch bookkeeping, version.
#include "common.cuh"
#include <atomic>
#include <cstdarg>
#include <cstdio>

namespace cwt {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return CWT_OK;
    set_error("CUDA error in %s: %s", what, cudaGetErrorString(e));
    return CWT_ERR_CUDA;
}

}  // namespace cwt

extern "C" int cwt_version(void) { return 100; }
extern "C" const char* cwt_last_error(void) { return cwt::g_err; }
extern "C" long long cwt_launch_count(void) { return cwt::g_launches.load(std::memory_order_relaxed); }
