// Error reporting, launch bookkeeping, version.
#include "common.cuh"
#include "../../include/cwt_b200_debug.h"
#include <cstdarg>
#include <cstdio>

namespace cwt {

void count_launch(int n);
int check_cuda(cudaError_t e, const char* what);
static thread_local char g_err[512] = "";
static thread_local long long g_launches = 0;     // per calling thread, like the error string

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches += n; }

int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return CWT_OK;
    set_error("CUDA error in %s: %s", what, cudaGetErrorString(e));
    return CWT_ERR_CUDA;
}

}  // namespace cwt

namespace cwt {
// L2 read-bandwidth microbenchmark (SURVEY.md §8d asks for one): every CTA sweeps the whole buffer `iters` times with
// 128-bit L1-bypassing loads; with a buffer that fits the 126 MB L2 the steady state is served by L2 only.
__global__ void __launch_bounds__(512) k_l2_read(const float4* __restrict__ buf, size_t n_vec, int iters, float* sink) {
    float acc = 0.f;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int it = 0; it < iters; ++it)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
            float4 v;
            asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(buf + i));
            acc += v.x + v.y + v.z + v.w;
        }
    if (acc == 123.456f) *sink = acc;      // keep the loads alive
}
}  // namespace cwt

extern "C" int cwt_debug_l2_read(const void* buf, size_t bytes, int iters, int ctas, void* sink, void* stream) {
    cwt::k_l2_read<<<ctas, 512, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const float4*>(buf), bytes / 16, iters,
                                                                         static_cast<float*>(sink));
    cwt::count_launch();
    return cwt::check_cuda(cudaGetLastError(), "l2_read");
}

extern "C" int cwt_version(void) { return 100; }
extern "C" const char* cwt_last_error(void) { return cwt::g_err; }
extern "C" long long cwt_launch_count(void) { return cwt::g_launches; }
