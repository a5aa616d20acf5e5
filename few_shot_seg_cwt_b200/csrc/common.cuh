// Shared helpers for the cwt_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/cwt_b200.h"

namespace cwt {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int  check_cuda(cudaError_t e, const char* what);

#define CWT_CUDA(expr)                                                          \
    do {                                                                        \
        int _rc = ::cwt::check_cuda((expr), #expr);                             \
        if (_rc != CWT_OK) return _rc;                                          \
    } while (0)

// after every kernel launch: catches bad launch configurations synchronously
#define CWT_LAUNCHED(name)                                                      \
    do {                                                                        \
        ::cwt::count_launch();                                                  \
        int _rc = ::cwt::check_cuda(cudaGetLastError(), name);                  \
        if (_rc != CWT_OK) return _rc;                                          \
    } while (0)

#define CWT_REQUIRE(cond, code, ...)                                            \
    do {                                                                        \
        if (!(cond)) { ::cwt::set_error(__VA_ARGS__); return (code); }          \
    } while (0)

static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// bump allocator over the caller's workspace
struct Carver {
    char* base; size_t off; size_t cap;
    Carver(void* p, size_t bytes) : base(static_cast<char*>(p)), off(0), cap(bytes) {}
    template <typename T> T* take(size_t n) {
        size_t o = align_up(off);
        off = o + n * sizeof(T);
        return reinterpret_cast<T*>(base + o);
    }
    bool ok() const { return off <= cap && (base != nullptr || off == 0); }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// streaming 128-bit load: read-only path, do not allocate in L1 (features are read once per pass)
__device__ __forceinline__ float4 ldg_stream4(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float ldg_stream1(const float* p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}

// label code: 0 / 1 / 2 (ignored) / 3 (invalid)
__device__ __forceinline__ int label_code(long long v, int ignore_index) {
    return v == 0 ? 0 : (v == 1 ? 1 : (v == ignore_index ? 2 : 3));
}
template <bool I64>
__device__ __forceinline__ int load_label_code(const void* lab, size_t i, int ignore_index) {
    if (I64) return label_code(reinterpret_cast<const long long*>(lab)[i], ignore_index);
    return label_code(reinterpret_cast<const uint8_t*>(lab)[i], ignore_index);
}

// Bilinear align_corners=True sample with scale exactly 1/8, in ATen's CPU rounding order
// (UpSampleKernel.cpp Interpolate<>: verified bit-exact against torch 2.11 F.interpolate):
//     row(a) = fma(v[a][b0], w0, v[a][b1]*w1) ;  out = fma(row(a0), h0, row(a1)*h1)
__device__ __forceinline__ float bilerp8(float v00, float v01, float v10, float v11,
                                         float w0, float w1, float h0, float h1) {
    float t0 = __fmaf_rn(v00, w0, __fmul_rn(v01, w1));
    float t1 = __fmaf_rn(v10, w0, __fmul_rn(v11, w1));
    return __fmaf_rn(t0, h0, __fmul_rn(t1, h1));
}

// packed fp32 pairs (sm_100: FFMA2 / FMUL2): each half is an ordinary round-to-nearest fp32 operation
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float a, float b) {
    f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}

}  // namespace cwt
