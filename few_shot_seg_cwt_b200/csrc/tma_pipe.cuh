// Shared pieces of the TMA-fed shared-memory ring kernels (iou_stream.cuh, skinny_stream.cuh): mbarrier and TMA tile-copy
// wrappers, bounded waits, the tensor-map encoder entry point.
#pragma once
#include "common.cuh"
#include <cuda.h>
#include <cstdlib>

namespace cwt {

constexpr unsigned LS_SPIN = 1u << 22;      // try_wait suspends for a while itself: this is seconds, not milliseconds

__device__ __forceinline__ uint32_t ls_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ls_mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ls_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void ls_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ls_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ls_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ls_u32(bar)) : "memory");
}
// bounded wait: a protocol bug traps (launch error) instead of hanging the GPU
__device__ __forceinline__ void ls_wait(uint64_t* bar, unsigned parity) {
    unsigned ok = 0, it = 0;
    for (;;) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(ls_u32(bar)), "r"(parity) : "memory");
        if (ok) return;
        if (++it > LS_SPIN) __trap();
    }
}
__device__ __forceinline__ void ls_bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(ls_u32(dst)), "l"(src), "r"(bytes), "r"(ls_u32(bar)) : "memory");
}
// one TMA tile copy: box (ns * w pixels) x (LS_CHT channels) of the [E*C][h*w] feature matrix (SASS: UTMALDG.2D)
__device__ __forceinline__ void ls_tma_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(ls_u32(dst)), "l"(map), "r"(ls_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// 128-bit shared-memory load as two packed fp32 pairs, from a 32-bit shared address
__device__ __forceinline__ ulonglong2 ls_lds128(uint32_t saddr) {
    ulonglong2 v;
    asm volatile("ld.shared.v2.b64 {%0,%1}, [%2];" : "=l"(v.x), "=l"(v.y) : "r"(saddr));
    return v;
}       // m[ns - 1]: box of ns low-res rows

typedef CUresult (*LsEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static LsEncodeFn ls_encode_fn() {
    static LsEncodeFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<LsEncodeFn>(ptr);
    }
    return fn;
}

// L2 promotion of the feature tensor maps; CWT_TMA_L2=0|64|128|256 overrides (measured DRAM reads of the logits kernel at
// E = 64: 545 MB with 256 B, 526 MB with 128 B or none, 513 MB with 64 B — the 960-byte tile rows start at 240-byte multiples)
static inline CUtensorMapL2promotion ls_l2_promotion() {
    const char* s = getenv("CWT_TMA_L2");
    const int v = s ? atoi(s) : 64;       // measured: 64 B gives the fewest DRAM bytes (513 vs 545 MB for the logits kernel)
    return v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : v == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
         : v == 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
}

// fp32 matrix [rows][cols] row-major (cols * 4 a multiple of 16 B), box [box_rows][box_cols <= 256], no swizzle;
// elements outside the matrix read as zero and count towards the transaction bytes
static inline int ls_make_map_f32(CUtensorMap* m, const float* base, uint64_t rows, uint64_t cols, uint32_t box_rows,
                                  uint32_t box_cols, const char* what) {
    LsEncodeFn enc = ls_encode_fn();
    CWT_REQUIRE(enc, CWT_ERR_CUDA, "%s: cuTensorMapEncodeTiled is not available from this driver", what);
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * 4};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, ls_l2_promotion(),
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CWT_REQUIRE(r == CUDA_SUCCESS, CWT_ERR_CUDA, "%s: cuTensorMapEncodeTiled failed (%d)", what, (int)r);
    return CWT_OK;
}

}  // namespace cwt
