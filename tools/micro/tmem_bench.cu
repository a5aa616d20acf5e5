// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/micro/tmem_bench tools/micro/tmem_bench.cu ; run on the GPU box.
// Tensor-memory microbenchmark for the resident fit: can a feature tile (204.8 KB) live in TMEM (256 KB per SM) instead of /
// beside shared memory?  Measures (one CTA per SM, clock64 on the SM):
//   1. tcgen05.st -> tcgen05.ld round trip of raw fp32 bit patterns (exactness),
//   2. tcgen05.ld (32x32b.x32) read throughput with 4 / 8 / 16 warps, whole 128 lanes x 400 columns = 204.8 KB per sweep,
//   3. the same while the other warps sweep 204.8 KB of shared memory with LDS.128 (are the two bandwidths additive?),
//   4. a P1-style sweep from TMEM: lane = channel, FMA with a per-lane weight, butterfly reduce-scatter over the 32 lanes.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float4 lds128_v(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void ldtm32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void sttm32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31};"
        :: "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
           "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
           "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
           "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]), "r"(taddr) : "memory");
}
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

constexpr int NCOL = 400;          // 128 lanes x 400 columns x 4 B = 204.8 KB
constexpr int NBLK = NCOL / 32;    // 12 blocks of 32 columns (+ 16 columns left: ignored by the timing loops -> 196.6 KB)

// MODE 0: TMEM sweep only (NW warps).  MODE 1: warps [0, NW) sweep TMEM, warps [NW, 16) sweep shared memory.
// MODE 2: shared-memory sweep only (warps [NW,16)).  MODE 3: P1-style TMEM sweep with FMA + butterfly reduce-scatter.
// DEPTH: tcgen05.ld instructions in flight before a wait.
template <int MODE, int NW, int DEPTH>
__global__ void __launch_bounds__(512, 1) k(long long* clk, float* out, int reps, int* bad) {
    extern __shared__ __align__(128) float sm[];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s32(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < 512 * 100; i += 512) sm[i] = 1.0f + (i & 7);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = tmem_slot;
    const uint32_t tq = tbase + ((uint32_t)((warp & 3) * 32) << 16);       // this warp's lane quarter
    // ---- fill + round trip: warps 0..3 write every column of their lanes with a distinctive bit pattern ----
    if (warp < 4) {
        for (int b = 0; b < 16; ++b) {
            uint32_t r[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = 0x3f800000u ^ (uint32_t)(((warp * 32 + lane) * 512 + b * 32 + i) * 2654435761u >> 9);
            sttm32(tq + b * 32, r);
        }
        wait_st();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {   // every warp checks its quarter (warps 4..15 re-read what warps 0..3 wrote)
        int nb = 0;
        for (int b = 0; b < 16; ++b) {
            uint32_t r[32];
            ldtm32(tq + b * 32, r);
            wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i)
                nb += r[i] != (0x3f800000u ^ (uint32_t)((((warp & 3) * 32 + lane) * 512 + b * 32 + i) * 2654435761u >> 9));
        }
        if (nb) atomicAdd(bad, nb);
    }
    __syncthreads();
    float acc = 0.f;
    const long long t0 = clock64();
    if ((MODE == 0 || MODE == 1) && warp < NW) {
        // warps sharing a lane quarter split the column blocks
        constexpr int WPQ = NW / 4 > 0 ? NW / 4 : 1;
        const int first = (warp >> 2) % WPQ;
#pragma unroll 1
        for (int r = 0; r < reps; ++r) {
#pragma unroll 1
            for (int b = first; b < NBLK; b += WPQ * DEPTH) {
                uint32_t v[DEPTH][32];
#pragma unroll
                for (int d = 0; d < DEPTH; ++d)
                    if (b + d * WPQ < NBLK) ldtm32(tq + (b + d * WPQ) * 32, v[d]);
                wait_ld();
#pragma unroll
                for (int d = 0; d < DEPTH; ++d)
                    if (b + d * WPQ < NBLK) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) acc += __uint_as_float(v[d][i]);
                    }
            }
        }
    }
    if ((MODE == 1 || MODE == 2) && warp >= NW) {
        const int nt = (16 - NW) * 32, t = tid - NW * 32;
        const uint32_t base = s32(sm);
#pragma unroll 1
        for (int r = 0; r < reps; ++r) {
#pragma unroll 5
            for (int i = t; i < 12800; i += nt) { float4 f = lds128_v(base + i * 16); acc += f.x + f.y + f.z + f.w; }
        }
    }
    if (MODE == 3 && warp < NW) {
        // lane = channel (128 lanes x 3 "channel chunks" of 128 columns... here: 4 chunks x 100 pixels = 400 columns);
        // the WPQ warps of a lane quarter split the 100 pixels in blocks of 25 (padded to 32 columns in this test: 3 blocks of 32)
        constexpr int WPQ = NW / 4 > 0 ? NW / 4 : 1;
        const int first = (warp >> 2) % WPQ;
        const float wd = 1.0f + lane * 0.01f;
#pragma unroll 1
        for (int r = 0; r < reps; ++r) {
#pragma unroll 1
            for (int pb = first; pb < 3; pb += WPQ) {
                float z[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) z[i] = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint32_t v[32];
                    ldtm32(tq + j * 96 + pb * 32, v);
                    wait_ld();
#pragma unroll
                    for (int i = 0; i < 32; ++i) z[i] = fmaf(wd, __uint_as_float(v[i]), z[i]);
                }
                // butterfly reduce-scatter: 32 values x 32 lanes -> lane i holds the sum of value i
#pragma unroll
                for (int s = 16; s >= 1; s >>= 1) {
                    const bool hi = (lane & s) != 0;
#pragma unroll
                    for (int i = 0; i < s; ++i) {
                        const float send = hi ? z[i] : z[i + s], keep = hi ? z[i + s] : z[i];
                        z[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
                    }
                }
                acc += z[0];
            }
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (tid == 0) clk[blockIdx.x] = t1 - t0;
    if (tid == 511) clk[148 + blockIdx.x] = t1 - t0;
    if (acc == 123.456f) out[0] = acc;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase) : "memory");
}

template <int MODE, int NW, int DEPTH>
static void run(const char* name, double bytes_tmem, double bytes_smem) {
    long long* clk; float* out; int* bad;
    cudaMalloc(&clk, 2 * 148 * 8); cudaMalloc(&out, 4); cudaMalloc(&bad, 4); cudaMemset(bad, 0, 4);
    const int reps = 200;
    auto fn = k<MODE, NW, DEPTH>;
    cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 204800);
    fn<<<148, 512, 204800>>>(clk, out, 2, bad);
    fn<<<148, 512, 204800>>>(clk, out, reps, bad);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[296]; int hb = 0;
    cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost); cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost);
    double s = 0, s2 = 0; for (int i = 0; i < 148; ++i) { s += h[i]; s2 += h[148 + i]; }
    const double c = s / 148 / reps, c2 = s2 / 148 / reps;      // clocks per sweep of thread 0 (TMEM role unless MODE 2) / thread 511 (smem role in MODE 1, 2)
    printf("%-58s %s  clk/sweep t0 %8.1f t511 %8.1f  tmem %6.1f B/clk  smem %6.1f B/clk  round-trip mismatches %d\n", name, cudaGetErrorString(e), c, c2,
           bytes_tmem / c, bytes_smem / c2, hb);
    cudaFree(clk); cudaFree(out); cudaFree(bad);
}

int main() {
    const double T = 128.0 * NBLK * 32 * 4, S = 204800.0;
    run<0, 4, 1>("tmem only, 4 warps, 1 ld in flight", T, 0);
    run<0, 4, 2>("tmem only, 4 warps, 2 ld in flight", T, 0);
    run<0, 4, 4>("tmem only, 4 warps, 4 ld in flight", T, 0);
    run<0, 8, 2>("tmem only, 8 warps, 2 ld in flight", T, 0);
    run<0, 16, 1>("tmem only, 16 warps, 1 ld in flight", T, 0);
    run<0, 16, 3>("tmem only, 16 warps, 3 ld in flight", T, 0);
    run<2, 0, 1>("smem only, 16 warps", 0, S);
    run<2, 8, 1>("smem only, 8 warps", 0, S);
    run<1, 8, 2>("tmem 8 warps (timed) + smem 8 warps", T, S);
    run<1, 4, 4>("tmem 4 warps (timed) + smem 12 warps", T, S);
    run<3, 16, 1>("P1-style from TMEM (FMA + butterfly), 16 warps, 384 col", 128.0 * 384 * 4, 0);
    run<3, 4, 1>("P1-style from TMEM (FMA + butterfly), 4 warps, 384 col", 128.0 * 384 * 4, 0);
    return 0;
}
