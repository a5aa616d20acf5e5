// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/micro/fma_bench tools/micro/fma_bench.cu ; run on the GPU box.
// FFMA (3-register) vs FFMA2 (packed fp32 pair, sm_100) issue rate and dependent latency per SM sub-partition:
// one CTA per SM, W warps per sub-partition, CH independent accumulator chains per thread.
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
__device__ __forceinline__ float fma1(float a, float b, float c) {
    float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d;
}

template <int PACKED, int CH>
__global__ void k(float* out, long long* clk, int reps, float a, float b) {
    long long t0 = 0;
    if (PACKED) {
        f32x2 acc[CH];
        f32x2 pa, pb;
        asm("mov.b64 %0, {%1, %2};" : "=l"(pa) : "f"(a), "f"(a + 1.f));
        asm("mov.b64 %0, {%1, %2};" : "=l"(pb) : "f"(b), "f"(b + 1.f));
#pragma unroll
        for (int i = 0; i < CH; ++i) acc[i] = (f32x2)threadIdx.x + i;
        __syncthreads();
        t0 = clock64();
#pragma unroll 1
        for (int r = 0; r < reps; ++r) {
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
                for (int i = 0; i < CH; ++i) acc[i] = fma2(acc[i], pa, pb);
        }
        long long t1 = clock64();
        f32x2 s = 0;
#pragma unroll
        for (int i = 0; i < CH; ++i) s ^= acc[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = (float)(s & 0xffff);
        if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
    } else {
        float acc[CH];
#pragma unroll
        for (int i = 0; i < CH; ++i) acc[i] = threadIdx.x + i;
        __syncthreads();
        t0 = clock64();
#pragma unroll 1
        for (int r = 0; r < reps; ++r) {
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
                for (int i = 0; i < CH; ++i) acc[i] = fma1(acc[i], a, b);
        }
        long long t1 = clock64();
        float s = 0;
#pragma unroll
        for (int i = 0; i < CH; ++i) s += acc[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
        if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
    }
}

template <int PACKED, int CH>
static void run(int warps_per_smsp, float* out, long long* clk) {
    const int reps = 2000;
    k<PACKED, CH><<<148, 128 * warps_per_smsp>>>(out, clk, reps, 0.999f, 0.001f);
    cudaDeviceSynchronize();
    k<PACKED, CH><<<148, 128 * warps_per_smsp>>>(out, clk, reps, 0.999f, 0.001f);
    cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, clk, sizeof(c), cudaMemcpyDeviceToHost);
    const double per_inst = (double)c / ((double)reps * 8 * CH);          // cycles per instruction of ONE warp
    printf("%s chains=%d warps/SMSP=%d: %.2f clk per instr per warp -> %.2f warp-instr/clk/SMSP, %.1f fp32 FMA lanes/clk/SM\n",
           PACKED ? "FFMA2" : "FFMA ", CH, warps_per_smsp, per_inst, warps_per_smsp / per_inst,
           4.0 * 32 * (PACKED ? 2 : 1) * warps_per_smsp / per_inst);
}

int main() {
    float* out; long long* clk;
    cudaMalloc(&out, 148 * 1024 * sizeof(float));
    cudaMalloc(&clk, 148 * sizeof(long long));
    run<0, 1>(1, out, clk); run<1, 1>(1, out, clk);       // dependent latency
    run<0, 4>(1, out, clk); run<1, 4>(1, out, clk);
    run<0, 8>(1, out, clk); run<1, 8>(1, out, clk);       // one warp, plenty of ILP
    run<0, 8>(4, out, clk); run<1, 8>(4, out, clk);       // pipe throughput
    run<0, 2>(4, out, clk); run<1, 2>(4, out, clk);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
