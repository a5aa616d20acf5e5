// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/micro/lds_bench tools/micro/lds_bench.cu ; run on the GPU box.
// Shared-memory LDS.128 access-pattern microbenchmark (one CTA per SM, 512 threads, 204.8 KB of smem).
// Patterns are the candidates for the resident fit's second sweep (rows of 100 floats = 400 B per channel).
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ float4 lds128_v(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}

template <int PAT>
__global__ void __launch_bounds__(512, 1) k(float* out, long long* clk, int reps, int zero, unsigned long long* acc64) {
    extern __shared__ __align__(128) float sm[];
    const int tid = threadIdx.x;
    for (int i = tid; i < 512 * 100; i += 512) sm[i] = 1.0f + (i & 7);
    __syncthreads();
    const uint32_t base0 = (uint32_t)__cvta_generic_to_shared(sm);
    float acc = 0.f;
    long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < reps; ++r) {
        const uint32_t base = base0 + (uint32_t)(r * zero);     // (zero == 0: keeps the loads inside the loop)
        if (PAT == 0) {            // contiguous: thread t reads quad t + 512 i   (25 loads cover the whole array)
#pragma unroll
            for (int i = 0; i < 25; ++i) { float4 f = lds128_v(base + (tid + 512 * i) * 16); acc += f.x + f.y + f.z + f.w; }
        } else if (PAT == 1) {     // 4 lanes / channel, quads {2p,2p+1}+8i: 128 channels per pass, 4 passes, 6 loads each (+ quad 24 skipped)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t row = base + (j * 128 + (tid >> 2)) * 400 + (tid & 3) * 32;
#pragma unroll
                for (int i = 0; i < 3; ++i) { float4 f = lds128_v(row + 128 * i); float4 g = lds128_v(row + 128 * i + 16); acc += f.x + f.y + f.z + f.w + g.x + g.y + g.z + g.w; }
            }
        } else if (PAT == 2) {     // 8 lanes / channel, quads p + 8i: 64 channels per pass, 8 passes, 3 loads each
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t row = base + (j * 64 + (tid >> 3)) * 400 + (tid & 7) * 16;
#pragma unroll
                for (int i = 0; i < 3; ++i) { float4 f = lds128_v(row + 128 * i); acc += f.x + f.y + f.z + f.w; }
            }
        } else if (PAT == 3) {     // thread per channel: 24 quads
            const uint32_t row = base + tid * 400;
#pragma unroll
            for (int i = 0; i < 24; ++i) { float4 f = lds128_v(row + 16 * i); acc += f.x + f.y + f.z + f.w; }
        } else if (PAT >= 6 && PAT <= 11) {   // the real P3: 4 lanes / channel, g in registers, FMAs (6), + shuffles (7), + RED (8), prefetch order (9)
            const int part = tid & 3, cl = tid >> 2;
            float4 gq[7];
#pragma unroll
            for (int i = 0; i < 3; ++i) { gq[2 * i] = lds128_v(base + 51200 * 4 - 512 + (2 * part + 8 * i) * 16); gq[2 * i + 1] = lds128_v(base + 51200 * 4 - 512 + (2 * part + 8 * i + 1) * 16); }
            gq[6] = lds128_v(base + 51200 * 4 - 512 + 24 * 16);
            const uint32_t row0 = base + cl * 400 + part * 32;
            float4 f[7];
            if (PAT == 9) {
#pragma unroll
                for (int i = 0; i < 3; ++i) { f[2 * i] = lds128_v(row0 + 128 * i); f[2 * i + 1] = lds128_v(row0 + 128 * i + 16); }
                f[6] = lds128_v(row0 - part * 32 + 384);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t row = row0 + j * 128 * 400;
                if (PAT != 9) {
#pragma unroll
                    for (int i = 0; i < 3; ++i) { f[2 * i] = lds128_v(row + 128 * i); f[2 * i + 1] = lds128_v(row + 128 * i + 16); }
                    f[6] = lds128_v(row - part * 32 + 384);
                }
                float d0 = 0.f, d1 = 0.f;
#pragma unroll
                for (int i = 0; i < 7; ++i) {
                    d0 = fmaf(gq[i].x, f[i].x, d0); d1 = fmaf(gq[i].y, f[i].y, d1);
                    d0 = fmaf(gq[i].z, f[i].z, d0); d1 = fmaf(gq[i].w, f[i].w, d1);
                }
                if (PAT == 9 && j < 3) {
                    const uint32_t rown = row + 128 * 400;
#pragma unroll
                    for (int i = 0; i < 3; ++i) { f[2 * i] = lds128_v(rown + 128 * i); f[2 * i + 1] = lds128_v(rown + 128 * i + 16); }
                    f[6] = lds128_v(rown - part * 32 + 384);
                }
                float d = d0 + d1;
                if (PAT >= 7) { d += __shfl_xor_sync(0xffffffffu, d, 1); d += __shfl_xor_sync(0xffffffffu, d, 2); }
                if (PAT >= 8) {
                    if (part == 0) {
                        unsigned long long v = ((unsigned long long)__float2ll_rn(d * 1048576.f) << 13) + 1ull;
                        asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(acc64 + (PAT == 10 ? (blockIdx.x / 36) * 512 : PAT == 11 ? 0 : blockIdx.x * 512) + j * 128 + cl), "l"(v));
                    }
                } else acc += d;
            }
        } else if (PAT == 4) {     // 4 lanes / channel, quads 4i + p (64 contiguous bytes per channel per load)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t row = base + (j * 128 + (tid >> 2)) * 400 + (tid & 3) * 16;
#pragma unroll
                for (int i = 0; i < 6; ++i) { float4 f = lds128_v(row + 64 * i); acc += f.x + f.y + f.z + f.w; }
            }
        } else if (PAT == 5) {     // 2 lanes / channel, quads 2i + p: 256 channels per pass, 2 passes, 12 loads each
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const uint32_t row = base + (j * 256 + (tid >> 1)) * 400 + (tid & 1) * 16;
#pragma unroll
                for (int i = 0; i < 12; ++i) { float4 f = lds128_v(row + 32 * i); acc += f.x + f.y + f.z + f.w; }
            }
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * 512 + tid] = acc;
    if (tid == 0) clk[blockIdx.x] = t1 - t0;
}

template <int PAT>
void run(const char* name, int loads_per_thread) {
    float* out; long long* clk; unsigned long long* acc64;
    cudaMalloc(&acc64, 148 * 512 * 8); cudaMemset(acc64, 0, 148 * 512 * 8);
    cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&clk, 148 * 8);
    const int smem = 512 * 100 * 4, reps = 200;
    cudaFuncSetAttribute(k<PAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k<PAT><<<148, 512, smem>>>(out, clk, reps, 0, acc64);
    k<PAT><<<148, 512, smem>>>(out, clk, reps, 0, acc64);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    double mean = 0; for (int i = 0; i < 148; ++i) mean += h[i]; mean /= 148;
    const double per_sweep = mean / reps, bytes = 512.0 * loads_per_thread * 16;
    printf("%-44s %8.0f clk / sweep   %6.1f B/clk   (%d LDS.128 per thread, err=%s)\n", name, per_sweep, bytes / per_sweep,
           loads_per_thread, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(clk);
}

int main() {
    run<0>("0 contiguous (P1 pattern)", 25);
    run<1>("1 4 lanes/channel, quads {2p,2p+1}+8i", 24);
    run<2>("2 8 lanes/channel, quads p+8i", 24);
    run<3>("3 thread per channel", 24);
    run<4>("4 4 lanes/channel, quads 4i+p", 24);
    run<6>("6 P3: pattern 1 + quad 24 + g regs + FMA", 28);
    run<7>("7 P3: 6 + shuffles", 28);
    run<8>("8 P3: 7 + F2I + RED.64", 28);
    run<9>("9 P3: 8 with next-chunk prefetch", 28);
    run<10>("10 P3: 8, groups of 36 CTAs share the words", 28);
    run<11>("11 P3: 8, all 148 CTAs share the words", 28);
    run<5>("5 2 lanes/channel, quads 2i+p", 24);
    return 0;
}
