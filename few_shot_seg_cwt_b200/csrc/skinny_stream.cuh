// Streaming versions of the two skinny contractions for R <= 8 rows (the re-associated attention of
// MultiHeadAttentionOne, src/model/transformer.py:60-76, with Lq * n_head = 8 rows at the script settings):
//
//   FTR  out[e][r][c] = sum_p P[e][r][p] * f[e][c][p]     (a.X, and dS.X in the backward)
//   RTF  out[e][r][p] = sum_c M[e][r][c] * f[e][c][p]     (+ |f[:,p]|^2)   (scores = (Q_h A_h) X^T)
//
// Same structure as the logits kernel (iou_stream.cuh): persistent CTAs, one producer warp issuing TMA tile copies
// (cp.async.bulk.tensor.2d, mbarrier complete_tx) into a shared-memory ring, eight consumer warps working out of shared
// memory with packed fp32 FMAs. The LDG versions (skinny.cuh) reach 2.0 / 3.4 TB/s at R = 8 (128 / 76 registers, 15 - 21
// resident warps, 8 row loads per 2 feature loads); here every feature byte is fetched by the TMA unit and each row value
// read from shared memory feeds four channels.
#pragma once
#include "common.cuh"
#include "tma_pipe.cuh"
#include "skinny.cuh"
#include <cstdlib>
#include <cstring>

namespace cwt {

constexpr int SS_W = 8;                 // consumer warps
constexpr int SS_THREADS = 32 * (1 + SS_W);
constexpr int SS_PX = 256;              // pixels per stage (TMA box width; 64 quads = two per lane)
constexpr int SS_CB = 32;               // channels per work item / per stage (four per consumer warp)
constexpr int SS_NS = 4;                // ring stages

struct SsMaps { CUtensorMap f, p; };
struct SsParams { float* out; int E, C, HW, R; };

// ---------------------------------------------------------------------------------------------------------------------
// FTR: work item = (episode, block of 32 channels); a stage = pixels [256 k, 256 k + 256) of the 32 feature rows and of
// the R probability rows (pixels past HW are zero-filled by the TMA unit and add nothing). Warp w owns channels
// 4 w .. 4 w + 3; a lane owns quads l and l + 32 of the stage and keeps 4 x R packed partial sums.
// ---------------------------------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(SS_THREADS, 1) k_ftr_stream(const __grid_constant__ SsMaps maps, SsParams p) {
    extern __shared__ __align__(128) unsigned char ss_raw[];
    constexpr unsigned F_BYTES = SS_CB * SS_PX * 4, P_BYTES = R * SS_PX * 4, STAGE = F_BYTES + P_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(ss_raw + (size_t)SS_NS * STAGE);
    uint64_t* empty = full + SS_NS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nblk = p.C / SS_CB, n_items = p.E * nblk, nchunk = (p.HW + SS_PX - 1) / SS_PX;
    if (tid == 0) {
        for (int i = 0; i < SS_NS; ++i) { ls_mbar_init(&full[i], 1); ls_mbar_init(&empty[i], SS_W); }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    unsigned slot = 0, ph = 0;
    if (warp == 0) {
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const int e = item / nblk, cb = item - e * nblk;
            for (int k = 0; k < nchunk; ++k) {
                ls_wait(&empty[slot], ph ^ 1u);
                if (lane == 0) {
                    unsigned char* st = ss_raw + (size_t)slot * STAGE;
                    ls_expect_tx(&full[slot], STAGE);
                    ls_tma_2d(st, &maps.f, k * SS_PX, e * p.C + cb * SS_CB, &full[slot]);
                    ls_tma_2d(st + F_BYTES, &maps.p, k * SS_PX, e * p.R, &full[slot]);
                }
                __syncwarp();
                if (++slot == SS_NS) { slot = 0; ph ^= 1u; }
            }
        }
        return;
    }
    const int cw = warp - 1;
    const uint32_t base = ls_u32(ss_raw) + (uint32_t)lane * 16u;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int e = item / nblk, cb = item - e * nblk;
        f32x2 acc[4][R];                                   // {even pixels, odd pixels} partial sums of channel j, row r
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int r = 0; r < R; ++r) acc[j][r] = 0ull;
        for (int k = 0; k < nchunk; ++k) {
            ls_wait(&full[slot], ph);
            const uint32_t st = base + slot * STAGE;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const uint32_t fa = st + (uint32_t)(cw * 4) * (SS_PX * 4) + q * 512u, pa = st + F_BYTES + q * 512u;
                ulonglong2 pv[R], fv[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) fv[j] = ls_lds128(fa + j * (SS_PX * 4));
#pragma unroll
                for (int r = 0; r < R; ++r) pv[r] = ls_lds128(pa + r * (SS_PX * 4));
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int r = 0; r < R; ++r) acc[j][r] = fma2(pv[r].y, fv[j].y, fma2(pv[r].x, fv[j].x, acc[j][r]));
            }
            __syncwarp();
            if (lane == 0) ls_arrive(&empty[slot]);
            if (++slot == SS_NS) { slot = 0; ph ^= 1u; }
        }
        // lane partials -> warp sums (fixed shuffle tree: deterministic), lane 0 writes the 4 x R results
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float a, b;
                upk2(acc[j][r], a, b);
                const float v = warp_sum(a + b);
                if (lane == 0 && r < p.R) p.out[((size_t)e * p.R + r) * p.C + cb * SS_CB + cw * 4 + j] = v;
            }
    }
}

// out[e][r][c], r < r_actual <= 8; P [E][r_actual][HW]; f [E][C][HW]. Returns CWT_ERR_UNSUPPORTED (no error text) for shapes
// the tile copies cannot take: the caller falls back to the LDG kernel.
static int launch_ftr_stream(const float* f, const float* P, float* out, int E, int C, int HW, int r_actual, cudaStream_t st) {
    if (r_actual < 1 || r_actual > 8 || C % SS_CB || HW % 4 || HW < SS_PX || (reinterpret_cast<uintptr_t>(f) & 15u) ||
        (reinterpret_cast<uintptr_t>(P) & 15u))
        return CWT_ERR_UNSUPPORTED;
    if ((long long)E * r_actual < 8) return CWT_ERR_UNSUPPORTED;          // the row box must fit inside the matrix
    const int R = r_actual <= 4 ? 4 : 8;                  // box height of the probability tile: rows past E * r_actual are zero-filled,
    SsMaps maps;                                          // rows of the NEXT episode are read and multiplied but never stored
    memset(&maps, 0, sizeof(maps));
    int rc = ls_make_map_f32(&maps.f, f, (uint64_t)E * C, (uint64_t)HW, SS_CB, SS_PX, "feat_times_rows");
    if (rc != CWT_OK) return rc;
    rc = ls_make_map_f32(&maps.p, P, (uint64_t)E * r_actual, (uint64_t)HW, (uint32_t)R, SS_PX, "feat_times_rows");
    if (rc != CWT_OK) return rc;
    SsParams p{out, E, C, HW, r_actual};
    int dev = 0, n_sm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    const int n_items = E * (C / SS_CB);
    const int grid = n_items < n_sm ? n_items : n_sm;
    if (R == 4) {
        const size_t sm = (size_t)SS_NS * (SS_CB * SS_PX * 4 + 4 * SS_PX * 4) + 16 * SS_NS;
        CWT_CUDA(cudaFuncSetAttribute(k_ftr_stream<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        k_ftr_stream<4><<<grid, SS_THREADS, sm, st>>>(maps, p);
    } else {
        const size_t sm = (size_t)SS_NS * (SS_CB * SS_PX * 4 + 8 * SS_PX * 4) + 16 * SS_NS;
        CWT_CUDA(cudaFuncSetAttribute(k_ftr_stream<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        k_ftr_stream<8><<<grid, SS_THREADS, sm, st>>>(maps, p);
    }
    CWT_LAUNCHED("feat_times_rows_stream");
    return CWT_OK;
}

// CWT_SKINNY=ldg forces the LDG kernels of skinny.cuh (tests run both)
static inline bool skinny_stream_enabled() {
    const char* s = getenv("CWT_SKINNY");
    return !(s && strcmp(s, "ldg") == 0);
}

// out[e][r][c] = sum_{s,p} P[e][r][s*HW+p] * f[e*S+s][c][p]: streaming kernel when it applies, LDG kernel otherwise
static int launch_feat_times_cols_auto(const float* f, const float* P, float* out, int E, int S, int C, int HW, int r_actual,
                                       cudaStream_t st) {
    if (S == 1 && skinny_stream_enabled()) {
        const int rc = launch_ftr_stream(f, P, out, E, C, HW, r_actual, st);
        if (rc != CWT_ERR_UNSUPPORTED) return rc;
    }
    return launch_feat_times_cols(f, P, out, E, S, C, HW, r_actual, st);
}

}  // namespace cwt
