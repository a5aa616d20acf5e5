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
// (no lower bound on the number of work items: even at E = 1 — 15 / 16 items — the streaming kernels beat the LDG ones,
//  transformer forward + backward + optimizer of the training step 0.83 ms vs 1.12 ms)

// CWT_SKINNY=ldg forces the LDG kernels of skinny.cuh (tests run both)
static inline bool skinny_stream_enabled() {
    const char* s = getenv("CWT_SKINNY");
    return !(s && strcmp(s, "ldg") == 0);
}

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


// ---------------------------------------------------------------------------------------------------------------------
// RTF: work item = (episode, chunk of 256 pixels); a stage = 32 channels of the chunk. Warp w owns the channels
// c = w (mod 8); a lane owns quads l and l + 32 and keeps R rows (+ |f_p|^2) x 2 quads of packed sums; the R weight rows of
// the episode sit in shared memory as {m, m} pairs. After the C channels of an item the eight partials are combined in a
// fixed order and written out. A CTA owns a CONTIGUOUS range of items, so the weights are reloaded about once per CTA.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int RS_NS = 3;
struct RsParams { const float* M; float* out; float* n2; int E, C, HW, R; };

template <int R, bool NORM2>
__global__ void __launch_bounds__(SS_THREADS, 1) k_rtf_stream(const __grid_constant__ CUtensorMap fmap, RsParams p) {
    extern __shared__ __align__(128) unsigned char ss_raw[];
    constexpr unsigned STAGE = SS_CB * SS_PX * 4;
    constexpr int RR = R + (NORM2 ? 1 : 0);
    float* red = reinterpret_cast<float*>(ss_raw + (size_t)RS_NS * STAGE);            // [SS_W][RR][SS_PX]
    f32x2* Ms = reinterpret_cast<f32x2*>(red + (size_t)SS_W * RR * SS_PX);           // [C][R] {m, m}
    uint64_t* full = reinterpret_cast<uint64_t*>(Ms + (size_t)p.C * R);
    uint64_t* empty = full + RS_NS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nchunk = (p.HW + SS_PX - 1) / SS_PX, n_items = p.E * nchunk, nstage = p.C / SS_CB;
    const int i0 = (int)((long long)n_items * blockIdx.x / gridDim.x), i1 = (int)((long long)n_items * (blockIdx.x + 1) / gridDim.x);
    if (tid == 0) {
        for (int i = 0; i < RS_NS; ++i) { ls_mbar_init(&full[i], 1); ls_mbar_init(&empty[i], SS_W); }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    unsigned slot = 0, ph = 0;
    if (warp == 0) {
        for (int item = i0; item < i1; ++item) {
            const int e = item / nchunk, k = item - e * nchunk;
            for (int s = 0; s < nstage; ++s) {
                ls_wait(&empty[slot], ph ^ 1u);
                if (lane == 0) {
                    ls_expect_tx(&full[slot], STAGE);
                    ls_tma_2d(ss_raw + (size_t)slot * STAGE, &fmap, k * SS_PX, e * p.C + s * SS_CB, &full[slot]);
                }
                __syncwarp();
                if (++slot == RS_NS) { slot = 0; ph ^= 1u; }
            }
        }
        return;
    }
    const int cw = warp - 1, ctid = tid - 32;
    const uint32_t lane_base = ls_u32(ss_raw) + (uint32_t)cw * (SS_PX * 4) + (uint32_t)lane * 16u;
    int cur_e = -1;
    for (int item = i0; item < i1; ++item) {
        const int e = item / nchunk, k = item - e * nchunk;
        asm volatile("bar.sync 1, %0;" ::"n"(32 * SS_W) : "memory");          // everyone has left the previous item (red, Ms)
        if (e != cur_e) {
            const float* Mg = p.M + (size_t)e * p.R * p.C;
            for (int i = ctid; i < p.C * R; i += 32 * SS_W) {
                const int r = i / p.C, c = i - r * p.C;
                const float m = r < p.R ? Mg[i] : 0.f;
                Ms[c * R + r] = pk2(m, m);
            }
            asm volatile("bar.sync 1, %0;" ::"n"(32 * SS_W) : "memory");
            cur_e = e;
        }
        f32x2 acc[R][4], n2[4];
#pragma unroll
        for (int r = 0; r < R; ++r) { acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0ull; }
        n2[0] = n2[1] = n2[2] = n2[3] = 0ull;
        uint32_t maddr = ls_u32(Ms) + (uint32_t)cw * (R * 8u);
        for (int s = 0; s < nstage; ++s) {
            ls_wait(&full[slot], ph);
            const uint32_t st = lane_base + slot * STAGE;
#pragma unroll
            for (int kk = 0; kk < SS_CB / SS_W; ++kk) {
                const uint32_t row = st + (uint32_t)(kk * SS_W) * (SS_PX * 4);
                const ulonglong2 d0 = ls_lds128(row), d1 = ls_lds128(row + 512u);
#pragma unroll
                for (int r2 = 0; r2 < R / 2; ++r2) {
                    const ulonglong2 m = ls_lds128(maddr + (uint32_t)(kk * SS_W * R * 8 + r2 * 16));
                    acc[2 * r2][0] = fma2(m.x, d0.x, acc[2 * r2][0]); acc[2 * r2][1] = fma2(m.x, d0.y, acc[2 * r2][1]);
                    acc[2 * r2][2] = fma2(m.x, d1.x, acc[2 * r2][2]); acc[2 * r2][3] = fma2(m.x, d1.y, acc[2 * r2][3]);
                    acc[2 * r2 + 1][0] = fma2(m.y, d0.x, acc[2 * r2 + 1][0]); acc[2 * r2 + 1][1] = fma2(m.y, d0.y, acc[2 * r2 + 1][1]);
                    acc[2 * r2 + 1][2] = fma2(m.y, d1.x, acc[2 * r2 + 1][2]); acc[2 * r2 + 1][3] = fma2(m.y, d1.y, acc[2 * r2 + 1][3]);
                }
                if (NORM2) {
                    n2[0] = fma2(d0.x, d0.x, n2[0]); n2[1] = fma2(d0.y, d0.y, n2[1]);
                    n2[2] = fma2(d1.x, d1.x, n2[2]); n2[3] = fma2(d1.y, d1.y, n2[3]);
                }
            }
            maddr += (uint32_t)(SS_CB * R * 8);
            __syncwarp();
            if (lane == 0) ls_arrive(&empty[slot]);
            if (++slot == RS_NS) { slot = 0; ph ^= 1u; }
        }
        {
            ulonglong2* rw = reinterpret_cast<ulonglong2*>(red + (size_t)cw * RR * SS_PX);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                rw[r * (SS_PX / 4) + lane] = make_ulonglong2(acc[r][0], acc[r][1]);
                rw[r * (SS_PX / 4) + lane + 32] = make_ulonglong2(acc[r][2], acc[r][3]);
            }
            if (NORM2) {
                rw[R * (SS_PX / 4) + lane] = make_ulonglong2(n2[0], n2[1]);
                rw[R * (SS_PX / 4) + lane + 32] = make_ulonglong2(n2[2], n2[3]);
            }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * SS_W) : "memory");
        const int px = k * SS_PX + ctid;                       // 256 consumer threads <-> the 256 pixels of the chunk
        if (px < p.HW) {
#pragma unroll
            for (int r = 0; r < RR; ++r) {
                float s = 0.f;
#pragma unroll
                for (int j = 0; j < SS_W; ++j) s += red[((size_t)j * RR + r) * SS_PX + ctid];
                if (r < R) { if (r < p.R) p.out[((size_t)e * p.R + r) * p.HW + px] = s; }
                else p.n2[(size_t)e * p.HW + px] = s;
            }
        }
    }
}

// out[e][r][p] (r < r_actual <= 8), n2[e][p] (nullable); M [E][r_actual][C]. CWT_ERR_UNSUPPORTED (no error text) when the
// shape does not suit the tile copies.
static int launch_rtf_stream(const float* f, const float* M, float* out, float* n2, int E, int C, int HW, int r_actual,
                             cudaStream_t st) {
    if (r_actual < 1 || r_actual > 8 || C % SS_CB || HW % 4 || HW < SS_PX || (reinterpret_cast<uintptr_t>(f) & 15u))
        return CWT_ERR_UNSUPPORTED;
    const int R = r_actual <= 4 ? 4 : 8, RR = R + (n2 ? 1 : 0);
    const size_t sm = (size_t)RS_NS * SS_CB * SS_PX * 4 + (size_t)SS_W * RR * SS_PX * 4 + (size_t)C * R * 8 + 16 * RS_NS;
    int dev = 0, n_sm = 148, smem_cap = 232448;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&smem_cap, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (sm > (size_t)smem_cap) return CWT_ERR_UNSUPPORTED;
    CUtensorMap fmap;
    memset(&fmap, 0, sizeof(fmap));
    int rc = ls_make_map_f32(&fmap, f, (uint64_t)E * C, (uint64_t)HW, SS_CB, SS_PX, "rows_times_feat");
    if (rc != CWT_OK) return rc;
    RsParams p{M, out, n2, E, C, HW, r_actual};
    const int n_items = E * ((HW + SS_PX - 1) / SS_PX);
    const int grid = n_items < n_sm ? n_items : n_sm;
#define CWT_RTF_STREAM(RV, NV)                                                                                          \
    do {                                                                                                                \
        CWT_CUDA(cudaFuncSetAttribute(k_rtf_stream<RV, NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));     \
        k_rtf_stream<RV, NV><<<grid, SS_THREADS, sm, st>>>(fmap, p);                                                    \
    } while (0)
    if (R == 4) { if (n2) CWT_RTF_STREAM(4, true); else CWT_RTF_STREAM(4, false); }
    else        { if (n2) CWT_RTF_STREAM(8, true); else CWT_RTF_STREAM(8, false); }
#undef CWT_RTF_STREAM
    CWT_LAUNCHED("rows_times_feat_stream");
    return CWT_OK;
}

// out[img][r][p], M[img][r][c] (one weight set per image): streaming kernel when it applies, LDG kernel otherwise
static int launch_rows_times_feat_auto(const float* f, const float* M, float* out, float* n2, int n_img, int C, int HW,
                                       int imgs_per_M, int r_actual, cudaStream_t st) {
    if (imgs_per_M == 1 && skinny_stream_enabled()) {
        const int rc = launch_rtf_stream(f, M, out, n2, n_img, C, HW, r_actual, st);
        if (rc != CWT_ERR_UNSUPPORTED) return rc;
    }
    return launch_rows_times_feat(f, M, out, n2, n_img, C, HW, imgs_per_M, r_actual, st);
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
