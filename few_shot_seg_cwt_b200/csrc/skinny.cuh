// The two "skinny contraction" work-horses of the head. Both stream a feature map
// f[img][c][p] (NCHW, p = y*w+x contiguous) exactly once, fully coalesced, and stay on
// the CUDA cores: the other operand has only R <= 16 rows, so there is no tile for the
// tensor cores to reuse (BASELINE.json north_star (a)).
//
//   rows_times_feat (RTF):  out[img][r][p] = sum_c M[img/ipm][r][c] * f[img][c][p]       (+ |f[:,p]|^2)
//       1x1-conv logits of the fit (src/test.py:181), baseline / adapted query logits
//       (src/test.py:192,204) and the re-associated attention scores (Q_h A_h) X^T.
//   feat_times_cols (FTC):  out[e][r][c] = sum_{s,p} P[e][r][s*HW+p] * f[e*S+s][c][p]
//       conv weight gradient of the fit (G60 F^T, autograd of src/test.py:186) and a.X of
//       the attention.
//
// HBM bytes per call = the feature map once (+ R*HW*4 per image), which is what the
// roofline in DESIGN.md counts.
#pragma once
#include "common.cuh"

namespace cwt {

// ------------------------------------------------------------------------------------
// RTF: CTA = 4 warps, tile = 32*VEC pixels; warp w sums channels [w*C/4, (w+1)*C/4);
// partials are combined through shared memory (deterministic order).
// ------------------------------------------------------------------------------------
template <int R, int VEC, bool NORM2>
__global__ void __launch_bounds__(128)
k_rows_times_feat(const float* __restrict__ f, const float* __restrict__ M, float* __restrict__ out,
                  float* __restrict__ n2out, int C, int HW, int imgs_per_M, int r_actual) {
    extern __shared__ __align__(16) float smem[];
    float* Ms = smem;                         // [C][R] (transposed so one LDS.128 feeds 4 rows)
    const int img = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int p0 = blockIdx.x * (32 * VEC) + lane * VEC;
    const float* Mg = M + (size_t)(img / imgs_per_M) * r_actual * C;
    for (int i = tid; i < C * R; i += 128) {
        int c = i / R, r = i - c * R;
        Ms[i] = (r < r_actual) ? Mg[(size_t)r * C + c] : 0.f;
    }
    __syncthreads();

    float acc[R][VEC];
    float n2[VEC];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int k = 0; k < VEC; ++k) acc[r][k] = 0.f;
#pragma unroll
    for (int k = 0; k < VEC; ++k) n2[k] = 0.f;

    const int cchunk = (C + 3) / 4;
    const int c_begin = warp * cchunk;
    const int c_end = min(C, c_begin + cchunk);
    const bool valid = p0 < HW;               // HW % VEC == 0 is guaranteed by the dispatcher
    if (valid) {
        const float* fp = f + ((size_t)img * C) * HW + p0;
#pragma unroll 4
        for (int c = c_begin; c < c_end; ++c) {
            float v[VEC];
            if (VEC == 4) {
                float4 t = ldg_stream4(fp + (size_t)c * HW);
                v[0] = t.x; v[1 % VEC] = t.y; v[2 % VEC] = t.z; v[3 % VEC] = t.w;
            } else {
                v[0] = ldg_stream1(fp + (size_t)c * HW);
            }
            const float* mrow = Ms + c * R;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float m = mrow[r];
#pragma unroll
                for (int k = 0; k < VEC; ++k) acc[r][k] = fmaf(m, v[k], acc[r][k]);
            }
            if (NORM2) {
#pragma unroll
                for (int k = 0; k < VEC; ++k) n2[k] = fmaf(v[k], v[k], n2[k]);
            }
        }
    }
    __syncthreads();                          // everyone is done reading Ms: reuse it
    constexpr int RR = R + (NORM2 ? 1 : 0);
    float* red = smem;                        // [4][RR][32*VEC]
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int k = 0; k < VEC; ++k) red[(warp * RR + r) * (32 * VEC) + lane * VEC + k] = acc[r][k];
    if (NORM2) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) red[(warp * RR + R) * (32 * VEC) + lane * VEC + k] = n2[k];
    }
    __syncthreads();
    const int tile_p0 = blockIdx.x * (32 * VEC);
    for (int i = tid; i < RR * 32 * VEC; i += 128) {
        int r = i / (32 * VEC), pp = i - r * (32 * VEC);
        int p = tile_p0 + pp;
        if (p >= HW) continue;
        float s = (red[(0 * RR + r) * (32 * VEC) + pp] + red[(1 * RR + r) * (32 * VEC) + pp]) +
                  (red[(2 * RR + r) * (32 * VEC) + pp] + red[(3 * RR + r) * (32 * VEC) + pp]);
        if (r < R) {
            if (r < r_actual) out[((size_t)img * r_actual + r) * HW + p] = s;
        } else {
            n2out[(size_t)img * HW + p] = s;
        }
    }
}

static inline size_t rtf_smem_bytes(int R, int VEC, bool norm2, int C) {
    size_t a = (size_t)C * R * 4, b = (size_t)4 * (R + (norm2 ? 1 : 0)) * 32 * VEC * 4;
    return a > b ? a : b;
}

template <int R, int VEC, bool NORM2>
static int launch_rtf_t(const float* f, const float* M, float* out, float* n2, int n_img, int C, int HW,
                        int imgs_per_M, int r_actual, cudaStream_t st) {
    size_t sm = rtf_smem_bytes(R, VEC, NORM2, C);
    auto kern = k_rows_times_feat<R, VEC, NORM2>;
    if (sm > 48 * 1024) CWT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    dim3 grid((HW + 32 * VEC - 1) / (32 * VEC), n_img);
    kern<<<grid, 128, sm, st>>>(f, M, out, n2, C, HW, imgs_per_M, r_actual);
    CWT_LAUNCHED("rows_times_feat");
    return CWT_OK;
}

template <int R>
static int launch_rtf_r(const float* f, const float* M, float* out, float* n2, int n_img, int C, int HW,
                        int imgs_per_M, int r_actual, cudaStream_t st) {
    const bool vec4 = (HW % 4 == 0);
    if (n2) {
        return vec4 ? launch_rtf_t<R, 4, true>(f, M, out, n2, n_img, C, HW, imgs_per_M, r_actual, st)
                    : launch_rtf_t<R, 1, true>(f, M, out, n2, n_img, C, HW, imgs_per_M, r_actual, st);
    }
    return vec4 ? launch_rtf_t<R, 4, false>(f, M, out, n2, n_img, C, HW, imgs_per_M, r_actual, st)
                : launch_rtf_t<R, 1, false>(f, M, out, n2, n_img, C, HW, imgs_per_M, r_actual, st);
}

// out[img][r][p], r < r_actual <= 16 ; M[img/imgs_per_M][r][c] ; n2 (nullable) [img][p]
static int launch_rows_times_feat(const float* f, const float* M, float* out, float* n2, int n_img, int C,
                                  int HW, int imgs_per_M, int r_actual, cudaStream_t st) {
    CWT_REQUIRE(r_actual >= 1 && r_actual <= 16, CWT_ERR_UNSUPPORTED, "rows_times_feat: %d rows (max 16)", r_actual);
    CWT_REQUIRE((size_t)C * 16 * 4 <= 200 * 1024, CWT_ERR_UNSUPPORTED, "rows_times_feat: C=%d too large", C);
    if (r_actual <= 1) return launch_rtf_r<1>(f, M, out, n2, n_img, C, HW, imgs_per_M, r_actual, st);
    if (r_actual <= 2) return launch_rtf_r<2>(f, M, out, n2, n_img, C, HW, imgs_per_M, r_actual, st);
    if (r_actual <= 4) return launch_rtf_r<4>(f, M, out, n2, n_img, C, HW, imgs_per_M, r_actual, st);
    if (r_actual <= 8) return launch_rtf_r<8>(f, M, out, n2, n_img, C, HW, imgs_per_M, r_actual, st);
    if (r_actual <= 12) return launch_rtf_r<12>(f, M, out, n2, n_img, C, HW, imgs_per_M, r_actual, st);
    return launch_rtf_r<16>(f, M, out, n2, n_img, C, HW, imgs_per_M, r_actual, st);
}

// ------------------------------------------------------------------------------------
// FTC: CTA = 4 warps, warp = CPW channels, lanes stride the pixels (VEC each), R row
// accumulators per channel per lane, one shuffle tree at the end. The epilogue functor
// receives the complete sums (lane 0): deterministic, no atomics.
// ------------------------------------------------------------------------------------
struct StoreEpilogue {
    float* out; int C; int r_actual;
    __device__ __forceinline__ void operator()(int e, int r, int c, float v) const {
        if (r < r_actual) out[((size_t)e * r_actual + r) * C + c] = v;
    }
};

template <int R, int CPW, int VEC, class Epi>
__global__ void __launch_bounds__(128)
k_feat_times_cols(const float* __restrict__ f, const float* __restrict__ P, int C, int HW, int S,
                  int r_actual, Epi epi) {
    const int e = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c0 = (blockIdx.x * 4 + warp) * CPW;
    if (c0 >= C) return;
    float acc[CPW][R];
#pragma unroll
    for (int j = 0; j < CPW; ++j)
#pragma unroll
        for (int r = 0; r < R; ++r) acc[j][r] = 0.f;

    for (int s = 0; s < S; ++s) {
        const float* fb = f + ((size_t)(e * S + s) * C) * HW;
        const float* Pb = P + ((size_t)e * r_actual) * S * HW + (size_t)s * HW;
#pragma unroll 2
        for (int p = lane * VEC; p < HW; p += 32 * VEC) {
            float pv[R][VEC];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (r < r_actual) {
                    if (VEC == 4) {
                        float4 t = *reinterpret_cast<const float4*>(Pb + (size_t)r * S * HW + p);
                        pv[r][0] = t.x; pv[r][1 % VEC] = t.y; pv[r][2 % VEC] = t.z; pv[r][3 % VEC] = t.w;
                    } else {
                        pv[r][0] = Pb[(size_t)r * S * HW + p];
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < VEC; ++k) pv[r][k] = 0.f;
                }
            }
#pragma unroll
            for (int j = 0; j < CPW; ++j) {
                if (c0 + j < C) {
                    float v[VEC];
                    if (VEC == 4) {
                        float4 t = ldg_stream4(fb + (size_t)(c0 + j) * HW + p);
                        v[0] = t.x; v[1 % VEC] = t.y; v[2 % VEC] = t.z; v[3 % VEC] = t.w;
                    } else {
                        v[0] = ldg_stream1(fb + (size_t)(c0 + j) * HW + p);
                    }
#pragma unroll
                    for (int r = 0; r < R; ++r)
#pragma unroll
                        for (int k = 0; k < VEC; ++k) acc[j][r] = fmaf(pv[r][k], v[k], acc[j][r]);
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < CPW; ++j)
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float v = warp_sum(acc[j][r]);
            if (lane == 0 && c0 + j < C) epi(e, r, c0 + j, v);
        }
}

template <int R, int CPW, class Epi>
static int launch_ftc_t(const float* f, const float* P, int E, int S, int C, int HW, int r_actual, Epi epi,
                        cudaStream_t st) {
    dim3 grid((C + 4 * CPW - 1) / (4 * CPW), E);
    if (HW % 4 == 0) k_feat_times_cols<R, CPW, 4, Epi><<<grid, 128, 0, st>>>(f, P, C, HW, S, r_actual, epi);
    else             k_feat_times_cols<R, CPW, 1, Epi><<<grid, 128, 0, st>>>(f, P, C, HW, S, r_actual, epi);
    CWT_LAUNCHED("feat_times_cols");
    return CWT_OK;
}

// out[e][r][c] = sum_{s,p} P[e][r][s*HW+p] * f[e*S+s][c][p]   (r < r_actual <= 16)
static int launch_feat_times_cols(const float* f, const float* P, float* out, int E, int S, int C, int HW,
                                  int r_actual, cudaStream_t st) {
    CWT_REQUIRE(r_actual >= 1 && r_actual <= 16, CWT_ERR_UNSUPPORTED, "feat_times_cols: %d rows (max 16)", r_actual);
    StoreEpilogue epi{out, C, r_actual};
    if (r_actual <= 1) return launch_ftc_t<1, 4>(f, P, E, S, C, HW, r_actual, epi, st);
    if (r_actual <= 2) return launch_ftc_t<2, 4>(f, P, E, S, C, HW, r_actual, epi, st);
    if (r_actual <= 4) return launch_ftc_t<4, 4>(f, P, E, S, C, HW, r_actual, epi, st);
    if (r_actual <= 8) return launch_ftc_t<8, 2>(f, P, E, S, C, HW, r_actual, epi, st);     // (4 channels per warp: 168 registers, measured slower)
    return launch_ftc_t<16, 2>(f, P, E, S, C, HW, r_actual, epi, st);
}

}  // namespace cwt
