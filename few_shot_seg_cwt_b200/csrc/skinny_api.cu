// C-ABI entry points of the two skinny contractions, used by the host layer for the training
// step (a-13): logits = W' X^T (src/train.py:259-261) and its adjoint dW' = dlogits X.
#include "common.cuh"
#include "skinny.cuh"
#include "skinny_stream.cuh"

namespace cwt {

// inv_n[e][p] = 1 / max(|f[e,:,p]|_2, 1e-12)  (F.normalize(dim=1), src/test.py:194)
__global__ void __launch_bounds__(256)
k_inv_norm_from_n2(const float* __restrict__ n2, float* __restrict__ inv_n, size_t total) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) inv_n[i] = 1.f / fmaxf(sqrtf(n2[i]), 1e-12f);
}
// x[e][r][p] *= inv_n[e][p]
__global__ void __launch_bounds__(256)
k_scale_rows(float* __restrict__ x, const float* __restrict__ inv_n, int R, int HW, size_t total) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int p = (int)(i % HW);
    const size_t e = i / ((size_t)R * HW);
    x[i] *= inv_n[e * HW + p];
}

// out[c][p] = scale * f[c][p] / max(|f[:,p]|_2, eps): the feature side of the cosine classifier
// (CosCls.forward, src/model/pspnet.py:302-310: F.normalize(x, p=2, dim=1, eps=1e-5), scores = 2.0 * conv(x_norm)).
// CTA = (tile of 128 or 32 pixels, image); 4 warps split the channels; the second sweep re-reads the tile from L2.
template <int VEC>
__global__ void __launch_bounds__(128)
k_normalize_features(const float* __restrict__ f, float* __restrict__ out, int C, int HW, float eps, float scale) {
    __shared__ float part[4][32 * VEC];
    const int img = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int p0 = blockIdx.x * (32 * VEC) + lane * VEC;
    const bool valid = p0 < HW;                               // HW % VEC == 0
    const int cchunk = (C + 3) / 4, c_begin = warp * cchunk, c_end = min(C, c_begin + cchunk);
    const float* fp = f + (size_t)img * C * HW + p0;
    float* op = out + (size_t)img * C * HW + p0;
    float n2[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) n2[k] = 0.f;
    if (valid) {
#pragma unroll 4
        for (int c = c_begin; c < c_end; ++c) {
            if (VEC == 4) {
                const float4 t = *reinterpret_cast<const float4*>(fp + (size_t)c * HW);
                n2[0] = fmaf(t.x, t.x, n2[0]); n2[1 % VEC] = fmaf(t.y, t.y, n2[1 % VEC]);
                n2[2 % VEC] = fmaf(t.z, t.z, n2[2 % VEC]); n2[3 % VEC] = fmaf(t.w, t.w, n2[3 % VEC]);
            } else {
                const float t = fp[(size_t)c * HW];
                n2[0] = fmaf(t, t, n2[0]);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) part[warp][lane * VEC + k] = n2[k];
    __syncthreads();
    float den[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
        const int i = lane * VEC + k;
        den[k] = fmaxf(sqrtf((part[0][i] + part[1][i]) + (part[2][i] + part[3][i])), eps);
    }
    if (!valid) return;
#pragma unroll 4
    for (int c = c_begin; c < c_end; ++c) {
        if (VEC == 4) {
            float4 t = *reinterpret_cast<const float4*>(fp + (size_t)c * HW);
            t.x = scale * (t.x / den[0]); t.y = scale * (t.y / den[1 % VEC]);
            t.z = scale * (t.z / den[2 % VEC]); t.w = scale * (t.w / den[3 % VEC]);
            *reinterpret_cast<float4*>(op + (size_t)c * HW) = t;
        } else {
            op[(size_t)c * HW] = scale * (fp[(size_t)c * HW] / den[0]);
        }
    }
}

}  // namespace cwt
using namespace cwt;

extern "C" int cwt_normalize_features_f32(const float* f, float* out, int n_img, int C, int HW, float eps, float scale,
                                          void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n_img == 0) return CWT_OK;
    CWT_REQUIRE(f && out && n_img > 0 && C >= 1 && HW >= 1 && eps > 0.f, CWT_ERR_INVALID_ARG, "normalize_features: bad argument");
    if (HW % 4 == 0 && (reinterpret_cast<uintptr_t>(f) & 15u) == 0 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0)
        k_normalize_features<4><<<dim3((HW + 127) / 128, n_img), 128, 0, st>>>(f, out, C, HW, eps, scale);
    else
        k_normalize_features<1><<<dim3((HW + 31) / 32, n_img), 128, 0, st>>>(f, out, C, HW, eps, scale);
    CWT_LAUNCHED("normalize_features");
    return CWT_OK;
}

extern "C" size_t cwt_skinny_workspace_bytes(int E, int R, int C, int HW) {
    return align_up((size_t)E * HW * 4) * 3 + align_up((size_t)E * R * HW * 4) + align_up((size_t)E * C * 4);
}

// out[e][r][p] = sum_c M[e][r][c] * fn[e][c][p],  fn = f or F.normalize(f, dim=1)
extern "C" int cwt_rows_times_feat(const float* M, const float* f, int normalize, float* out,
                                   int E, int R, int C, int HW, void* workspace, size_t ws_bytes, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (E == 0) return CWT_OK;
    CWT_REQUIRE(M && f && out && E > 0 && R >= 1 && C >= 1 && HW >= 1, CWT_ERR_INVALID_ARG, "rows_times_feat: bad argument");
    Carver cv(workspace, ws_bytes);
    float* n2 = cv.take<float>((size_t)E * HW);
    float* inv_n = cv.take<float>((size_t)E * HW);
    CWT_REQUIRE(!normalize || (workspace && cv.ok()), CWT_ERR_WORKSPACE, "rows_times_feat: workspace too small");
    int rc = launch_rows_times_feat_auto(f, M, out, normalize ? n2 : nullptr, E, C, HW, 1, R, st);
    if (rc) return rc;
    if (normalize) {
        size_t t1 = (size_t)E * HW, t2 = (size_t)E * R * HW;
        k_inv_norm_from_n2<<<(unsigned)((t1 + 255) / 256), 256, 0, st>>>(n2, inv_n, t1);
        CWT_LAUNCHED("inv_norm");
        k_scale_rows<<<(unsigned)((t2 + 255) / 256), 256, 0, st>>>(out, inv_n, R, HW, t2);
        CWT_LAUNCHED("scale_rows");
    }
    return CWT_OK;
}

// out[e][r][c] = sum_p P[e][r][p] * fn[e][c][p]
extern "C" int cwt_feat_times_rows(const float* P, const float* f, int normalize, float* out,
                                   int E, int R, int C, int HW, void* workspace, size_t ws_bytes, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (E == 0) return CWT_OK;
    CWT_REQUIRE(P && f && out && E > 0 && R >= 1 && C >= 1 && HW >= 1, CWT_ERR_INVALID_ARG, "feat_times_rows: bad argument");
    if (!normalize) return launch_feat_times_cols_auto(f, P, out, E, 1, C, HW, R, st);
    Carver cv(workspace, ws_bytes);
    float* n2 = cv.take<float>((size_t)E * HW);
    float* inv_n = cv.take<float>((size_t)E * HW);
    float* dummy = cv.take<float>((size_t)E * HW);
    float* Pm = cv.take<float>((size_t)E * R * HW);
    float* ones = cv.take<float>((size_t)E * C);
    CWT_REQUIRE(workspace && cv.ok(), CWT_ERR_WORKSPACE, "feat_times_rows: workspace too small");
    CWT_CUDA(cudaMemsetAsync(ones, 0, sizeof(float) * (size_t)E * C, st));
    int rc = launch_rows_times_feat(f, ones, dummy, n2, E, C, HW, 1, 1, st);     // one pass for |f|^2
    if (rc) return rc;
    size_t t1 = (size_t)E * HW, t2 = (size_t)E * R * HW;
    k_inv_norm_from_n2<<<(unsigned)((t1 + 255) / 256), 256, 0, st>>>(n2, inv_n, t1);
    CWT_LAUNCHED("inv_norm");
    CWT_CUDA(cudaMemcpyAsync(Pm, P, sizeof(float) * t2, cudaMemcpyDeviceToDevice, st));
    k_scale_rows<<<(unsigned)((t2 + 255) / 256), 256, 0, st>>>(Pm, inv_n, R, HW, t2);
    CWT_LAUNCHED("scale_rows");
    return launch_feat_times_cols_auto(f, Pm, out, E, 1, C, HW, R, st);
}
