// (f-3) PSPNet.inner_loop on the reference's cosine classifier CosCls with ANY cls_type
// (src/model/pspnet.py:290-323; fitted by inner_loop, pspnet.py:189-205, SGD over classifier.parameters()):
//     x^    = F.normalize(x, p=2, dim=1, eps=1e-5)                      (the caller passes x^: cwt_normalize_features_f32)
//     'n'   : cls.weight.data = F.normalize(cls.weight.data, dim=1, eps=1e-5) at every forward (in place, outside autograd)
//     'r'   : WeightNorm.apply(cls, 'weight', dim=0): weight = weight_g * weight_v / ||weight_v||_row (parameters g, v;
//             the pre-forward hook recomputes weight, so 'n' has no effect when 'r' is set)
//     'b'   : cls has a bias                 't' : scale_factor is a parameter (initial value 2.0)
//     scores = scale_factor * (weight . x^ + bias)  ->  bilinear up  ->  CE (SegLoss 'wt_ce' / 'ce')
// Two classes, so the CE sees only the logit difference zd = s ((W1 - W0) . x^ + (b1 - b0)) and, with g60 = dL/dzd at 60x60
// (hires.cuh):   dW1 = s g60 . x^T = -dW0 ;  db1 = s sum(g60) = -db0 ;  ds = sum(g60 zd) / s ;
//                'r':  dg_r = dW_r . v^_r ,  dv_r = (g_r / ||v_r||) (dW_r - (dW_r . v^_r) v^_r)        (torch._weight_norm backward)
// Per SGD step: k_coscls_prepare (effective difference row), rows_times_feat<1>, k_fit_hires, feat_times_cols<1> (store),
// k_coscls_update (all parameters; deterministic block reductions). Streaming algorithm; not on the episodic hot path.
#include "common.cuh"
#include "skinny.cuh"
#include "hires.cuh"

namespace cwt {

int pack_label_cells(const void* labels, int label_kind, int n_img, int h, int w, int H, int W, int ignore_index,
                     uint4* cells, int32_t* counts, cudaStream_t st);

// deterministic sum over the 256 threads of a CTA (result in every thread)
__device__ __forceinline__ float block_sum_256(float v, float* red /* [8] */) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k];
    return t;
}

// start of a step, one CTA per episode: 'n' normalises the weight rows in place; wd_eff = s (Weff1 - Weff0), bd_eff = s (b1 - b0)
__global__ void __launch_bounds__(256)
k_coscls_prepare(float* __restrict__ weight, const float* __restrict__ weight_g, const float* __restrict__ bias,
                 const float* __restrict__ scale, int flags, int C, float* __restrict__ wd_eff, float* __restrict__ bd_eff) {
    __shared__ float red[8];
    const int e = blockIdx.x, tid = threadIdx.x;
    float* w0 = weight + (size_t)(e * 2) * C;
    float* w1 = w0 + C;
    float q0 = 0.f, q1 = 0.f;
    for (int c = tid; c < C; c += 256) { q0 = fmaf(w0[c], w0[c], q0); q1 = fmaf(w1[c], w1[c], q1); }
    const float n0 = sqrtf(block_sum_256(q0, red)), n1 = sqrtf(block_sum_256(q1, red));
    const float s = scale[e];
    float m0 = 1.f, m1 = 1.f;                                  // effective row = m_r * stored row
    if (flags & CWT_COSCLS_R) { m0 = weight_g[e * 2] / n0; m1 = weight_g[e * 2 + 1] / n1; }
    else if (flags & CWT_COSCLS_N) {
        const float i0 = 1.f / fmaxf(n0, 1e-5f), i1 = 1.f / fmaxf(n1, 1e-5f);
        for (int c = tid; c < C; c += 256) { w0[c] *= i0; w1[c] *= i1; }      // .data is overwritten (pspnet.py:304-305)
    }
    for (int c = tid; c < C; c += 256) wd_eff[(size_t)e * C + c] = s * (m1 * w1[c] - m0 * w0[c]);
    if (tid == 0) bd_eff[e] = bias ? s * (bias[e * 2 + 1] - bias[e * 2]) : 0.f;
}

// end of a step, one CTA per episode: plain SGD on every parameter from dwd = g60 . x^T (unscaled), g60 and zd_raw = wd_eff . x^
__global__ void __launch_bounds__(256)
k_coscls_update(const float* __restrict__ dwd, const float* __restrict__ g60, const float* __restrict__ zd_raw, int n,
                float* __restrict__ weight, float* __restrict__ weight_g, float* __restrict__ bias, float* __restrict__ scale,
                int flags, float lr, int C) {
    __shared__ float red[8];
    const int e = blockIdx.x, tid = threadIdx.x;
    float sg = 0.f, sgz = 0.f;
    for (int i = tid; i < n; i += 256) {
        const float g = g60[(size_t)e * n + i];
        sg += g;
        sgz = fmaf(g, zd_raw[(size_t)e * n + i], sgz);
    }
    sg = block_sum_256(sg, red);
    sgz = block_sum_256(sgz, red);
    const float s = scale[e];
    const float bdiff = bias ? bias[e * 2 + 1] - bias[e * 2] : 0.f;
    float* w0 = weight + (size_t)(e * 2) * C;
    float* w1 = w0 + C;
    const float* d = dwd + (size_t)e * C;
    if (flags & CWT_COSCLS_R) {
        float q0 = 0.f, q1 = 0.f, p0 = 0.f, p1 = 0.f;
        for (int c = tid; c < C; c += 256) {
            q0 = fmaf(w0[c], w0[c], q0); q1 = fmaf(w1[c], w1[c], q1);
            p0 = fmaf(d[c], w0[c], p0); p1 = fmaf(d[c], w1[c], p1);
        }
        const float n0 = sqrtf(block_sum_256(q0, red)), n1 = sqrtf(block_sum_256(q1, red));
        // dW1 = s dwd, dW0 = -s dwd ; dot_r = dW_r . v^_r
        const float dot0 = -s * block_sum_256(p0, red) / n0, dot1 = s * block_sum_256(p1, red) / n1;
        const float g0 = weight_g[e * 2], g1 = weight_g[e * 2 + 1];
        for (int c = tid; c < C; c += 256) {
            const float dv0 = (g0 / n0) * (-s * d[c] - dot0 * w0[c] / n0);
            const float dv1 = (g1 / n1) * (s * d[c] - dot1 * w1[c] / n1);
            w0[c] = fmaf(-lr, dv0, w0[c]);
            w1[c] = fmaf(-lr, dv1, w1[c]);
        }
        if (tid == 0) { weight_g[e * 2] = fmaf(-lr, dot0, g0); weight_g[e * 2 + 1] = fmaf(-lr, dot1, g1); }
    } else {
        for (int c = tid; c < C; c += 256) {
            const float dw1 = s * d[c];
            w0[c] = fmaf(lr, dw1, w0[c]);                    // dW0 = -dW1
            w1[c] = fmaf(-lr, dw1, w1[c]);
        }
    }
    if (tid == 0) {
        if (bias) {
            const float db1 = s * sg;
            bias[e * 2] = fmaf(lr, db1, bias[e * 2]);
            bias[e * 2 + 1] = fmaf(-lr, db1, bias[e * 2 + 1]);
        }
        if (flags & CWT_COSCLS_T) scale[e] = fmaf(-lr, sgz / s + bdiff * sg, s);     // ds = sum g60 (cos1 - cos0)
    }
}

struct CosWs {
    uint4* cells; int32_t* counts_img; float2* cw; float* wd; float* bd; float* zd; float* g60; float* dwd; float* loss_part; int nblk;
};

static size_t carve_cos(Carver& cv, CosWs& ws, int E, int S, int C, int h, int w) {
    ws.nblk = hires_bands(h);
    ws.cells = cv.take<uint4>((size_t)E * S * h * w);
    ws.counts_img = cv.take<int32_t>((size_t)E * S * 4);
    ws.cw = cv.take<float2>((size_t)E);
    ws.wd = cv.take<float>((size_t)E * C);
    ws.bd = cv.take<float>((size_t)E);
    ws.zd = cv.take<float>((size_t)E * S * h * w);
    ws.g60 = cv.take<float>((size_t)E * S * h * w);
    ws.dwd = cv.take<float>((size_t)E * C);
    ws.loss_part = cv.take<float>((size_t)E * S * ws.nblk);
    return align_up(cv.off);
}

}  // namespace cwt

using namespace cwt;

extern "C" size_t cwt_fit_coscls_workspace_bytes(int E, int S, int C, int h, int w, int H, int W) {
    (void)H; (void)W;
    Carver cv(nullptr, 0);
    CosWs ws;
    return carve_cos(cv, ws, E, S, C, h, w);
}

extern "C" int cwt_fit_coscls_f32(const float* x_norm, const void* s_label, int label_kind,
                                  float* weight, float* weight_g_or_null, float* bias_or_null, float* scale,
                                  const float* class_weight_or_null, float* loss_trace_or_null, int32_t* label_counts_or_null,
                                  int flags, int E, int S, int C, int h, int w, int H, int W,
                                  int n_iter, float lr, int ignore_index,
                                  void* workspace, size_t ws_bytes, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CWT_REQUIRE(E >= 0 && S >= 1 && C >= 1 && h >= 1 && w >= 1 && n_iter >= 0, CWT_ERR_INVALID_ARG,
                "fit_coscls: bad sizes E=%d S=%d C=%d h=%d w=%d n_iter=%d", E, S, C, h, w, n_iter);
    if (E == 0) return CWT_OK;
    CWT_REQUIRE(x_norm && s_label && weight && scale, CWT_ERR_INVALID_ARG, "fit_coscls: null pointer");
    CWT_REQUIRE((flags & ~7) == 0, CWT_ERR_INVALID_ARG, "fit_coscls: flags %d", flags);
    CWT_REQUIRE(!(flags & CWT_COSCLS_R) || weight_g_or_null, CWT_ERR_INVALID_ARG, "fit_coscls: flag 'r' needs weight_g");
    CWT_REQUIRE(H == 8 * (h - 1) + 1 && W == 8 * (w - 1) + 1, CWT_ERR_UNSUPPORTED,
                "fit_coscls: label size %dx%d is not 8*(%dx%d - 1) + 1 (align_corners scale must be 1/8)", H, W, h, w);
    CWT_REQUIRE(w <= HIRES_MAXW, CWT_ERR_UNSUPPORTED, "fit_coscls: w=%d exceeds the supported width (%d)", w, HIRES_MAXW);
    CWT_REQUIRE(label_kind == CWT_LABEL_U8 || label_kind == CWT_LABEL_I64, CWT_ERR_INVALID_ARG,
                "fit_coscls: label_kind %d", label_kind);
    Carver cv(workspace, ws_bytes);
    CosWs ws;
    const size_t need = carve_cos(cv, ws, E, S, C, h, w);
    CWT_REQUIRE(workspace && ws_bytes >= need, CWT_ERR_WORKSPACE, "fit_coscls: workspace %zu < %zu bytes", ws_bytes, need);

    const int HWl = h * w;
    int rc = pack_label_cells(s_label, label_kind, E * S, h, w, H, W, ignore_index, ws.cells, ws.counts_img, st);
    if (rc != CWT_OK) return rc;
    k_class_consts<<<(E + 127) / 128, 128, 0, st>>>(ws.counts_img, class_weight_or_null, 0.0, ws.cw,
                                                    label_counts_or_null, E, S);
    CWT_LAUNCHED("class_consts");
    StoreEpilogue epi{ws.dwd, C, 1};
    for (int it = 0; it < n_iter; ++it) {
        k_coscls_prepare<<<E, 256, 0, st>>>(weight, weight_g_or_null, bias_or_null, scale, flags, C, ws.wd, ws.bd);
        CWT_LAUNCHED("coscls_prepare");
        rc = launch_rows_times_feat(x_norm, ws.wd, ws.zd, nullptr, E * S, C, HWl, S, 1, st);
        if (rc != CWT_OK) return rc;
        if (loss_trace_or_null) {
            rc = launch_fit_hires<true>(ws.zd, ws.cells, ws.cw, ws.g60, ws.loss_part, E * S, h, w, S, st, ws.bd);
            if (rc != CWT_OK) return rc;
            k_reduce_loss<<<E, 32, 0, st>>>(ws.loss_part, loss_trace_or_null + (size_t)it * E, S * ws.nblk);
            CWT_LAUNCHED("reduce_loss");
        } else {
            rc = launch_fit_hires<false>(ws.zd, ws.cells, ws.cw, ws.g60, nullptr, E * S, h, w, S, st, ws.bd);
            if (rc != CWT_OK) return rc;
        }
        rc = launch_ftc_t<1, 4>(x_norm, ws.g60, E, S, C, HWl, 1, epi, st);
        if (rc != CWT_OK) return rc;
        k_coscls_update<<<E, 256, 0, st>>>(ws.dwd, ws.g60, ws.zd, S * HWl, weight, weight_g_or_null, bias_or_null, scale,
                                           flags, lr, C);
        CWT_LAUNCHED("coscls_update");
    }
    return CWT_OK;
}
