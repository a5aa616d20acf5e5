// (a-1..a-3) Support-classifier fit — replaces the 200-step autograd loop of the reference
// (src/test.py:164-187 = src/train.py:206-231 = PSPNet.inner_loop, src/model/pspnet.py:189-205).
//
// Math per SGD step, for every episode of the batch:
//     zd   = (W1 - W0) . F                       logit difference, [S, h*w]          (RTF<1>)
//     d    = up(zd)   bilinear, align_corners, scale exactly 1/8, at H x W
//     p    = sigmoid(d) = softmax(U)[1]          2-class softmax == sigmoid of the difference
//     g    = w[y] (p - y) / sum_i w[y_i]         (0 on ignored pixels)
//     g60  = up^T(g)                             adjoint as a deterministic gather   (k_fit_hires)
//     dW1  = g60 . F^T ;  dW0 = -dW1             (the two softmax gradients cancel)   (FTC<1>)
//     W1  -= lr dW1 ; W0 += lr dW1               plain SGD, no momentum / weight decay
//
// CWT_FIT_STREAM: three launches per step; the feature map is streamed twice per step
// from HBM/L2 (bytes_fit = (2T+1) S F per episode) — any S, any batch size.
#include "common.cuh"
#include "skinny.cuh"
#include "hires.cuh"
#include "../../include/cwt_b200_debug.h"

namespace cwt {

int pack_label_cells(const void* labels, int label_kind, int n_img, int h, int w, int H, int W, int ignore_index,
                     uint4* cells, int32_t* counts, cudaStream_t st);
size_t fit_resident_workspace_bytes(int E, int C, int h, int w);
int fit_resident(const float* f_s, const uint4* cells, const float2* cw, float* w_io, int E, int C, int h, int w,
                 int n_iter, float lr, void* ws, size_t ws_bytes, long long* prof_out, cudaStream_t st);
bool fit_l2_supported(int S, int C, int h, int w);
size_t fit_l2_workspace_bytes(int E, int S, int C, int h, int w);
int fit_l2(const float* f_s, const uint4* cells, const float2* cw, float* w_io, int E, int S, int C, int h, int w,
           int n_iter, float lr, void* ws, size_t ws_bytes, cudaStream_t st);

// one CTA per episode: W <- W0, Wd = W1 - W0
__global__ void __launch_bounds__(256)
k_fit_init_weights(const float* __restrict__ w0, float* __restrict__ w_out, float* __restrict__ wd, int C) {
    const int e = blockIdx.x;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float a = w0[(size_t)(e * 2) * C + c], b = w0[(size_t)(e * 2 + 1) * C + c];
        w_out[(size_t)(e * 2) * C + c] = a;
        w_out[(size_t)(e * 2 + 1) * C + c] = b;
        wd[(size_t)e * C + c] = b - a;
    }
}

// pass-2 epilogue: complete dW1[c] -> SGD update of both rows and the next step's difference
struct FitUpdateEpilogue {
    float* W; float* wd; float lr; int C;
    __device__ __forceinline__ void operator()(int e, int r, int c, float dw) const {
        if (r != 0) return;
        float* p0 = W + (size_t)(e * 2) * C + c;
        float* p1 = W + (size_t)(e * 2 + 1) * C + c;
        const float n0 = fmaf(lr, dw, *p0);       // dW0 = -dW1
        const float n1 = fmaf(-lr, dw, *p1);
        *p0 = n0; *p1 = n1;
        wd[(size_t)e * C + c] = n1 - n0;
    }
};

struct FitWs {
    uint4* cells; int32_t* counts_img; float2* cw; float* wd; float* zd; float* g60; float* loss_part;
    void* resident; size_t resident_bytes;
    int nblk;
};

static size_t carve_fit(Carver& cv, FitWs& ws, int E, int S, int C, int h, int w, int H, int W) {
    ws.nblk = hires_bands(h);
    ws.cells = cv.take<uint4>((size_t)E * S * h * w);
    ws.counts_img = cv.take<int32_t>((size_t)E * S * 4);
    ws.cw = cv.take<float2>((size_t)E);
    ws.wd = cv.take<float>((size_t)E * C);
    ws.zd = cv.take<float>((size_t)E * S * h * w);
    ws.g60 = cv.take<float>((size_t)E * S * h * w);
    ws.loss_part = cv.take<float>((size_t)E * S * ws.nblk);
    // the persistent kernels' exchange areas: resident (1 shot) or L2-streamed (several shots, head geometry)
    ws.resident_bytes = fit_l2_workspace_bytes(E, S, C, h, w);
    if (S == 1) { const size_t r = fit_resident_workspace_bytes(E, C, h, w); if (r > ws.resident_bytes) ws.resident_bytes = r; }
    ws.resident = cv.take<char>(ws.resident_bytes);
    return align_up(cv.off);
}

}  // namespace cwt

using namespace cwt;

extern "C" size_t cwt_fit_workspace_bytes(int E, int S, int C, int h, int w, int H, int W) {
    Carver cv(nullptr, 0);
    FitWs ws;
    return carve_fit(cv, ws, E, S, C, h, w, H, W);
}

static int fit_classifier_impl(const float* f_s, const void* s_label, int label_kind, const float* w0,
                               const float* class_weight_or_null, float* w_out,
                               float* loss_trace_or_null, int32_t* label_counts_or_null,
                               int E, int S, int C, int h, int w, int H, int W,
                               int n_iter, float lr, int ignore_index, int algo,
                               void* workspace, size_t ws_bytes, long long* prof_out, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CWT_REQUIRE(E >= 0 && S >= 1 && C >= 1 && h >= 1 && w >= 1 && n_iter >= 0, CWT_ERR_INVALID_ARG,
                "fit: bad sizes E=%d S=%d C=%d h=%d w=%d n_iter=%d", E, S, C, h, w, n_iter);
    if (E == 0) return CWT_OK;
    CWT_REQUIRE(f_s && s_label && w0 && w_out, CWT_ERR_INVALID_ARG, "fit: null pointer");
    CWT_REQUIRE(H == 8 * (h - 1) + 1 && W == 8 * (w - 1) + 1, CWT_ERR_UNSUPPORTED,
                "fit: label size %dx%d is not 8*(%dx%d - 1) + 1 (align_corners scale must be 1/8)", H, W, h, w);
    CWT_REQUIRE(w <= HIRES_MAXW, CWT_ERR_UNSUPPORTED, "fit: w=%d exceeds the supported width (%d)", w, HIRES_MAXW);
    CWT_REQUIRE(label_kind == CWT_LABEL_U8 || label_kind == CWT_LABEL_I64, CWT_ERR_INVALID_ARG,
                "fit: label_kind %d", label_kind);
    CWT_REQUIRE(algo == CWT_FIT_AUTO || algo == CWT_FIT_STREAM || algo == CWT_FIT_RESIDENT || algo == CWT_FIT_L2, CWT_ERR_INVALID_ARG,
                "fit: algo %d", algo);
    Carver cv(workspace, ws_bytes);
    FitWs ws;
    size_t need = carve_fit(cv, ws, E, S, C, h, w, H, W);
    CWT_REQUIRE(workspace && ws_bytes >= need, CWT_ERR_WORKSPACE, "fit: workspace %zu < %zu bytes", ws_bytes, need);

    const int HWl = h * w;
    int rc = pack_label_cells(s_label, label_kind, E * S, h, w, H, W, ignore_index, ws.cells, ws.counts_img, st);
    if (rc != CWT_OK) return rc;
    k_class_consts<<<(E + 127) / 128, 128, 0, st>>>(ws.counts_img, class_weight_or_null, 0.0, ws.cw,
                                                    label_counts_or_null, E, S);
    CWT_LAUNCHED("class_consts");
    k_fit_init_weights<<<E, 256, 0, st>>>(w0, w_out, ws.wd, C);
    CWT_LAUNCHED("fit_init_weights");

    // L2-streamed persistent kernel: several shots, head geometry, no loss trace (also on request for one shot)
    if ((algo == CWT_FIT_L2 || (algo == CWT_FIT_AUTO && S > 1)) && n_iter > 0) {
        if (!loss_trace_or_null && !prof_out && fit_l2_supported(S, C, h, w)) {
            rc = fit_l2(f_s, ws.cells, ws.cw, w_out, E, S, C, h, w, n_iter, lr, ws.resident, ws.resident_bytes, st);
            if (rc == CWT_OK) return CWT_OK;
            if (rc != CWT_ERR_UNSUPPORTED || algo == CWT_FIT_L2) return rc;
        } else if (algo == CWT_FIT_L2) {
            CWT_REQUIRE(false, CWT_ERR_UNSUPPORTED, "fit: CWT_FIT_L2 needs the head geometry (C=512, 60x60) and no loss trace (C=%d %dx%d)", C, h, w);
        }
    }
    // resident (shared-memory) algorithm: 1-shot, no loss trace, shape must fit on chip
    if (algo != CWT_FIT_STREAM && algo != CWT_FIT_L2 && n_iter > 0) {
        if (S == 1 && !loss_trace_or_null && HWl % 4 == 0) {
            rc = fit_resident(f_s, ws.cells, ws.cw, w_out, E, C, h, w, n_iter, lr, ws.resident, ws.resident_bytes,
                              prof_out, st);
            if (rc == CWT_OK) return CWT_OK;
            if (rc != CWT_ERR_UNSUPPORTED || algo == CWT_FIT_RESIDENT) return rc;
        } else if (algo == CWT_FIT_RESIDENT) {
            CWT_REQUIRE(false, CWT_ERR_UNSUPPORTED,
                        "fit: CWT_FIT_RESIDENT needs shot == 1, h*w %% 4 == 0 and no loss trace (S=%d, h*w=%d)", S, HWl);
        }
    }

    FitUpdateEpilogue epi{w_out, ws.wd, lr, C};
    for (int it = 0; it < n_iter; ++it) {
        rc = launch_rows_times_feat(f_s, ws.wd, ws.zd, nullptr, E * S, C, HWl, S, 1, st);
        if (rc != CWT_OK) return rc;
        if (loss_trace_or_null) {
            rc = launch_fit_hires<true>(ws.zd, ws.cells, ws.cw, ws.g60, ws.loss_part, E * S, h, w, S, st);
            if (rc != CWT_OK) return rc;
            k_reduce_loss<<<E, 32, 0, st>>>(ws.loss_part, loss_trace_or_null + (size_t)it * E, S * ws.nblk);
            CWT_LAUNCHED("reduce_loss");
        } else {
            rc = launch_fit_hires<false>(ws.zd, ws.cells, ws.cw, ws.g60, nullptr, E * S, h, w, S, st);
            if (rc != CWT_OK) return rc;
        }
        rc = launch_ftc_t<1, 4>(f_s, ws.g60, E, S, C, HWl, 1, epi, st);
        if (rc != CWT_OK) return rc;
    }
    return CWT_OK;
}

extern "C" int cwt_fit_classifier_f32(const float* f_s, const void* s_label, int label_kind, const float* w0,
                                      const float* class_weight_or_null, float* w_out,
                                      float* loss_trace_or_null, int32_t* label_counts_or_null,
                                      int E, int S, int C, int h, int w, int H, int W,
                                      int n_iter, float lr, int ignore_index, int algo,
                                      void* workspace, size_t ws_bytes, void* stream) {
    return fit_classifier_impl(f_s, s_label, label_kind, w0, class_weight_or_null, w_out, loss_trace_or_null,
                               label_counts_or_null, E, S, C, h, w, H, W, n_iter, lr, ignore_index, algo, workspace, ws_bytes,
                               nullptr, stream);
}

// developer entry point (include/cwt_b200_debug.h): the same fit on the instrumented resident kernel
extern "C" int cwt_debug_fit_classifier_prof_f32(const float* f_s, const void* s_label, int label_kind, const float* w0,
                                                 const float* class_weight_or_null, float* w_out,
                                                 int E, int C, int h, int w, int H, int W,
                                                 int n_iter, float lr, int ignore_index,
                                                 void* workspace, size_t ws_bytes, long long* prof_out, void* stream) {
    CWT_REQUIRE(prof_out, CWT_ERR_INVALID_ARG, "debug_fit_prof: null profile buffer");
    return fit_classifier_impl(f_s, s_label, label_kind, w0, class_weight_or_null, w_out, nullptr, nullptr,
                               E, 1, C, h, w, H, W, n_iter, lr, ignore_index, CWT_FIT_RESIDENT, workspace, ws_bytes,
                               prof_out, stream);
}

namespace cwt {
// one warp per episode: deferred error word of a fit (see cwt_fit_status in include/cwt_b200.h)
__global__ void __launch_bounds__(128)
k_fit_status(const int32_t* __restrict__ counts, const float* __restrict__ w, int has_cw, int32_t* __restrict__ status, int E, int C) {
    const int e = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (e >= E) return;
    bool bad = false;
    for (int i = lane; i < 2 * C; i += 32) bad |= (__float_as_uint(w[(size_t)e * 2 * C + i]) & 0x7f800000u) == 0x7f800000u;
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0)
        status[e] = (counts[e * 4 + 3] > 0 ? CWT_FIT_BAD_LABEL : 0) | ((!has_cw && counts[e * 4 + 1] == 0) ? CWT_FIT_NO_FG : 0) |
                    (bad ? CWT_FIT_NONFINITE : 0);
}
}  // namespace cwt

extern "C" int cwt_fit_status(const int32_t* label_counts, const float* w_fit, int has_class_weight, int32_t* status,
                              int E, int C, void* stream) {
    CWT_REQUIRE(E >= 0 && C >= 1, CWT_ERR_INVALID_ARG, "fit_status: bad sizes E=%d C=%d", E, C);
    if (E == 0) return CWT_OK;
    CWT_REQUIRE(label_counts && w_fit && status, CWT_ERR_INVALID_ARG, "fit_status: null pointer");
    k_fit_status<<<(E + 3) / 4, 128, 0, static_cast<cudaStream_t>(stream)>>>(label_counts, w_fit, has_class_weight, status, E, C);
    CWT_LAUNCHED("fit_status");
    return CWT_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// The same fit for a classifier WITH a bias (nn.Conv2d(C, 2, 1, bias=True); the reference builds one inside CosCls when
// cls_type[2] == 'b', src/model/pspnet.py:294,319): logits l_c = W_c . F + bias_scale * b_c, so the logit difference is
// zd + bias_scale * (b1 - b0) and  db1 = bias_scale * sum_p g60[p] = -db0  (plain SGD on both, like the weights).
// bias_scale = 1 for the dot classifier; CosCls multiplies conv(x_norm) INCLUDING its bias by scale_factor, and the
// caller folds scale_factor into the features, so it passes bias_scale = scale_factor. Streaming algorithm.
// ---------------------------------------------------------------------------------------------------------------
namespace cwt {
// one CTA per episode: b <- b0 (or, with g60, one SGD step on the bias from the step's logit gradient); bd = scale (b1 - b0)
__global__ void __launch_bounds__(256)
k_fit_bias_update(const float* __restrict__ g60, const float* __restrict__ b_init, float* __restrict__ b, float* __restrict__ bd,
                  float lr, float bias_scale, int n /* S*h*w */) {
    __shared__ float red[8];
    const int e = blockIdx.x, tid = threadIdx.x;
    float s = 0.f;
    if (g60) for (int i = tid; i < n; i += 256) s += g60[(size_t)e * n + i];
    s = warp_sum(s);
    if ((tid & 31) == 0) red[tid >> 5] = s;
    __syncthreads();
    if (tid == 0) {
        float tot = 0.f;
        for (int k = 0; k < 8; ++k) tot += red[k];                         // fixed order
        const float db1 = bias_scale * tot;
        const float o0 = b_init ? b_init[e * 2] : b[e * 2], o1 = b_init ? b_init[e * 2 + 1] : b[e * 2 + 1];
        const float n0 = fmaf(lr, db1, o0), n1 = fmaf(-lr, db1, o1);      // db0 = -db1
        b[e * 2] = n0; b[e * 2 + 1] = n1;
        bd[e] = bias_scale * (n1 - n0);
    }
}
}  // namespace cwt

extern "C" size_t cwt_fit_bias_workspace_bytes(int E, int S, int C, int h, int w, int H, int W) {
    Carver cv(nullptr, 0);
    FitWs ws;
    carve_fit(cv, ws, E, S, C, h, w, H, W);
    cv.take<float>((size_t)E);
    return align_up(cv.off);
}

extern "C" int cwt_fit_classifier_bias_f32(const float* f_s, const void* s_label, int label_kind, const float* w0,
                                           const float* b0, const float* class_weight_or_null, float* w_out, float* b_out,
                                           float* loss_trace_or_null, int32_t* label_counts_or_null,
                                           int E, int S, int C, int h, int w, int H, int W,
                                           int n_iter, float lr, float bias_scale, int ignore_index,
                                           void* workspace, size_t ws_bytes, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CWT_REQUIRE(E >= 0 && S >= 1 && C >= 1 && h >= 1 && w >= 1 && n_iter >= 0, CWT_ERR_INVALID_ARG,
                "fit_bias: bad sizes E=%d S=%d C=%d h=%d w=%d n_iter=%d", E, S, C, h, w, n_iter);
    if (E == 0) return CWT_OK;
    CWT_REQUIRE(f_s && s_label && w0 && b0 && w_out && b_out, CWT_ERR_INVALID_ARG, "fit_bias: null pointer");
    CWT_REQUIRE(H == 8 * (h - 1) + 1 && W == 8 * (w - 1) + 1, CWT_ERR_UNSUPPORTED,
                "fit_bias: label size %dx%d is not 8*(%dx%d - 1) + 1 (align_corners scale must be 1/8)", H, W, h, w);
    CWT_REQUIRE(w <= HIRES_MAXW, CWT_ERR_UNSUPPORTED, "fit_bias: w=%d exceeds the supported width (%d)", w, HIRES_MAXW);
    CWT_REQUIRE(label_kind == CWT_LABEL_U8 || label_kind == CWT_LABEL_I64, CWT_ERR_INVALID_ARG,
                "fit_bias: label_kind %d", label_kind);
    Carver cv(workspace, ws_bytes);
    FitWs ws;
    carve_fit(cv, ws, E, S, C, h, w, H, W);
    float* bd = cv.take<float>((size_t)E);
    const size_t need = align_up(cv.off);
    CWT_REQUIRE(workspace && ws_bytes >= need, CWT_ERR_WORKSPACE, "fit_bias: workspace %zu < %zu bytes", ws_bytes, need);

    const int HWl = h * w;
    int rc = pack_label_cells(s_label, label_kind, E * S, h, w, H, W, ignore_index, ws.cells, ws.counts_img, st);
    if (rc != CWT_OK) return rc;
    k_class_consts<<<(E + 127) / 128, 128, 0, st>>>(ws.counts_img, class_weight_or_null, 0.0, ws.cw,
                                                    label_counts_or_null, E, S);
    CWT_LAUNCHED("class_consts");
    k_fit_init_weights<<<E, 256, 0, st>>>(w0, w_out, ws.wd, C);
    CWT_LAUNCHED("fit_init_weights");
    k_fit_bias_update<<<E, 256, 0, st>>>(nullptr, b0, b_out, bd, 0.f, bias_scale, 0);
    CWT_LAUNCHED("fit_bias_init");

    FitUpdateEpilogue epi{w_out, ws.wd, lr, C};
    for (int it = 0; it < n_iter; ++it) {
        rc = launch_rows_times_feat(f_s, ws.wd, ws.zd, nullptr, E * S, C, HWl, S, 1, st);
        if (rc != CWT_OK) return rc;
        if (loss_trace_or_null) {
            rc = launch_fit_hires<true>(ws.zd, ws.cells, ws.cw, ws.g60, ws.loss_part, E * S, h, w, S, st, bd);
            if (rc != CWT_OK) return rc;
            k_reduce_loss<<<E, 32, 0, st>>>(ws.loss_part, loss_trace_or_null + (size_t)it * E, S * ws.nblk);
            CWT_LAUNCHED("reduce_loss");
        } else {
            rc = launch_fit_hires<false>(ws.zd, ws.cells, ws.cw, ws.g60, nullptr, E * S, h, w, S, st, bd);
            if (rc != CWT_OK) return rc;
        }
        rc = launch_ftc_t<1, 4>(f_s, ws.g60, E, S, C, HWl, 1, epi, st);
        if (rc != CWT_OK) return rc;
        k_fit_bias_update<<<E, 256, 0, st>>>(ws.g60, nullptr, b_out, bd, lr, bias_scale, S * HWl);
        CWT_LAUNCHED("fit_bias_update");
    }
    return CWT_OK;
}
