// (f-3) PSPNet.increment_inner_loop with MORE THAN TWO classes — the multi-way incremental setting of the fork
// (src/model/pspnet.py:207-221, called from src/train_cca.py:146,322 with num_classes_tr (+1) = 16 / 17 on PASCAL,
// 61 / 62 on COCO): continue fitting a K-class bias-free 1x1 classifier with plain SGD on
// Adapt_SegLoss -> weighted_adpt_ce_loss (src/model/model_util.py:76-98) = CrossEntropyLoss(weight[K], ignore 255) of the
// logits up-sampled to H x W (bilinear, align_corners, scale exactly 1/8).
//
// One classifier per call (the reference fits one episode at a time). Per SGD step:
//     L60[s][k]  = W_k . F_s                         rows_times_feat, <= 16 rows per launch
//     G[s][k]    = w[y] (softmax_k(up(L60)) - 1[k == y]) / sum_i w[y_i]   at H x W     (k_mc_softmax_grad)
//     g60[k][s]  = up^T(G[s][k])                     deterministic gather over the <= 15 x 15 window  (k_mc_adjoint)
//     W_k       -= lr * sum_s g60[k][s] . F_s^T      feat_times_cols + SGD epilogue, <= 16 rows per launch
// Straightforward streaming kernels (the hi-res gradient IS materialised here, K * H * W floats): this variant is not on
// the episodic hot path, it completes the inner-loop family for the fork's scripts.
#include "common.cuh"
#include "skinny.cuh"

namespace cwt {

template <bool I64>
__device__ __forceinline__ long long mc_label(const void* lab, size_t i) {
    return I64 ? reinterpret_cast<const long long*>(lab)[i] : (long long)reinterpret_cast<const uint8_t*>(lab)[i];
}

// one thread per hi-res pixel of one image: K up-sampled logits (recomputed per pass), softmax, weighted CE gradient
template <bool I64>
__global__ void __launch_bounds__(256)
k_mc_softmax_grad(const float* __restrict__ l60 /* [S][K][h*w] */, const void* __restrict__ labels /* [S][H][W] */,
                  const float* __restrict__ cls_w /* [K] */, const float* __restrict__ inv_sumw /* [1] */,
                  float* __restrict__ G /* [S][K][H*W] */, int K, int h, int w, int H, int W, int ignore_index) {
    const int s = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= H * W) return;
    const int Y = i / W, X = i - Y * W;
    const int y0 = Y >> 3, x0 = X >> 3, y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
    const float ly = (float)(Y & 7) * 0.125f, lx = (float)(X & 7) * 0.125f;
    const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
    const int i00 = y0 * w + x0, i01 = y0 * w + x1, i10 = y1 * w + x0, i11 = y1 * w + x1;
    const float* base = l60 + (size_t)s * K * h * w;
    float* g = G + (size_t)s * K * H * W + i;
    const long long y = mc_label<I64>(labels, (size_t)s * H * W + i);
    if (y == (long long)ignore_index || y < 0 || y >= K) {            // ignored (or invalid: the caller checks) -> no gradient
        for (int k = 0; k < K; ++k) g[(size_t)k * H * W] = 0.f;
        return;
    }
    auto up = [&](int k) {
        const float* l = base + (size_t)k * h * w;
        return w00 * l[i00] + w01 * l[i01] + w10 * l[i10] + w11 * l[i11];
    };
    float m = -INFINITY;
    for (int k = 0; k < K; ++k) m = fmaxf(m, up(k));
    float sum = 0.f;
    for (int k = 0; k < K; ++k) sum += __expf(up(k) - m);
    const float coef = cls_w[(int)y] * inv_sumw[0], inv = 1.f / sum;
    for (int k = 0; k < K; ++k) {
        const float p = __expf(up(k) - m) * inv;
        g[(size_t)k * H * W] = coef * (p - (k == (int)y ? 1.f : 0.f));
    }
}

// one thread per (class, low-res pixel) of one image: adjoint bilinear map as a gather
__global__ void __launch_bounds__(256)
k_mc_adjoint(const float* __restrict__ G /* [S][K][H*W] */, float* __restrict__ g60 /* [K][S][h*w] */,
             int K, int S, int h, int w, int H, int W) {
    const int s = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K * h * w) return;
    const int k = i / (h * w), q = i - k * h * w, a = q / w, b = q - a * w;
    const float* g = G + ((size_t)s * K + k) * H * W;
    const int Y0 = max(8 * a - 7, 0), Y1 = min(8 * a + 7, H - 1), X0 = max(8 * b - 7, 0), X1 = min(8 * b + 7, W - 1);
    float acc = 0.f;
    for (int Y = Y0; Y <= Y1; ++Y) {
        const float wy = 1.f - fabsf((float)(Y - 8 * a)) * 0.125f;
        float row = 0.f;
        for (int X = X0; X <= X1; ++X) row = fmaf(1.f - fabsf((float)(X - 8 * b)) * 0.125f, g[(size_t)Y * W + X], row);
        acc = fmaf(wy, row, acc);
    }
    g60[((size_t)k * S + s) * h * w + q] = acc;
}

// FTC epilogue: rows r0 .. r0 + r_actual of the classifier take one SGD step
struct McUpdateEpilogue {
    float* W; float lr; int C; int r_actual;
    __device__ __forceinline__ void operator()(int, int r, int c, float dw) const {
        if (r >= r_actual) return;
        float* p = W + (size_t)r * C + c;
        *p = fmaf(-lr, dw, *p);
    }
};

}  // namespace cwt

using namespace cwt;

extern "C" size_t cwt_fit_multiclass_workspace_bytes(int K, int S, int C, int h, int w, int H, int W) {
    (void)C;
    return align_up((size_t)S * K * h * w * 4) * 2 + align_up((size_t)S * K * H * W * 4);
}

extern "C" int cwt_fit_multiclass_f32(const float* f_s, const void* s_label, int label_kind, float* weight,
                                      const float* class_weight, const float* inv_sum_weight,
                                      int K, int S, int C, int h, int w, int H, int W,
                                      int n_iter, float lr, int ignore_index,
                                      void* workspace, size_t ws_bytes, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CWT_REQUIRE(K >= 2 && S >= 1 && C >= 1 && h >= 1 && w >= 1 && n_iter >= 0, CWT_ERR_INVALID_ARG,
                "fit_multiclass: bad sizes K=%d S=%d C=%d h=%d w=%d n_iter=%d", K, S, C, h, w, n_iter);
    CWT_REQUIRE(f_s && s_label && weight && class_weight && inv_sum_weight, CWT_ERR_INVALID_ARG, "fit_multiclass: null pointer");
    CWT_REQUIRE(H == 8 * (h - 1) + 1 && W == 8 * (w - 1) + 1, CWT_ERR_UNSUPPORTED,
                "fit_multiclass: label size %dx%d is not 8*(%dx%d - 1) + 1 (align_corners scale must be 1/8)", H, W, h, w);
    CWT_REQUIRE(label_kind == CWT_LABEL_U8 || label_kind == CWT_LABEL_I64, CWT_ERR_INVALID_ARG,
                "fit_multiclass: label_kind %d", label_kind);
    const size_t need = cwt_fit_multiclass_workspace_bytes(K, S, C, h, w, H, W);
    CWT_REQUIRE(workspace && ws_bytes >= need, CWT_ERR_WORKSPACE, "fit_multiclass: workspace %zu < %zu bytes", ws_bytes, need);
    Carver cv(workspace, ws_bytes);
    const int HWl = h * w;
    float* l60 = cv.take<float>((size_t)S * K * HWl);
    float* g60 = cv.take<float>((size_t)S * K * HWl);
    float* G = cv.take<float>((size_t)S * K * H * W);

    for (int it = 0; it < n_iter; ++it) {
        for (int s = 0; s < S; ++s)
            for (int r0 = 0; r0 < K; r0 += 16) {
                const int nr = K - r0 < 16 ? K - r0 : 16;
                int rc = launch_rows_times_feat(f_s + (size_t)s * C * HWl, weight + (size_t)r0 * C, l60 + ((size_t)s * K + r0) * HWl,
                                                nullptr, 1, C, HWl, 1, nr, st);
                if (rc != CWT_OK) return rc;
            }
        const dim3 gp((H * W + 255) / 256, S);
        if (label_kind == CWT_LABEL_I64)
            k_mc_softmax_grad<true><<<gp, 256, 0, st>>>(l60, s_label, class_weight, inv_sum_weight, G, K, h, w, H, W, ignore_index);
        else
            k_mc_softmax_grad<false><<<gp, 256, 0, st>>>(l60, s_label, class_weight, inv_sum_weight, G, K, h, w, H, W, ignore_index);
        CWT_LAUNCHED("mc_softmax_grad");
        k_mc_adjoint<<<dim3((K * HWl + 255) / 256, S), 256, 0, st>>>(G, g60, K, S, h, w, H, W);
        CWT_LAUNCHED("mc_adjoint");
        for (int r0 = 0; r0 < K; r0 += 16) {
            const int nr = K - r0 < 16 ? K - r0 : 16;
            McUpdateEpilogue epi{weight + (size_t)r0 * C, lr, C, nr};
            int rc = launch_ftc_t<16, 2>(f_s, g60 + (size_t)r0 * S * HWl, 1, S, C, HWl, nr, epi, st);
            if (rc != CWT_OK) return rc;
        }
    }
    return CWT_OK;
}
