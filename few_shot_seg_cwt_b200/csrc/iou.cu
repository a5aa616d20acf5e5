// (a-4, a-7..a-12) query logits -> bilinear up to H x W -> argmax -> intersection / union (+ CE),
// and (a-13) the weighted training loss on the query with its gradient w.r.t. the 60x60 logits.
// Replaces src/test.py:192,200-204,214-223, src/util.py:237-308 and src/train.py:259-264.
//
// Nothing is ever written at H x W: the up-sampled logits live in registers, counts are
// reduced with warp votes -> integer atomics in shared memory -> one global atomic per
// counter per CTA. U is counted directly as |{pred==c or tgt==c} and valid| which equals the
// reference's area_output + area_target - area_intersection (src/util.py:306).
#include "common.cuh"
#include "skinny.cuh"
#include "hires.cuh"

namespace cwt {

int pack_label_cells(const void* labels, int label_kind, int n_img, int h, int w, int H, int W, int ignore_index,
                     uint4* cells, int32_t* counts, cudaStream_t st);

constexpr int IOU_THREADS = 256;
constexpr int IOU_MAXW = 128;

// grid (h, n_maps): CTA a0 owns hi-res rows [8 a0, 8 a0 + 8) of one 2-channel logit map
template <bool I64>
__global__ void __launch_bounds__(IOU_THREADS)
k_upsample_argmax_iou(const float* __restrict__ l60, const void* __restrict__ lab, int maps_per_label,
                      unsigned long long* __restrict__ counts, double* __restrict__ ce,
                      int h, int w, int H, int W, int ignore_index) {
    __shared__ float ls[2][2][IOU_MAXW];
    __shared__ int scnt[6];
    __shared__ float sloss[IOU_THREADS / 32];
    const int n = blockIdx.y, a0 = blockIdx.x, a1 = min(a0 + 1, h - 1);
    const int tid = threadIdx.x;
    const float* lm = l60 + (size_t)n * 2 * h * w;
    for (int i = tid; i < 4 * w; i += IOU_THREADS) {
        int ch = i / (2 * w), rem = i - ch * 2 * w, rr = rem / w, b = rem - rr * w;
        ls[ch][rr][b] = lm[(size_t)ch * h * w + (rr ? a1 : a0) * w + b];
    }
    if (tid < 6) scnt[tid] = 0;
    __syncthreads();
    const size_t lbase = (size_t)(n / maps_per_label) * H * W;
    const int rows = min(8, H - 8 * a0);
    int cI0 = 0, cI1 = 0, cU0 = 0, cU1 = 0, cT0 = 0, cT1 = 0, nvalid = 0;
    float loss = 0.f;
    for (int idx = tid; idx < rows * W; idx += IOU_THREADS) {
        const int r = idx / W, X = idx - r * W, Y = 8 * a0 + r;
        const float h1 = r * 0.125f, h0 = 1.f - h1;
        const int b0 = X >> 3, b1 = min(b0 + 1, w - 1);
        const float w1 = (X & 7) * 0.125f, w0 = 1.f - w1;
        const float u0 = bilerp8(ls[0][0][b0], ls[0][0][b1], ls[0][1][b0], ls[0][1][b1], w0, w1, h0, h1);
        const float u1 = bilerp8(ls[1][0][b0], ls[1][0][b1], ls[1][1][b0], ls[1][1][b1], w0, w1, h0, h1);
        const int pred = (u1 > u0) ? 1 : 0;                       // torch.argmax: first index wins ties
        const int code = load_label_code<I64>(lab, lbase + (size_t)Y * W + X, ignore_index);
        const bool valid = code < 2;
        cI0 += (valid && code == 0 && pred == 0);
        cI1 += (valid && code == 1 && pred == 1);
        cU0 += (valid && (code == 0 || pred == 0));
        cU1 += (valid && (code == 1 || pred == 1));
        cT0 += (code == 0);
        cT1 += (code == 1);
        if (valid) {
            nvalid += 1;
            const float t = (code == 1) ? (u0 - u1) : (u1 - u0);   // -log softmax(u)[y] = softplus(t)
            loss += fmaxf(t, 0.f) + log1pf(expf(-fabsf(t)));
        }
    }
    int v[6] = {cI0, cU0, cT0, cI1, cU1, cT1};                    // layout [class][I,U,T]
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        int s = __reduce_add_sync(0xffffffffu, v[k]);
        if ((tid & 31) == 0 && s) atomicAdd(&scnt[k], s);
    }
    loss = warp_sum(loss);
    nvalid = __reduce_add_sync(0xffffffffu, nvalid);
    if ((tid & 31) == 0) sloss[tid >> 5] = loss;
    __shared__ int svalid;
    if (tid == 0) svalid = 0;
    __syncthreads();
    if ((tid & 31) == 0 && nvalid) atomicAdd(&svalid, nvalid);
    __syncthreads();
    if (tid < 6 && scnt[tid]) atomicAdd(&counts[(size_t)n * 6 + tid], (unsigned long long)scnt[tid]);
    if (tid == 0 && ce) {
        double s = 0.0;
        for (int k = 0; k < IOU_THREADS / 32; ++k) s += (double)sloss[k];
        atomicAdd(&ce[(size_t)n * 2], s);
        atomicAdd(&ce[(size_t)n * 2 + 1], (double)svalid);
    }
}

static int launch_upsample_iou(const float* l60, const void* lab, int label_kind, int maps_per_label,
                               long long* counts, double* ce, int n, int h, int w, int H, int W,
                               int ignore_index, cudaStream_t st) {
    CWT_REQUIRE(H == 8 * (h - 1) + 1 && W == 8 * (w - 1) + 1, CWT_ERR_UNSUPPORTED,
                "iou: label size %dx%d is not 8*(%dx%d - 1) + 1", H, W, h, w);
    CWT_REQUIRE(w <= IOU_MAXW, CWT_ERR_UNSUPPORTED, "iou: w=%d > %d", w, IOU_MAXW);
    CWT_CUDA(cudaMemsetAsync(counts, 0, sizeof(long long) * 6 * (size_t)n, st));
    if (ce) CWT_CUDA(cudaMemsetAsync(ce, 0, sizeof(double) * 2 * (size_t)n, st));
    dim3 grid(h, n);
    auto* cnt = reinterpret_cast<unsigned long long*>(counts);
    if (label_kind == CWT_LABEL_I64)
        k_upsample_argmax_iou<true><<<grid, IOU_THREADS, 0, st>>>(l60, lab, maps_per_label, cnt, ce, h, w, H, W, ignore_index);
    else
        k_upsample_argmax_iou<false><<<grid, IOU_THREADS, 0, st>>>(l60, lab, maps_per_label, cnt, ce, h, w, H, W, ignore_index);
    CWT_LAUNCHED("upsample_argmax_iou");
    return CWT_OK;
}

// rows of variants flagged in normalize_mask are divided by max(|f_q[:,p]|, 1e-12)  (F.normalize)
__global__ void __launch_bounds__(256)
k_scale_by_inv_norm(float* __restrict__ l, const float* __restrict__ n2, int V, int HW, int normalize_mask,
                    size_t total) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int p = (int)(i % HW);
    const size_t row = i / HW;                 // (e*V + v)*2 + k
    const int v = (int)((row / 2) % V);
    const size_t e = row / (2 * (size_t)V);
    if ((normalize_mask >> v) & 1) l[i] = l[i] / fmaxf(sqrtf(n2[e * HW + p]), 1e-12f);
}

// generic intersection / union on integer predictions (src/util.py:280-308)
template <bool I64>
__global__ void __launch_bounds__(256)
k_intersection_union(const void* __restrict__ preds, const void* __restrict__ target, long long npix,
                     int num_classes, int ignore_index, unsigned long long* __restrict__ iot) {
    extern __shared__ int hist[];              // [num_classes][3] I, O, T
    const int n = blockIdx.y;
    for (int i = threadIdx.x; i < 3 * num_classes; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const size_t base = (size_t)n * (size_t)npix;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npix;
         i += (long long)gridDim.x * blockDim.x) {
        long long p, t;
        if (I64) { p = ((const long long*)preds)[base + i]; t = ((const long long*)target)[base + i]; }
        else     { p = ((const uint8_t*)preds)[base + i];   t = ((const uint8_t*)target)[base + i]; }
        if (t == ignore_index) p = ignore_index;            // preds[target == ignore] = ignore
        const bool pin = (p >= 0 && p < num_classes), tin = (t >= 0 && t < num_classes);
        if (pin && p == t) atomicAdd(&hist[(int)p * 3 + 0], 1);
        if (pin) atomicAdd(&hist[(int)p * 3 + 1], 1);       // histc drops values outside [0, C-1]
        if (tin) atomicAdd(&hist[(int)t * 3 + 2], 1);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * num_classes; i += blockDim.x)
        if (hist[i]) atomicAdd(&iot[(size_t)n * 3 * num_classes + i], (unsigned long long)hist[i]);
}

__global__ void k_finalize_union(unsigned long long* iot, size_t n_rows) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_rows) iot[i * 3 + 1] = iot[i * 3 + 1] + iot[i * 3 + 2] - iot[i * 3 + 0];
}

// training loss glue: zd = l1 - l0 ; d_logits = [-g60, +g60]
__global__ void k_logit_diff(const float* __restrict__ l, float* __restrict__ zd, int HW, size_t total) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    size_t e = i / HW; int p = (int)(i - e * HW);
    zd[i] = l[(e * 2 + 1) * HW + p] - l[(e * 2) * HW + p];
}
__global__ void k_spread_grad(const float* __restrict__ g60, float* __restrict__ dl, int HW, size_t total) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    size_t e = i / HW; int p = (int)(i - e * HW);
    const float g = g60[i];
    dl[(e * 2) * HW + p] = -g;
    dl[(e * 2 + 1) * HW + p] = g;
}

struct LogitsIouWs { float* l60; float* n2; };
static size_t carve_logits_iou(Carver& cv, LogitsIouWs& ws, int E, int V, int HW) {
    ws.l60 = cv.take<float>((size_t)E * V * 2 * HW);
    ws.n2 = cv.take<float>((size_t)E * HW);
    return align_up(cv.off);
}

struct QueryLossWs { uint4* cells; int32_t* counts; float2* cw; float* zd; float* g60; float* part; int nblk; };
static size_t carve_query_loss(Carver& cv, QueryLossWs& ws, int E, int h, int w, int H, int W) {
    ws.nblk = hires_bands(h);
    ws.cells = cv.take<uint4>((size_t)E * h * w);
    ws.counts = cv.take<int32_t>((size_t)E * 4);
    ws.cw = cv.take<float2>((size_t)E);
    ws.zd = cv.take<float>((size_t)E * h * w);
    ws.g60 = cv.take<float>((size_t)E * h * w);
    ws.part = cv.take<float>((size_t)E * ws.nblk);
    return align_up(cv.off);
}

}  // namespace cwt

using namespace cwt;

extern "C" size_t cwt_logits_iou_workspace_bytes(int E, int V, int C, int h, int w, int H, int W) {
    Carver cv(nullptr, 0);
    LogitsIouWs ws;
    return carve_logits_iou(cv, ws, E, V, h * w);
}

extern "C" int cwt_upsample_argmax_iou(const float* logits60, const void* label, int label_kind,
                                       long long* iu_counts, double* ce_or_null,
                                       int n, int h, int w, int H, int W, int ignore_index, void* stream) {
    if (n == 0) return CWT_OK;
    CWT_REQUIRE(logits60 && label && iu_counts && n > 0, CWT_ERR_INVALID_ARG, "upsample_argmax_iou: bad argument");
    CWT_REQUIRE(label_kind == CWT_LABEL_U8 || label_kind == CWT_LABEL_I64, CWT_ERR_INVALID_ARG, "label_kind %d", label_kind);
    return launch_upsample_iou(logits60, label, label_kind, 1, iu_counts, ce_or_null, n, h, w, H, W, ignore_index,
                               static_cast<cudaStream_t>(stream));
}

extern "C" int cwt_logits_iou(const float* wts, const float* f_q, const void* q_label, int label_kind,
                              int normalize_mask, long long* iu_counts, float* logits60_or_null,
                              double* ce_or_null, int E, int V, int C, int h, int w, int H, int W,
                              int ignore_index, void* workspace, size_t ws_bytes, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (E == 0) return CWT_OK;
    CWT_REQUIRE(wts && f_q && q_label && iu_counts, CWT_ERR_INVALID_ARG, "logits_iou: null pointer");
    CWT_REQUIRE(E > 0 && V >= 1 && V <= 8 && C >= 1, CWT_ERR_INVALID_ARG, "logits_iou: bad sizes E=%d V=%d C=%d", E, V, C);
    CWT_REQUIRE(label_kind == CWT_LABEL_U8 || label_kind == CWT_LABEL_I64, CWT_ERR_INVALID_ARG, "label_kind %d", label_kind);
    const int HWl = h * w;
    Carver cv(workspace, ws_bytes);
    LogitsIouWs ws;
    size_t need = carve_logits_iou(cv, ws, E, V, HWl);
    CWT_REQUIRE(workspace && ws_bytes >= need, CWT_ERR_WORKSPACE, "logits_iou: workspace %zu < %zu", ws_bytes, need);
    float* l60 = logits60_or_null ? logits60_or_null : ws.l60;
    int rc = launch_rows_times_feat(f_q, wts, l60, normalize_mask ? ws.n2 : nullptr, E, C, HWl, 1, V * 2, st);
    if (rc != CWT_OK) return rc;
    if (normalize_mask) {
        size_t total = (size_t)E * V * 2 * HWl;
        k_scale_by_inv_norm<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(l60, ws.n2, V, HWl, normalize_mask, total);
        CWT_LAUNCHED("scale_by_inv_norm");
    }
    return launch_upsample_iou(l60, q_label, label_kind, V, iu_counts, ce_or_null, E * V, h, w, H, W, ignore_index, st);
}

extern "C" int cwt_intersection_union(const void* preds, const void* target, int label_kind, long long* counts,
                                      int n, long long npix, int num_classes, int ignore_index, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n == 0) return CWT_OK;
    CWT_REQUIRE(preds && target && counts && n > 0 && npix >= 0, CWT_ERR_INVALID_ARG, "intersection_union: bad argument");
    CWT_REQUIRE(num_classes >= 1 && num_classes <= 1024, CWT_ERR_UNSUPPORTED, "intersection_union: num_classes %d", num_classes);
    CWT_REQUIRE(label_kind == CWT_LABEL_U8 || label_kind == CWT_LABEL_I64, CWT_ERR_INVALID_ARG, "label_kind %d", label_kind);
    CWT_CUDA(cudaMemsetAsync(counts, 0, sizeof(long long) * 3 * (size_t)num_classes * n, st));
    if (npix == 0) return CWT_OK;
    int bx = (int)((npix + 256 * 16 - 1) / (256 * 16));
    bx = bx < 1 ? 1 : (bx > 128 ? 128 : bx);
    dim3 grid(bx, n);
    auto* iot = reinterpret_cast<unsigned long long*>(counts);
    size_t sm = sizeof(int) * 3 * num_classes;
    if (label_kind == CWT_LABEL_I64)
        k_intersection_union<true><<<grid, 256, sm, st>>>(preds, target, npix, num_classes, ignore_index, iot);
    else
        k_intersection_union<false><<<grid, 256, sm, st>>>(preds, target, npix, num_classes, ignore_index, iot);
    CWT_LAUNCHED("intersection_union");
    size_t rows = (size_t)n * num_classes;
    k_finalize_union<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(iot, rows);
    CWT_LAUNCHED("finalize_union");
    return CWT_OK;
}

extern "C" size_t cwt_query_loss_workspace_bytes(int E, int h, int w, int H, int W) {
    Carver cv(nullptr, 0);
    QueryLossWs ws;
    return carve_query_loss(cv, ws, E, h, w, H, W);
}

extern "C" int cwt_query_loss_grad(const float* logits60, const void* label, int label_kind,
                                   float* loss, float* d_logits60, int E, int h, int w, int H, int W,
                                   int ignore_index, void* workspace, size_t ws_bytes, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (E == 0) return CWT_OK;
    CWT_REQUIRE(logits60 && label && loss && d_logits60 && E > 0, CWT_ERR_INVALID_ARG, "query_loss_grad: bad argument");
    CWT_REQUIRE(H == 8 * (h - 1) + 1 && W == 8 * (w - 1) + 1, CWT_ERR_UNSUPPORTED,
                "query_loss_grad: label size %dx%d is not 8*(%dx%d - 1) + 1", H, W, h, w);
    CWT_REQUIRE(w <= HIRES_MAXW, CWT_ERR_UNSUPPORTED, "query_loss_grad: w=%d > %d", w, HIRES_MAXW);
    CWT_REQUIRE(label_kind == CWT_LABEL_U8 || label_kind == CWT_LABEL_I64, CWT_ERR_INVALID_ARG, "label_kind %d", label_kind);
    Carver cv(workspace, ws_bytes);
    QueryLossWs ws;
    size_t need = carve_query_loss(cv, ws, E, h, w, H, W);
    CWT_REQUIRE(workspace && ws_bytes >= need, CWT_ERR_WORKSPACE, "query_loss_grad: workspace %zu < %zu", ws_bytes, need);
    const int HWl = h * w;
    int rc = pack_label_cells(label, label_kind, E, h, w, H, W, ignore_index, ws.cells, ws.counts, st);
    if (rc != CWT_OK) return rc;
    k_class_consts<<<(E + 127) / 128, 128, 0, st>>>(ws.counts, nullptr, 1e-12, ws.cw, nullptr, E, 1);
    CWT_LAUNCHED("class_consts");
    size_t total = (size_t)E * HWl;
    k_logit_diff<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(logits60, ws.zd, HWl, total);
    CWT_LAUNCHED("logit_diff");
    rc = launch_fit_hires<true>(ws.zd, ws.cells, ws.cw, ws.g60, ws.part, E, h, w, 1, st);
    if (rc != CWT_OK) return rc;
    k_reduce_loss<<<E, 32, 0, st>>>(ws.part, loss, ws.nblk);
    CWT_LAUNCHED("reduce_loss");
    k_spread_grad<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(ws.g60, d_logits60, HWl, total);
    CWT_LAUNCHED("spread_grad");
    return CWT_OK;
}
