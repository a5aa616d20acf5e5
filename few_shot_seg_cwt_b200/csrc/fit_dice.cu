// (f-3) PSPNet.inner_loop with SegLoss('wt_dc' | 'dc') — the per-channel sigmoid dice loss of the reference
// (src/model/pspnet.py:189-205 with criterion = SegLoss(loss_type), src/model/model_util.py:18-19 ->
// weighted_dice_loss, model_util.py:40-73: weighted_val 1, reduction 'sum', input_type 'lg').
//
// Math per SGD step, for every episode (S support images pooled into one loss, n = S):
//     L    = W . F                                both logit rows, [S, 2, h*w]                  (RTF<2>)
//     u    = up(L)   bilinear, align_corners, scale exactly 1/8, at H x W, per channel
//     p    = sigmoid(u)                           per channel (NOT a softmax over the two classes)
//     row (s,c):  A = sum_i p_i^2  (every pixel, 255 included) ; N = sum_i t_i p_i ; T = sum_i t_i^2 = #(label == c)
//                 D = clamp(A + T, 1e-8) ; loss_row = 1 - 2 N / D ; loss = sum_rows loss_row / S   (k_dice_hires<0>, k_dice_coef)
//     dl/dp_i = (-2 t_i / D + 4 N p_i / D^2) / S  ; du_i = dl/dp_i . p_i (1 - p_i)
//     g60  = up^T(du)                             adjoint as a deterministic gather             (k_dice_hires<1>)
//     dW_c = sum_s g60[s,c] . F[s]^T ; W_c -= lr dW_c                                           (FTC<2> + SGD epilogue)
// The two rows do not cancel here (two independent sigmoids), so both are contracted. Streaming algorithm only: five
// launches per step over the whole batch; the variant is not on the episodic hot path (every shipped config uses
// inner_loss_type wt_ce), it completes SegLoss for the inner_loop drop-in.
#include "common.cuh"
#include "skinny.cuh"
#include "hires.cuh"

namespace cwt {

int pack_label_cells(const void* labels, int label_kind, int n_img, int h, int w, int H, int W, int ignore_index,
                     uint4* cells, int32_t* counts, cudaStream_t st);

// One cell (64 pixels) of ONE logit channel. ch = the label code that is this channel's target (0 bg / 1 fg);
// nr / ns = rows / columns of the cell that lie inside the image (8, or 1 in the last cell row / column: H = 8(h-1)+1).
// PASS 0: a0 = sum p^2, a1 = sum t p over the in-image pixels. PASS 1: a0..a3 = gradient mass sent to the corners
// 00 01 10 11, with du = (alpha t + beta p) p (1 - p).
template <int PASS>
__device__ __forceinline__ void dice_cell(float z00, float z01, float z10, float z11, uint4 bits, uint32_t ch, int nr, int ns,
                                          float alpha, float beta, float& a0, float& a1, float& a2, float& a3) {
    const uint32_t words[4] = {bits.x, bits.y, bits.z, bits.w};
    const float dl = (z10 - z00) * 0.125f, dr = (z11 - z01) * 0.125f;
    float c00 = 0.f, c01 = 0.f, c10 = 0.f, c11 = 0.f, sp2 = 0.f, stp = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const uint32_t rb = (words[r >> 1] >> ((r & 1) * 16)) & 0xffffu;
        const float left = fmaf((float)r, dl, z00), right = fmaf((float)r, dr, z01);
        const float slope = (right - left) * 0.125f;
        const float nleft = NEG_LOG2E * left, nslope = NEG_LOG2E * slope;
        float gs = 0.f, gr = 0.f;
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            const float p = fast_rcp(1.f + fast_ex2(fmaf((float)s, nslope, nleft)));
            const bool valid = (r < nr) && (s < ns);
            const float t = (((rb >> (2 * s)) & 3u) == ch) ? 1.f : 0.f;
            if (PASS == 0) {
                if (valid) { sp2 = fmaf(p, p, sp2); stp = fmaf(t, p, stp); }
            } else {
                const float g = valid ? fmaf(alpha, t, beta * p) * (p * (1.f - p)) : 0.f;
                gs += g;
                gr = fmaf((float)s, g, gr);
            }
        }
        if (PASS == 1) {
            gr *= 0.125f;                         // mass sent to the right column
            const float gl = gs - gr;             // ... and to the left one
            const float h1 = (float)r * 0.125f, h0 = 1.f - h1;
            c00 = fmaf(h0, gl, c00); c01 = fmaf(h0, gr, c01);
            c10 = fmaf(h1, gl, c10); c11 = fmaf(h1, gr, c11);
        }
    }
    if (PASS == 0) { a0 = sp2; a1 = stp; a2 = 0.f; a3 = 0.f; }
    else { a0 = c00; a1 = c01; a2 = c10; a3 = c11; }
}

// CTA = (band of RROWS low-res rows, image, channel); same band / halo-row scheme as k_fit_hires (hires.cuh).
//   lg   [img][2][h*w]      logits of both rows (rows_times_feat layout)
//   coef [img*2 + c]        (alpha, beta) of the row, PASS 1
//   part [img*2 + c][band][2]   partial (sum p^2, sum t p), PASS 0 (each cell is owned by exactly one band)
//   g60  [e][2][S][h*w]     feat_times_cols layout, PASS 1
template <int RROWS, int PASS>
__global__ void __launch_bounds__((RROWS + 1) * HIRES_MAXW > 1024 ? 1024 : 512)
k_dice_hires(const float* __restrict__ lg, const uint4* __restrict__ cells, const float2* __restrict__ coef,
             float* __restrict__ g60, float* __restrict__ part, int h, int w, int S) {
    extern __shared__ float hsm[];
    float* zs = hsm;                                  // [(RROWS+2)][w]   rows a_first-1 .. a_last+1
    float* cc = zs + (RROWS + 2) * w;                 // [4][(RROWS+1)][w] corner contributions
    __shared__ float lred[2][32];
    const int img = blockIdx.y, c = blockIdx.z, e = img / S;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int a_first = blockIdx.x * RROWS;
    const int a_last = min(a_first + RROWS - 1, h - 1);
    const int row = img * 2 + c;
    const float* zimg = lg + (size_t)row * h * w;
    float alpha = 0.f, beta = 0.f;
    if (PASS == 1) { const float2 ab = coef[row]; alpha = ab.x; beta = ab.y; }
    for (int i = tid; i < (RROWS + 2) * w; i += nthr) {
        const int ar = i / w, b = i - ar * w;
        const int a = min(max(a_first - 1 + ar, 0), h - 1);       // clamped (out-of-range rows carry zero weight)
        zs[i] = zimg[a * w + b];
    }
    __syncthreads();
    float sp2 = 0.f, stp = 0.f;
    const uint4* cimg = cells + (size_t)img * h * w;
    for (int i = tid; i < (RROWS + 1) * w; i += nthr) {
        const int ar = i / w, b = i - ar * w;
        const int a = a_first - 1 + ar;
        float o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
        const bool active = (PASS == 0) ? (ar >= 1 && a <= a_last) : (a >= 0 && a <= a_last);
        if (active) {
            const int b1 = min(b + 1, w - 1);
            const float* z0 = zs + ar * w;
            const float* z1 = zs + (ar + 1) * w;
            dice_cell<PASS>(z0[b], z0[b1], z1[b], z1[b1], cimg[a * w + b], (uint32_t)c, (a < h - 1) ? 8 : 1, (b < w - 1) ? 8 : 1,
                            alpha, beta, o0, o1, o2, o3);
            if (PASS == 0) { sp2 += o0; stp += o1; }
        }
        if (PASS == 1) {
            cc[(0 * (RROWS + 1) + ar) * w + b] = o0;
            cc[(1 * (RROWS + 1) + ar) * w + b] = o1;
            cc[(2 * (RROWS + 1) + ar) * w + b] = o2;
            cc[(3 * (RROWS + 1) + ar) * w + b] = o3;
        }
    }
    if (PASS == 0) {
        sp2 = warp_sum(sp2); stp = warp_sum(stp);
        if ((tid & 31) == 0) { lred[0][tid >> 5] = sp2; lred[1][tid >> 5] = stp; }
        __syncthreads();
        if (tid == 0) {
            float s0 = 0.f, s1 = 0.f;
            for (int k = 0; k < (nthr + 31) / 32; ++k) { s0 += lred[0][k]; s1 += lred[1][k]; }      // fixed order
            part[((size_t)row * gridDim.x + blockIdx.x) * 2 + 0] = s0;
            part[((size_t)row * gridDim.x + blockIdx.x) * 2 + 1] = s1;
        }
        return;
    }
    __syncthreads();
    // g60(a,b) = c00(a,b) + c01(a,b-1) + c10(a-1,b) + c11(a-1,b-1)
    float* gout = g60 + (((size_t)e * 2 + c) * S + (img - e * S)) * h * w;
    for (int i = tid; i < RROWS * w; i += nthr) {
        const int ar = 1 + i / w, b = i - (ar - 1) * w;
        const int a = a_first - 1 + ar;
        if (a > a_last) break;
        float s = cc[(0 * (RROWS + 1) + ar) * w + b] + cc[(2 * (RROWS + 1) + ar - 1) * w + b];
        if (b > 0) s += cc[(1 * (RROWS + 1) + ar) * w + b - 1] + cc[(3 * (RROWS + 1) + ar - 1) * w + b - 1];
        gout[a * w + b] = s;
    }
}

// one thread per (image, channel) row: dice denominators -> gradient coefficients and the row's loss share
__global__ void __launch_bounds__(128)
k_dice_coef(const float* __restrict__ part, const int32_t* __restrict__ counts_img, float2* __restrict__ coef,
            float* __restrict__ lossrow, int n_rows, int nb, int S) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows) return;
    float A = 0.f, N = 0.f;
    for (int b = 0; b < nb; ++b) { A += part[((size_t)i * nb + b) * 2]; N += part[((size_t)i * nb + b) * 2 + 1]; }
    const float T = (float)counts_img[(i >> 1) * 4 + (i & 1)];
    const float lp = A + T;
    const bool clamped = lp < 1e-8f;                  // torch.clamp(min=eps): no gradient through the clamped branch
    const float D = clamped ? 1e-8f : lp;
    const float invS = 1.f / (float)S;
    coef[i] = make_float2(-2.f / D * invS, clamped ? 0.f : 4.f * N / (D * D) * invS);
    lossrow[i] = (1.f - 2.f * N / D) * invS;
}

// FTC epilogue: complete dW[e][r][c] -> plain SGD on both rows
struct DiceUpdateEpilogue {
    float* W; float lr; int C;
    __device__ __forceinline__ void operator()(int e, int r, int c, float dw) const {
        if (r >= 2) return;
        float* p = W + (size_t)(e * 2 + r) * C + c;
        *p = fmaf(-lr, dw, *p);
    }
};

struct DiceWs {
    uint4* cells; int32_t* counts_img; float* lg; float* g60; float* part; float2* coef; float* lossrow; int nblk;
};

static size_t carve_dice(Carver& cv, DiceWs& ws, int E, int S, int h, int w) {
    ws.nblk = hires_bands(h);
    ws.cells = cv.take<uint4>((size_t)E * S * h * w);
    ws.counts_img = cv.take<int32_t>((size_t)E * S * 4);
    ws.lg = cv.take<float>((size_t)E * S * 2 * h * w);
    ws.g60 = cv.take<float>((size_t)E * S * 2 * h * w);
    ws.part = cv.take<float>((size_t)E * S * 2 * ws.nblk * 2);
    ws.coef = cv.take<float2>((size_t)E * S * 2);
    ws.lossrow = cv.take<float>((size_t)E * S * 2);
    return align_up(cv.off);
}

}  // namespace cwt

using namespace cwt;

extern "C" size_t cwt_fit_dice_workspace_bytes(int E, int S, int C, int h, int w, int H, int W) {
    (void)C; (void)H; (void)W;
    Carver cv(nullptr, 0);
    DiceWs ws;
    return carve_dice(cv, ws, E, S, h, w);
}

extern "C" int cwt_fit_classifier_dice_f32(const float* f_s, const void* s_label, int label_kind, const float* w0,
                                           float* w_out, float* loss_trace_or_null,
                                           int E, int S, int C, int h, int w, int H, int W,
                                           int n_iter, float lr, int ignore_index,
                                           void* workspace, size_t ws_bytes, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CWT_REQUIRE(E >= 0 && S >= 1 && C >= 1 && h >= 1 && w >= 1 && n_iter >= 0, CWT_ERR_INVALID_ARG,
                "fit_dice: bad sizes E=%d S=%d C=%d h=%d w=%d n_iter=%d", E, S, C, h, w, n_iter);
    if (E == 0) return CWT_OK;
    CWT_REQUIRE(f_s && s_label && w0 && w_out, CWT_ERR_INVALID_ARG, "fit_dice: null pointer");
    CWT_REQUIRE(H == 8 * (h - 1) + 1 && W == 8 * (w - 1) + 1, CWT_ERR_UNSUPPORTED,
                "fit_dice: label size %dx%d is not 8*(%dx%d - 1) + 1 (align_corners scale must be 1/8)", H, W, h, w);
    CWT_REQUIRE(w <= HIRES_MAXW, CWT_ERR_UNSUPPORTED, "fit_dice: w=%d exceeds the supported width (%d)", w, HIRES_MAXW);
    CWT_REQUIRE(label_kind == CWT_LABEL_U8 || label_kind == CWT_LABEL_I64, CWT_ERR_INVALID_ARG,
                "fit_dice: label_kind %d", label_kind);
    Carver cv(workspace, ws_bytes);
    DiceWs ws;
    const size_t need = carve_dice(cv, ws, E, S, h, w);
    CWT_REQUIRE(workspace && ws_bytes >= need, CWT_ERR_WORKSPACE, "fit_dice: workspace %zu < %zu bytes", ws_bytes, need);

    const int HWl = h * w, n_img = E * S;
    int rc = pack_label_cells(s_label, label_kind, n_img, h, w, H, W, ignore_index, ws.cells, ws.counts_img, st);
    if (rc != CWT_OK) return rc;
    if (w_out != w0) CWT_CUDA(cudaMemcpyAsync(w_out, w0, sizeof(float) * 2 * (size_t)E * C, cudaMemcpyDeviceToDevice, st));

    const dim3 grid(ws.nblk, n_img, 2);
    const int thr = hires_threads(w);
    const size_t sm = hires_smem(w);
    DiceUpdateEpilogue epi{w_out, lr, C};
    for (int it = 0; it < n_iter; ++it) {
        rc = launch_rows_times_feat(f_s, w_out, ws.lg, nullptr, n_img, C, HWl, S, 2, st);
        if (rc != CWT_OK) return rc;
        k_dice_hires<HIRES_RROWS, 0><<<grid, thr, sm, st>>>(ws.lg, ws.cells, nullptr, nullptr, ws.part, h, w, S);
        CWT_LAUNCHED("dice_hires<0>");
        k_dice_coef<<<(n_img * 2 + 127) / 128, 128, 0, st>>>(ws.part, ws.counts_img, ws.coef, ws.lossrow, n_img * 2, ws.nblk, S);
        CWT_LAUNCHED("dice_coef");
        if (loss_trace_or_null) {
            k_reduce_loss<<<E, 32, 0, st>>>(ws.lossrow, loss_trace_or_null + (size_t)it * E, 2 * S);
            CWT_LAUNCHED("reduce_loss");
        }
        k_dice_hires<HIRES_RROWS, 1><<<grid, thr, sm, st>>>(ws.lg, ws.cells, ws.coef, ws.g60, nullptr, h, w, S);
        CWT_LAUNCHED("dice_hires<1>");
        rc = launch_ftc_t<2, 4>(f_s, ws.g60, E, S, C, HWl, 2, epi, st);
        if (rc != CWT_OK) return rc;
    }
    return CWT_OK;
}
