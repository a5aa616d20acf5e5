// Full-resolution (H x W) stage shared by the support fit (a-3) and the query training
// loss (a-13): bilinear up (scale exactly 1/8) of a logit difference, 2-class softmax as a
// sigmoid, class-weighted CE gradient, and the adjoint bilinear map as a deterministic
// gather (the reference's upsample_bilinear2d_backward uses float atomics on CUDA).
#pragma once
#include "common.cuh"

namespace cwt {

// Per-episode CE constants: class weight [wt0, wt1] (given, or [1, n0/(n1+eps)] counted over the
// episode's S label images: src/test.py:169-175 with eps=0, src/train.py:237-243 with eps=1e-12;
// python float64 division, stored fp32) pre-divided by sum_i w[y_i]  (= n0*wt0 + n1*wt1).
static __global__ void __launch_bounds__(128)
k_class_consts(const int32_t* __restrict__ counts_img, const float* __restrict__ class_weight, double eps,
               float2* __restrict__ cw, int32_t* __restrict__ counts_ep, int E, int S) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    long long n[4] = {0, 0, 0, 0};
    for (int s = 0; s < S; ++s)
#pragma unroll
        for (int k = 0; k < 4; ++k) n[k] += counts_img[(e * S + s) * 4 + k];
    float wt0 = 1.f, wt1;
    if (class_weight) { wt0 = class_weight[e * 2]; wt1 = class_weight[e * 2 + 1]; }
    else wt1 = (float)((double)n[0] / ((double)n[1] + eps));
    const float sumw = (float)((double)n[0] * (double)wt0 + (double)n[1] * (double)wt1);
    cw[e] = make_float2(wt0 / sumw, wt1 / sumw);
    if (counts_ep)
#pragma unroll
        for (int k = 0; k < 4; ++k) counts_ep[e * 4 + k] = (int32_t)n[k];
}

// Full-resolution stage: CTA = (band of RROWS low-res rows, image). Recomputes the sigmoid of
// the hi-res rows it gathers from (8 halo rows per band: (RROWS+1)/RROWS redundancy).
constexpr int HIRES_THREADS = 256;
constexpr int HIRES_MAXCOL = 4;      // W <= 1024
constexpr int HIRES_MAXW = 128;      // w <= 128

template <int RROWS, bool LOSS>
__global__ void __launch_bounds__(HIRES_THREADS)
k_fit_hires(const float* __restrict__ zd, const uint8_t* __restrict__ lab, const float2* __restrict__ cw,
            float* __restrict__ g60, float* __restrict__ loss_part, int h, int w, int H, int W, int S) {
    __shared__ float zs[(RROWS + 2) * HIRES_MAXW];
    __shared__ float colbuf[HIRES_THREADS * HIRES_MAXCOL];
    __shared__ float lred[HIRES_THREADS / 32];
    const int img = blockIdx.y, e = img / S;
    const int tid = threadIdx.x;
    const int a_first = blockIdx.x * RROWS;
    const int a_last = min(a_first + RROWS - 1, h - 1);
    const float2 c01 = cw[e];
    const float* zimg = zd + (size_t)img * h * w;
    for (int i = tid; i < (RROWS + 2) * w; i += HIRES_THREADS) {
        int ar = i / w, b = i - ar * w;
        int a = a_first - 1 + ar;
        zs[ar * w + b] = (a >= 0 && a < h) ? zimg[a * w + b] : 0.f;
    }
    __syncthreads();

    const uint8_t* limg = lab + (size_t)img * H * W;
    float acc_cur[HIRES_MAXCOL], acc_next[HIRES_MAXCOL];
#pragma unroll
    for (int j = 0; j < HIRES_MAXCOL; ++j) { acc_cur[j] = 0.f; acc_next[j] = 0.f; }
    float loss = 0.f;
    const int ncol = (W + HIRES_THREADS - 1) / HIRES_THREADS;

    for (int a0 = a_first - 1; a0 <= a_last; ++a0) {
        if (a0 >= 0) {
            const int a1 = min(a0 + 1, h - 1);
            const float* z0 = zs + (a0 - a_first + 1) * w;
            const float* z1 = zs + (a1 - a_first + 1) * w;
            const bool need_cur = (a0 >= a_first);
            for (int r = 0; r < 8; ++r) {
                const int Y = 8 * a0 + r;
                if (Y >= H) break;
                const bool need_next = (r > 0) && (a0 + 1 <= a_last);
                if (!need_cur && !need_next) continue;
                const float h1 = r * 0.125f, h0 = 1.f - h1;
#pragma unroll
                for (int j = 0; j < HIRES_MAXCOL; ++j) {
                    const int X = tid + j * HIRES_THREADS;
                    if (j < ncol && X < W) {
                        const int b0 = X >> 3, b1 = min(b0 + 1, w - 1);
                        const float w1 = (X & 7) * 0.125f, w0 = 1.f - w1;
                        const float d = bilerp8(z0[b0], z0[b1], z1[b0], z1[b1], w0, w1, h0, h1);
                        const int code = limg[(size_t)Y * W + X];
                        const float p = __fdividef(1.f, 1.f + __expf(-d));
                        const float g = (code == 0) ? c01.x * p : ((code == 1) ? c01.y * (p - 1.f) : 0.f);
                        acc_cur[j] = fmaf(h0, g, acc_cur[j]);
                        acc_next[j] = fmaf(h1, g, acc_next[j]);
                        if (LOSS && need_cur && code < 2) {
                            const float t = (code == 1) ? -d : d;          // -log p_y = softplus(t)
                            const float sp = fmaxf(t, 0.f) + log1pf(expf(-fabsf(t)));
                            loss = fmaf((code == 1) ? c01.y : c01.x, sp, loss);
                        }
                    }
                }
            }
        }
        if (a0 >= a_first) {
#pragma unroll
            for (int j = 0; j < HIRES_MAXCOL; ++j) {
                const int X = tid + j * HIRES_THREADS;
                if (j < ncol && X < W) colbuf[X] = acc_cur[j];
            }
            __syncthreads();
            for (int b = tid; b < w; b += HIRES_THREADS) {
                float s = 0.f;
#pragma unroll
                for (int k = -7; k <= 7; ++k) {
                    const int X = 8 * b + k;
                    if (X >= 0 && X < W) s = fmaf(1.f - fabsf((float)k) * 0.125f, colbuf[X], s);
                }
                g60[(size_t)img * h * w + a0 * w + b] = s;
            }
            __syncthreads();
        }
#pragma unroll
        for (int j = 0; j < HIRES_MAXCOL; ++j) { acc_cur[j] = acc_next[j]; acc_next[j] = 0.f; }
    }
    if (LOSS) {
        loss = warp_sum(loss);
        if ((tid & 31) == 0) lred[tid >> 5] = loss;
        __syncthreads();
        if (tid == 0) {
            float s = 0.f;
            for (int k = 0; k < HIRES_THREADS / 32; ++k) s += lred[k];
            loss_part[(size_t)e * (S * gridDim.x) + (img - e * S) * gridDim.x + blockIdx.x] = s;
        }
    }
}

static __global__ void k_reduce_loss(const float* __restrict__ part, float* __restrict__ out, int n_per_ep) {
    const int e = blockIdx.x;
    float s = 0.f;
    for (int i = threadIdx.x; i < n_per_ep; i += 32) s += part[(size_t)e * n_per_ep + i];
    s = warp_sum(s);
    if (threadIdx.x == 0) out[e] = s;
}


constexpr int HIRES_RROWS = 4;

}  // namespace cwt
