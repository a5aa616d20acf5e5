// Full-resolution (H x W) stage shared by the support fit (a-3) and the query training loss (a-13):
// bilinear up (align_corners, scale exactly 1/8) of a 2-class logit difference, softmax as a
// sigmoid, class-weighted CE gradient, and the adjoint bilinear map.
//
// Work unit = one CELL of the align_corners grid: cell (a,b) covers pixels [8a,8a+8) x [8b,8b+8),
// is interpolated from the four low-res corners z(a,b) z(a,b+1) z(a+1,b) z(a+1,b+1) and sends its
// gradient back to exactly those four corners. A thread evaluates one cell (64 sigmoids, all in
// registers, labels as 8 x uint16 = one 128-bit load); the four corner contributions are then
// combined by a deterministic gather (the reference's upsample_bilinear2d_backward uses float
// atomics on CUDA and is run-to-run non-deterministic).
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


// sigmoid pieces as single MUFU ops (no denormal fix-up code, no branches)
__device__ __forceinline__ float fast_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// One pixel of the CE gradient, branch-free. nd = -log2(e) * d (pre-scaled logit difference);
// sel = the row's label bits masked to this pixel's 2-bit field, one = 1 << field, c0/c1 = w/sum_w.
//   code 0: g = c0 * p ; code 1: g = c1 * (p - 1) ; ignored / invalid: 0        (p = sigmoid(d))
__device__ __forceinline__ float ce_grad_pixel(float nd, uint32_t sel, uint32_t one, float c0, float c1) {
    const float p = fast_rcp(1.f + fast_ex2(nd));
    // selects on integer masks (kept branch-free: label codes differ from lane to lane)
    const uint32_t bB = (sel == one) ? __float_as_uint(c1) : 0u;
    const uint32_t bA = (sel == 0u) ? __float_as_uint(c0) : bB;
    const float A = __uint_as_float(bA), B = __uint_as_float(bB);
    return fmaf(A, p, -B);
}
constexpr float NEG_LOG2E = -1.4426950408889634f;

// One cell: 64 pixels. bits = 8 rows x 16 bit (2 bit label code per pixel, see k_pack_label_cells).
// c0 / c1 = w[0]/sum_w, w[1]/sum_w. Outputs the gradient mass sent to the four corners.
template <bool LOSS>
__device__ __forceinline__ void hires_cell(float z00, float z01, float z10, float z11, uint4 bits,
                                           float c0, float c1, float& o00, float& o01, float& o10, float& o11,
                                           float& loss) {
    const uint32_t words[4] = {bits.x, bits.y, bits.z, bits.w};
    const float dl = (z10 - z00) * 0.125f, dr = (z11 - z01) * 0.125f;
    float a00 = 0.f, a01 = 0.f, a10 = 0.f, a11 = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const uint32_t rb = (words[r >> 1] >> ((r & 1) * 16)) & 0xffffu;
        const float left = fmaf((float)r, dl, z00), right = fmaf((float)r, dr, z01);
        const float slope = (right - left) * 0.125f;
        float gs = 0.f, gr = 0.f;
        const float nleft = NEG_LOG2E * left, nslope = NEG_LOG2E * slope;
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            const float g = ce_grad_pixel(fmaf((float)s, nslope, nleft), rb & (3u << (2 * s)), 1u << (2 * s), c0, c1);
            gs += g;
            gr = fmaf((float)s, g, gr);
            if (LOSS) {
                const uint32_t code = (rb >> (2 * s)) & 3u;
                if (code < 2u) {
                    const float d = fmaf((float)s, slope, left);
                    const float t = (code == 1u) ? -d : d;                 // -log p_y = softplus(t)
                    loss = fmaf((code == 1u) ? c1 : c0, fmaxf(t, 0.f) + log1pf(expf(-fabsf(t))), loss);
                }
            }
        }
        gr *= 0.125f;                         // mass sent to the right column
        const float gl = gs - gr;             // ... and to the left one
        const float h1 = (float)r * 0.125f, h0 = 1.f - h1;
        a00 = fmaf(h0, gl, a00); a01 = fmaf(h0, gr, a01);
        a10 = fmaf(h1, gl, a10); a11 = fmaf(h1, gr, a11);
    }
    o00 = a00; o01 = a01; o10 = a10; o11 = a11;
}

// Streaming form: CTA = (band of RROWS low-res rows, image); one thread per cell of the band plus the
// cell row above it (the halo row is recomputed: (RROWS+1)/RROWS redundancy, no cross-CTA traffic).
constexpr int HIRES_MAXW = 128;      // w <= 128
constexpr int HIRES_RROWS = 4;

template <int RROWS, bool LOSS>
__global__ void __launch_bounds__((RROWS + 1) * HIRES_MAXW > 1024 ? 1024 : 512)
k_fit_hires(const float* __restrict__ zd, const uint4* __restrict__ cells, const float2* __restrict__ cw,
            float* __restrict__ g60, float* __restrict__ loss_part, int h, int w, int S,
            const float* __restrict__ zbias) {        // zbias (nullable) [E]: classifier bias difference added to z
    extern __shared__ float hsm[];
    float* zs = hsm;                                  // [(RROWS+2)][w]   rows a_first-1 .. a_last+1
    float* cc = zs + (RROWS + 2) * w;                 // [4][(RROWS+1)][w] corner contributions
    __shared__ float lred[32];
    const int img = blockIdx.y, e = img / S;
    const float zb = zbias ? zbias[e] : 0.f;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int a_first = blockIdx.x * RROWS;
    const int a_last = min(a_first + RROWS - 1, h - 1);
    const float2 c01 = cw[e];
    const float* zimg = zd + (size_t)img * h * w;
    for (int i = tid; i < (RROWS + 2) * w; i += nthr) {
        const int ar = i / w, b = i - ar * w;
        const int a = min(max(a_first - 1 + ar, 0), h - 1);       // clamped (out-of-range rows carry zero weight)
        zs[i] = zimg[a * w + b] + zb;
    }
    __syncthreads();
    float loss = 0.f;
    const uint4* cimg = cells + (size_t)img * h * w;
    for (int i = tid; i < (RROWS + 1) * w; i += nthr) {
        const int ar = i / w, b = i - ar * w;
        const int a = a_first - 1 + ar;
        float o00 = 0.f, o01 = 0.f, o10 = 0.f, o11 = 0.f;
        if (a >= 0 && a <= a_last) {
            const int b1 = min(b + 1, w - 1);
            const float* z0 = zs + ar * w;
            const float* z1 = zs + (ar + 1) * w;
            float l = 0.f;
            hires_cell<LOSS>(z0[b], z0[b1], z1[b], z1[b1], cimg[a * w + b], c01.x, c01.y, o00, o01, o10, o11, l);
            if (LOSS && ar >= 1) loss += l;                       // each cell is owned by exactly one band
        }
        cc[(0 * (RROWS + 1) + ar) * w + b] = o00;
        cc[(1 * (RROWS + 1) + ar) * w + b] = o01;
        cc[(2 * (RROWS + 1) + ar) * w + b] = o10;
        cc[(3 * (RROWS + 1) + ar) * w + b] = o11;
    }
    __syncthreads();
    // g60(a,b) = c00(a,b) + c01(a,b-1) + c10(a-1,b) + c11(a-1,b-1)
    for (int i = tid; i < RROWS * w; i += nthr) {
        const int ar = 1 + i / w, b = i - (ar - 1) * w;
        const int a = a_first - 1 + ar;
        if (a > a_last) break;
        float s = cc[(0 * (RROWS + 1) + ar) * w + b] + cc[(2 * (RROWS + 1) + ar - 1) * w + b];
        if (b > 0) s += cc[(1 * (RROWS + 1) + ar) * w + b - 1] + cc[(3 * (RROWS + 1) + ar - 1) * w + b - 1];
        g60[(size_t)img * h * w + a * w + b] = s;
    }
    if (LOSS) {
        loss = warp_sum(loss);
        if ((tid & 31) == 0) lred[tid >> 5] = loss;
        __syncthreads();
        if (tid == 0) {
            float s = 0.f;
            for (int k = 0; k < (nthr + 31) / 32; ++k) s += lred[k];
            loss_part[(size_t)e * (S * gridDim.x) + (img - e * S) * gridDim.x + blockIdx.x] = s;
        }
    }
}

static inline int hires_threads(int w) {
    int t = (HIRES_RROWS + 1) * w;
    t = (t + 31) / 32 * 32;
    return t > 512 ? 512 : t;
}
static inline size_t hires_smem(int w) { return sizeof(float) * ((HIRES_RROWS + 2) * w + 4 * (HIRES_RROWS + 1) * w); }
static inline int hires_bands(int h) { return (h + HIRES_RROWS - 1) / HIRES_RROWS; }

template <bool LOSS>
static int launch_fit_hires(const float* zd, const uint4* cells, const float2* cw, float* g60, float* loss_part,
                            int n_img, int h, int w, int S, cudaStream_t st, const float* zbias = nullptr) {
    dim3 grid(hires_bands(h), n_img);
    k_fit_hires<HIRES_RROWS, LOSS><<<grid, hires_threads(w), hires_smem(w), st>>>(zd, cells, cw, g60, loss_part, h, w, S, zbias);
    CWT_LAUNCHED("fit_hires");
    return CWT_OK;
}

static __global__ void k_reduce_loss(const float* __restrict__ part, float* __restrict__ out, int n_per_ep) {
    const int e = blockIdx.x;
    float s = 0.f;
    for (int i = threadIdx.x; i < n_per_ep; i += 32) s += part[(size_t)e * n_per_ep + i];
    s = warp_sum(s);
    if (threadIdx.x == 0) out[e] = s;
}

}  // namespace cwt
