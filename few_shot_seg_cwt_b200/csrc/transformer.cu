// (a-5, a-6, a-13) Classifier Weight Transformer — MultiHeadAttentionOne forward / backward.
// Replaces src/model/transformer.py:12-83 for the call transformer(W, f_q, f_q) (k is v, one
// shared projection), src/test.py:194-197 and src/train.py:250-257.
//
// Notation: X = normalised query features [HW,C] (stored as k[c][p] * inv_n[p]),
// A = w_qkvs.weight [nH*C, C], A_h its h-th row block, tau = sqrt(C), rows r = l*nH + h.
//
//   Q   = q A^T                                   small GEMM
//   QA  = Q_h A_h            [E,R,C]               small GEMM      (re-association: S_h = Q_h K_h^T = (Q_h A_h) X^T)
//   S   = QA X^T / tau       [E,R,HW]              rows_times_feat (CWT_ATTN_REASSOC)
//                                                  or the tcgen05 K-projection GEMM with the score epilogue (CWT_ATTN_TCGEN05)
//   P   = softmax_HW(S) ; Pt = P*keep/(1-p_attn)
//   aX  = Pt X               [E,R,C]               feat_times_cols (O_h = Pt K_h = (Pt X) A_h^T)
//   O   = aX_h A_h^T         [E*Lq, nH*C]          small GEMM
//   Z   = O Fc^T + b ; Rr = Z*keep_out/(1-p_out) + q ; Y = LayerNorm(Rr)
//
// The backward uses the same re-association (SURVEY.md §8 math block): no dense HW x C x C GEMM.
#include "common.cuh"
#include "skinny.cuh"
#include "skinny_stream.cuh"

namespace cwt {

int kproj_scores_tcgen05(const float* k, const float* w_qkvs, const float* Qp, float* sraw, float* n2,
                         int E, int Lq, int nH, int C, int HW, void* ws, size_t ws_bytes, cudaStream_t st);
size_t kproj_tcgen05_workspace_bytes(int E, int Lq, int nH, int C, int HW);

// ------------------------------------------------------------------------------------
// small strided SGEMM: C[m][n] (+)= alpha * sum_k A(m,k) B(k,n) (+ bias[n]); batch = gridDim.z
// ------------------------------------------------------------------------------------
struct GemmP {
    const float* A; const float* B; float* C; const float* bias;
    int M, N, K;
    long long a_ms, a_ks, b_ks, b_ns, ldc, a_bs, b_bs, c_bs;
    float alpha; int accumulate;
};

constexpr int G_TM = 32, G_TN = 32, G_TK = 16, G_KG = 4;

// The GEMMs of this path are small (M = 2 E rows, N, K = 512 .. 2048): at most ~128 CTAs, one per SM, so a CTA's serial
// K loop is what takes the time. The CTA therefore splits K over G_KG = 4 groups of 64 threads: group g multiplies the
// k-tiles g, g + 4, ... of the same 32 x 32 output tile (own shared-memory tiles, own named barrier, next tile prefetched
// into registers), thread = 4 x 4 outputs; the four partial tiles are added in a fixed order.
// (Measured at E = 64, whole transformer block: one group per CTA and 32 x 64 tiles 0.352 ms; 64 x 64 tiles with half the CTAs
// 0.392 ms; four groups 0.314 ms; four groups and 32 x 32 tiles — this — 0.298 ms; eight groups 0.315 ms.)
__global__ void __launch_bounds__(64 * G_KG) k_sgemm_small(GemmP p) {
    __shared__ __align__(16) float smem[G_KG * G_TK * (G_TM + 4 + G_TN + 4)];
    const int tid = threadIdx.x, g = tid >> 6, gt = tid & 63, tx = gt & 7, ty = gt >> 3;
    float* As = smem + g * (G_TK * (G_TM + 4 + G_TN + 4));        // [G_TK][G_TM + 4]
    float* Bs = As + G_TK * (G_TM + 4);                           // [G_TK][G_TN + 4]
    const int m0 = blockIdx.y * G_TM, n0 = blockIdx.x * G_TN;
    const float* A = p.A + (size_t)blockIdx.z * p.a_bs;
    const float* B = p.B + (size_t)blockIdx.z * p.b_bs;
    float* C = p.C + (size_t)blockIdx.z * p.c_bs;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    float ra[8], rb[8];
    auto a_idx = [&](int i, int& m, int& k) {
        const int idx = gt + i * 64;
        if (p.a_ks == 1) { k = idx % G_TK; m = idx / G_TK; } else { m = idx % G_TM; k = idx / G_TM; }
    };
    auto b_idx = [&](int i, int& n, int& k) {
        const int idx = gt + i * 64;
        if (p.b_ns == 1) { n = idx % G_TN; k = idx / G_TN; } else { k = idx % G_TK; n = idx / G_TK; }
    };
    auto fetch = [&](int k0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int m, k;
            a_idx(i, m, k);
            ra[i] = (m0 + m < p.M && k0 + k < p.K) ? A[(size_t)(m0 + m) * p.a_ms + (size_t)(k0 + k) * p.a_ks] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int n, k;
            b_idx(i, n, k);
            rb[i] = (n0 + n < p.N && k0 + k < p.K) ? B[(size_t)(k0 + k) * p.b_ks + (size_t)(n0 + n) * p.b_ns] : 0.f;
        }
    };
    const int kstep = G_TK * G_KG;
    if (g * G_TK < p.K) fetch(g * G_TK);
    for (int k0 = g * G_TK; k0 < p.K; k0 += kstep) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { int m, k; a_idx(i, m, k); As[k * (G_TM + 4) + m] = ra[i]; }
#pragma unroll
        for (int i = 0; i < 8; ++i) { int n, k; b_idx(i, n, k); Bs[k * (G_TN + 4) + n] = rb[i]; }
        asm volatile("bar.sync %0, 64;" ::"r"(1 + g) : "memory");
        if (k0 + kstep < p.K) fetch(k0 + kstep);
#pragma unroll
        for (int k = 0; k < G_TK; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(&As[k * (G_TM + 4) + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k * (G_TN + 4) + tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        asm volatile("bar.sync %0, 64;" ::"r"(1 + g) : "memory");
    }
    // groups 1..3 park their partial tiles in shared memory; group 0 adds them in order and stores
    __syncthreads();
    float* part = smem;                                           // [G_KG - 1][G_TM][G_TN]  (12 KB <= the tile storage)
    if (g > 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) part[((g - 1) * G_TM + ty * 4 + i) * G_TN + tx * 4 + j] = acc[i][j];
    }
    __syncthreads();
    if (g > 0) return;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= p.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= p.N) continue;
            float v = acc[i][j];
#pragma unroll
            for (int g2 = 0; g2 < G_KG - 1; ++g2) v += part[(g2 * G_TM + ty * 4 + i) * G_TN + tx * 4 + j];
            v *= p.alpha;
            if (p.bias) v += p.bias[n];
            float* dst = C + (size_t)m * p.ldc + n;
            *dst = p.accumulate ? (*dst + v) : v;
        }
    }
}

static int gemm(const float* A, long long a_ms, long long a_ks, const float* B, long long b_ks, long long b_ns,
                float* C, long long ldc, int M, int N, int K, float alpha, bool accumulate, const float* bias,
                int batch, long long a_bs, long long b_bs, long long c_bs, cudaStream_t st) {
    GemmP p{A, B, C, bias, M, N, K, a_ms, a_ks, b_ks, b_ns, ldc, a_bs, b_bs, c_bs, alpha, accumulate ? 1 : 0};
    dim3 grid((N + G_TN - 1) / G_TN, (M + G_TM - 1) / G_TM, batch);
    k_sgemm_small<<<grid, 64 * G_KG, 0, st>>>(p);
    CWT_LAUNCHED("sgemm_small");
    return CWT_OK;
}

// ------------------------------------------------------------------------------------
// softmax over HW of one score row; writes P (kept for backward) and Pm = P*keep/(1-p)*inv_n
// grid (R, E), rows r = l*nH + h ; keep_attn [(h*E + e)*Lq + l][HW]
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_softmax_rows(const float* __restrict__ sraw, const float* __restrict__ n2, const uint8_t* __restrict__ keep,
               float inv_tau, float drop_scale, float* __restrict__ P, float* __restrict__ Pm,
               float* __restrict__ inv_n_out, int E, int Lq, int nH, int HW) {
    __shared__ float red[8];
    __shared__ float bc;
    const int r = blockIdx.x, e = blockIdx.y, l = r / nH, hh = r - l * nH;
    const int tid = threadIdx.x;
    const size_t row = ((size_t)e * (Lq * nH) + r) * HW;
    const float* s = sraw + row;
    const float* nn = n2 ? n2 + (size_t)e * HW : nullptr;
    const uint8_t* kp = keep ? keep + (((size_t)hh * E + e) * Lq + l) * HW : nullptr;
    float mx = -INFINITY;
    for (int p = tid; p < HW; p += 256) {
        float in = nn ? 1.f / fmaxf(sqrtf(nn[p]), 1e-12f) : 1.f;
        mx = fmaxf(mx, s[p] * in * inv_tau);
    }
    mx = warp_max(mx);
    if ((tid & 31) == 0) red[tid >> 5] = mx;
    __syncthreads();
    if (tid == 0) { float m = red[0]; for (int k = 1; k < 8; ++k) m = fmaxf(m, red[k]); bc = m; }
    __syncthreads();
    mx = bc;
    float sum = 0.f;
    for (int p = tid; p < HW; p += 256) {
        float in = nn ? 1.f / fmaxf(sqrtf(nn[p]), 1e-12f) : 1.f;
        sum += expf(s[p] * in * inv_tau - mx);
    }
    sum = warp_sum(sum);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = sum;
    __syncthreads();
    if (tid == 0) { float t = 0.f; for (int k = 0; k < 8; ++k) t += red[k]; bc = t; }
    __syncthreads();
    const float inv_sum = 1.f / bc;
    for (int p = tid; p < HW; p += 256) {
        float in = nn ? 1.f / fmaxf(sqrtf(nn[p]), 1e-12f) : 1.f;
        float pr = expf(s[p] * in * inv_tau - mx) * inv_sum;
        if (P) P[row + p] = pr;
        float pt = kp ? (kp[p] ? pr * drop_scale : 0.f) : pr;
        Pm[row + p] = pt * in;
        if (inv_n_out && r == 0) inv_n_out[(size_t)e * HW + p] = in;
    }
}

// Rr = Z*keep/(1-p) + q ; LayerNorm over C (eps 1e-5, biased variance)   grid (rows)
__global__ void __launch_bounds__(128)
k_residual_layernorm(const float* __restrict__ Z, int nparts, size_t part_stride, const float* __restrict__ zbias,
                     const float* __restrict__ q, const uint8_t* __restrict__ keep,
                     float drop_scale, const float* __restrict__ g, const float* __restrict__ b,
                     float* __restrict__ out, float* __restrict__ xhat, float* __restrict__ rstd_out, int C) {
    extern __shared__ float rbuf[];            // [C]
    __shared__ float red[4];
    __shared__ float bc;
    const size_t row = blockIdx.x;
    const int tid = threadIdx.x;
    float s = 0.f;
    for (int c = tid; c < C; c += 128) {
        float z = Z[row * C + c];                           // Z = sum over heads of the per-head fc partials (+ bias), fixed order
        for (int h2 = 1; h2 < nparts; ++h2) z += Z[(size_t)h2 * part_stride + row * C + c];
        if (zbias) z += zbias[c];
        if (keep) z = keep[row * C + c] ? z * drop_scale : 0.f;
        float v = z + q[row * C + c];
        rbuf[c] = v;
        s += v;
    }
    s = warp_sum(s);
    if ((tid & 31) == 0) red[tid >> 5] = s;
    __syncthreads();
    if (tid == 0) bc = (red[0] + red[1] + red[2] + red[3]) / (float)C;
    __syncthreads();
    const float mean = bc;
    float v2 = 0.f;
    for (int c = tid; c < C; c += 128) { float d = rbuf[c] - mean; v2 = fmaf(d, d, v2); }
    v2 = warp_sum(v2);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = v2;
    __syncthreads();
    if (tid == 0) bc = rsqrtf((red[0] + red[1] + red[2] + red[3]) / (float)C + 1e-5f);
    __syncthreads();
    const float rstd = bc;
    for (int c = tid; c < C; c += 128) {
        float xh = (rbuf[c] - mean) * rstd;
        out[row * C + c] = fmaf(xh, g[c], b[c]);
        if (xhat) xhat[row * C + c] = xh;
    }
    if (rstd_out && tid == 0) rstd_out[row] = rstd;
}

// ---- backward element-wise pieces ---------------------------------------------------
// dR = (g*dY - mean(g*dY) - xhat*mean(g*dY*xhat)) * rstd ; dZ = dR*keep/(1-p)     grid (rows)
__global__ void __launch_bounds__(128)
k_layernorm_bwd_rows(const float* __restrict__ dY, const float* __restrict__ xhat, const float* __restrict__ rstd,
                     const float* __restrict__ g, const uint8_t* __restrict__ keep, float drop_scale,
                     float* __restrict__ dZ, int C) {
    __shared__ float red[2][4];
    __shared__ float bc[2];
    const size_t row = blockIdx.x;
    const int tid = threadIdx.x;
    float s1 = 0.f, s2 = 0.f;
    for (int c = tid; c < C; c += 128) {
        float gd = g[c] * dY[row * C + c];
        s1 += gd;
        s2 = fmaf(gd, xhat[row * C + c], s2);
    }
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    if ((tid & 31) == 0) { red[0][tid >> 5] = s1; red[1][tid >> 5] = s2; }
    __syncthreads();
    if (tid < 2) bc[tid] = (red[tid][0] + red[tid][1] + red[tid][2] + red[tid][3]) / (float)C;
    __syncthreads();
    const float m1 = bc[0], m2 = bc[1], rs = rstd[row];
    for (int c = tid; c < C; c += 128) {
        float gd = g[c] * dY[row * C + c];
        float dr = (gd - m1 - xhat[row * C + c] * m2) * rs;
        if (keep) dr = keep[row * C + c] ? dr * drop_scale : 0.f;
        dZ[row * C + c] = dr;
    }
}

// column sums over rows: out0[c] = sum_rows a[row][c]*b[row][c] (b nullable => a only)   grid (ceil(C/128))
__global__ void __launch_bounds__(128)
k_colsum(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int rows, int C) {
    const int c = blockIdx.x * 128 + threadIdx.x;
    if (c >= C) return;
    float s = 0.f;
    for (int r = 0; r < rows; ++r) {
        float v = a[(size_t)r * C + c];
        s = b ? fmaf(v, b[(size_t)r * C + c], s) : s + v;
    }
    out[c] = s;
}

// dS = P * (dP - sum_p dP*P), dP = dPraw*inv_n*keep/(1-p) ; writes dSm = dS*inv_n/tau (in place over dPraw)
__global__ void __launch_bounds__(256)
k_softmax_bwd_rows(float* __restrict__ dPraw, const float* __restrict__ P, const float* __restrict__ inv_n,
                   const uint8_t* __restrict__ keep, float drop_scale, float inv_tau, int E, int Lq, int nH, int HW) {
    __shared__ float red[8];
    __shared__ float bc;
    const int r = blockIdx.x, e = blockIdx.y, l = r / nH, hh = r - l * nH;
    const int tid = threadIdx.x;
    const size_t row = ((size_t)e * (Lq * nH) + r) * HW;
    const float* in = inv_n + (size_t)e * HW;
    const uint8_t* kp = keep ? keep + (((size_t)hh * E + e) * Lq + l) * HW : nullptr;
    float dot = 0.f;
    for (int p = tid; p < HW; p += 256) {
        float dp = dPraw[row + p] * in[p];
        if (kp) dp = kp[p] ? dp * drop_scale : 0.f;
        dot = fmaf(dp, P[row + p], dot);
    }
    dot = warp_sum(dot);
    if ((tid & 31) == 0) red[tid >> 5] = dot;
    __syncthreads();
    if (tid == 0) { float t = 0.f; for (int k = 0; k < 8; ++k) t += red[k]; bc = t; }
    __syncthreads();
    dot = bc;
    for (int p = tid; p < HW; p += 256) {
        float dp = dPraw[row + p] * in[p];
        if (kp) dp = kp[p] ? dp * drop_scale : 0.f;
        dPraw[row + p] = P[row + p] * (dp - dot) * in[p] * inv_tau;
    }
}

// ---- buffers ------------------------------------------------------------------------
struct TSaved { float* Qp; float* P; float* inv_n; float* aX; float* O; float* xhat; float* rstd; };
static size_t carve_saved(Carver& cv, TSaved& s, int E, int Lq, int nH, int C, int HW) {
    const size_t R = (size_t)Lq * nH;
    s.Qp = cv.take<float>((size_t)E * Lq * nH * C);
    s.P = cv.take<float>((size_t)E * R * HW);
    s.inv_n = cv.take<float>((size_t)E * HW);
    s.aX = cv.take<float>((size_t)E * R * C);
    s.O = cv.take<float>((size_t)E * Lq * nH * C);
    s.xhat = cv.take<float>((size_t)E * Lq * C);
    s.rstd = cv.take<float>((size_t)E * Lq);
    return align_up(cv.off);
}
struct TFwdWs { TSaved local; float* QA; float* sraw; float* n2; float* Pm; float* Z; void* tc; size_t tc_bytes; };
static size_t carve_fwd(Carver& cv, TFwdWs& w, int E, int Lq, int nH, int C, int HW, int algo) {
    const size_t R = (size_t)Lq * nH;
    carve_saved(cv, w.local, E, Lq, nH, C, HW);   // used when the caller keeps nothing for backward
    w.QA = cv.take<float>((size_t)E * R * C);
    w.sraw = cv.take<float>((size_t)E * R * HW);
    w.n2 = cv.take<float>((size_t)E * HW);
    w.Pm = cv.take<float>((size_t)E * R * HW);
    w.Z = cv.take<float>((size_t)nH * E * Lq * C);          // per-head partials of fc
    w.tc_bytes = (algo == CWT_ATTN_TCGEN05) ? kproj_tcgen05_workspace_bytes(E, Lq, nH, C, HW) : 0;
    w.tc = cv.take<char>(w.tc_bytes);
    return align_up(cv.off);
}
struct TBwdWs { float* dZ; float* dO; float* dOA; float* dP; float* dSX; float* dQ; };
static size_t carve_bwd(Carver& cv, TBwdWs& w, int E, int Lq, int nH, int C, int HW) {
    const size_t R = (size_t)Lq * nH;
    w.dZ = cv.take<float>((size_t)E * Lq * C);
    w.dO = cv.take<float>((size_t)E * Lq * nH * C);
    w.dOA = cv.take<float>((size_t)E * R * C);
    w.dP = cv.take<float>((size_t)E * R * HW);
    w.dSX = cv.take<float>((size_t)E * R * C);
    w.dQ = cv.take<float>((size_t)E * Lq * nH * C);
    return align_up(cv.off);
}

}  // namespace cwt

using namespace cwt;

extern "C" size_t cwt_transformer_saved_bytes(int E, int Lq, int nH, int C, int HW) {
    Carver cv(nullptr, 0);
    TSaved s;
    return carve_saved(cv, s, E, Lq, nH, C, HW);
}

extern "C" size_t cwt_transformer_workspace_bytes(int E, int Lq, int nH, int C, int HW, int algo) {
    Carver cv(nullptr, 0);
    TFwdWs w;
    size_t f = carve_fwd(cv, w, E, Lq, nH, C, HW, algo);
    Carver cv2(nullptr, 0);
    TBwdWs b;
    size_t bw = carve_bwd(cv2, b, E, Lq, nH, C, HW);
    return f > bw ? f : bw;
}

extern "C" int cwt_transformer_fwd_f32(const float* q, const float* k, int normalize_k,
                                       const float* w_qkvs, const float* fc_w, const float* fc_b,
                                       const float* ln_g, const float* ln_b,
                                       const uint8_t* keep_attn, const uint8_t* keep_out,
                                       float p_attn, float p_out, float* out, void* saved_or_null,
                                       int E, int Lq, int nH, int C, int HW, int algo,
                                       void* workspace, size_t ws_bytes, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (E == 0) return CWT_OK;
    CWT_REQUIRE(q && k && w_qkvs && fc_w && fc_b && ln_g && ln_b && out, CWT_ERR_INVALID_ARG, "transformer_fwd: null pointer");
    CWT_REQUIRE(E > 0 && Lq >= 1 && nH >= 1 && C >= 1 && HW >= 1, CWT_ERR_INVALID_ARG, "transformer_fwd: bad sizes");
    CWT_REQUIRE(Lq * nH <= 16, CWT_ERR_UNSUPPORTED, "transformer_fwd: Lq*n_head = %d > 16", Lq * nH);
    CWT_REQUIRE(algo == CWT_ATTN_REASSOC || algo == CWT_ATTN_TCGEN05, CWT_ERR_INVALID_ARG, "transformer_fwd: algo %d", algo);
    CWT_REQUIRE(!(keep_attn && !(p_attn < 1.f)) && !(keep_out && !(p_out < 1.f)), CWT_ERR_INVALID_ARG,
                "transformer_fwd: dropout probability must be < 1");
    Carver cv(workspace, ws_bytes);
    TFwdWs w;
    size_t need = carve_fwd(cv, w, E, Lq, nH, C, HW, algo);
    CWT_REQUIRE(workspace && ws_bytes >= need, CWT_ERR_WORKSPACE, "transformer_fwd: workspace %zu < %zu", ws_bytes, need);
    TSaved sv = w.local;
    if (saved_or_null) {
        Carver cs(saved_or_null, (size_t)-1);
        carve_saved(cs, sv, E, Lq, nH, C, HW);
    }
    const int R = Lq * nH, ML = E * Lq, NC = nH * C;
    const float inv_tau = 1.f / sqrtf((float)C);
    int rc;
    // Q = q A^T
    rc = gemm(q, C, 1, w_qkvs, 1, C, sv.Qp, NC, ML, NC, C, 1.f, false, nullptr, 1, 0, 0, 0, st);
    if (rc) return rc;
    if (algo == CWT_ATTN_TCGEN05) {
        rc = kproj_scores_tcgen05(k, w_qkvs, sv.Qp, w.sraw, normalize_k ? w.n2 : nullptr, E, Lq, nH, C, HW,
                                  w.tc, w.tc_bytes, st);
        if (rc) return rc;
    } else {
        // QA[e][l*nH+h][:] = Q_h[e,l,:] A_h        (batch over heads)
        rc = gemm(sv.Qp, NC, 1, w_qkvs, C, 1, w.QA, NC, ML, C, C, 1.f, false, nullptr, nH, C, (long long)C * C, C, st);
        if (rc) return rc;
        rc = launch_rows_times_feat_auto(k, w.QA, w.sraw, normalize_k ? w.n2 : nullptr, E, C, HW, 1, R, st);
        if (rc) return rc;
    }
    k_softmax_rows<<<dim3(R, E), 256, 0, st>>>(w.sraw, normalize_k ? w.n2 : nullptr, keep_attn, inv_tau,
                                               keep_attn ? 1.f / (1.f - p_attn) : 1.f, sv.P, w.Pm, sv.inv_n,
                                               E, Lq, nH, HW);
    CWT_LAUNCHED("softmax_rows");
    rc = launch_feat_times_cols_auto(k, w.Pm, sv.aX, E, 1, C, HW, R, st);
    if (rc) return rc;
    // O[(e,l)][h*C+n] = aX[e][l*nH+h][:] . A_h[n][:]
    rc = gemm(sv.aX, NC, 1, w_qkvs, 1, C, sv.O, NC, ML, C, C, 1.f, false, nullptr, nH, C, (long long)C * C, C, st);
    if (rc) return rc;
    // Z = O Fc^T + b
    //   one GEMM per head (K = C instead of nH*C, nH times the CTAs): Z_h = O_h Fc_h^T; the partials are summed, with the bias,
    //   by the LayerNorm kernel in a fixed order
    rc = gemm(sv.O, NC, 1, fc_w, 1, NC, w.Z, C, ML, C, C, 1.f, false, nullptr, nH, C, C, (long long)ML * C, st);
    if (rc) return rc;
    k_residual_layernorm<<<ML, 128, sizeof(float) * C, st>>>(w.Z, nH, (size_t)ML * C, fc_b, q, keep_out, keep_out ? 1.f / (1.f - p_out) : 1.f,
                                                             ln_g, ln_b, out, sv.xhat, sv.rstd, C);
    CWT_LAUNCHED("residual_layernorm");
    return CWT_OK;
}

extern "C" int cwt_transformer_bwd_f32(const float* d_out, const float* q, const float* k, int normalize_k,
                                       const float* w_qkvs, const float* fc_w, const float* ln_g,
                                       const uint8_t* keep_attn, const uint8_t* keep_out,
                                       float p_attn, float p_out, const void* saved,
                                       float* d_w_qkvs, float* d_fc_w, float* d_fc_b, float* d_ln_g, float* d_ln_b,
                                       int E, int Lq, int nH, int C, int HW,
                                       void* workspace, size_t ws_bytes, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CWT_REQUIRE(d_out && q && k && w_qkvs && fc_w && ln_g && saved && d_w_qkvs && d_fc_w && d_fc_b && d_ln_g && d_ln_b,
                CWT_ERR_INVALID_ARG, "transformer_bwd: null pointer");
    CWT_REQUIRE(E > 0 && Lq >= 1 && nH >= 1 && Lq * nH <= 16, CWT_ERR_INVALID_ARG, "transformer_bwd: bad sizes");
    (void)normalize_k;   // inv_n was saved by the forward (all ones when the forward did not normalise)
    Carver cv(workspace, ws_bytes);
    TBwdWs w;
    size_t need = carve_bwd(cv, w, E, Lq, nH, C, HW);
    CWT_REQUIRE(workspace && ws_bytes >= need, CWT_ERR_WORKSPACE, "transformer_bwd: workspace %zu < %zu", ws_bytes, need);
    TSaved sv;
    Carver cs(const_cast<void*>(saved), (size_t)-1);
    carve_saved(cs, sv, E, Lq, nH, C, HW);
    const int R = Lq * nH, ML = E * Lq, NC = nH * C;
    const float inv_tau = 1.f / sqrtf((float)C);
    int rc;
    // LayerNorm affine grads and dZ
    k_colsum<<<(C + 127) / 128, 128, 0, st>>>(d_out, sv.xhat, d_ln_g, ML, C);
    CWT_LAUNCHED("colsum");
    k_colsum<<<(C + 127) / 128, 128, 0, st>>>(d_out, nullptr, d_ln_b, ML, C);
    CWT_LAUNCHED("colsum");
    k_layernorm_bwd_rows<<<ML, 128, 0, st>>>(d_out, sv.xhat, sv.rstd, ln_g, keep_out,
                                             keep_out ? 1.f / (1.f - p_out) : 1.f, w.dZ, C);
    CWT_LAUNCHED("layernorm_bwd_rows");
    // fc: dFc = dZ^T O ; db = colsum dZ ; dO = dZ Fc
    rc = gemm(w.dZ, 1, C, sv.O, NC, 1, d_fc_w, NC, C, NC, ML, 1.f, false, nullptr, 1, 0, 0, 0, st);
    if (rc) return rc;
    k_colsum<<<(C + 127) / 128, 128, 0, st>>>(w.dZ, nullptr, d_fc_b, ML, C);
    CWT_LAUNCHED("colsum");
    rc = gemm(w.dZ, C, 1, fc_w, NC, 1, w.dO, NC, ML, NC, C, 1.f, false, nullptr, 1, 0, 0, 0, st);
    if (rc) return rc;
    // dPt_h = dO_h K_h^T = (dO_h A_h) X^T
    rc = gemm(w.dO, NC, 1, w_qkvs, C, 1, w.dOA, NC, ML, C, C, 1.f, false, nullptr, nH, C, (long long)C * C, C, st);
    if (rc) return rc;
    rc = launch_rows_times_feat_auto(k, w.dOA, w.dP, nullptr, E, C, HW, 1, R, st);
    if (rc) return rc;
    k_softmax_bwd_rows<<<dim3(R, E), 256, 0, st>>>(w.dP, sv.P, sv.inv_n, keep_attn,
                                                   keep_attn ? 1.f / (1.f - p_attn) : 1.f, inv_tau, E, Lq, nH, HW);
    CWT_LAUNCHED("softmax_bwd_rows");
    // dSX = (dS X)/tau ; dQ_h = dSX_h A_h^T
    rc = launch_feat_times_cols_auto(k, w.dP, w.dSX, E, 1, C, HW, R, st);
    if (rc) return rc;
    rc = gemm(w.dSX, NC, 1, w_qkvs, 1, C, w.dQ, NC, ML, C, C, 1.f, false, nullptr, nH, C, (long long)C * C, C, st);
    if (rc) return rc;
    // dA_h = dQ_h^T q + dO_h^T (Pt_h X) + Q_h^T (dS_h X)/tau     (K = E*Lq rows; sums over episodes)
    rc = gemm(w.dQ, 1, NC, q, C, 1, d_w_qkvs, C, NC, C, ML, 1.f, false, nullptr, 1, 0, 0, 0, st);
    if (rc) return rc;
    rc = gemm(w.dO, 1, NC, sv.aX, NC, 1, d_w_qkvs, C, C, C, ML, 1.f, true, nullptr, nH, C, C, (long long)C * C, st);
    if (rc) return rc;
    rc = gemm(sv.Qp, 1, NC, w.dSX, NC, 1, d_w_qkvs, C, C, C, ML, 1.f, true, nullptr, nH, C, C, (long long)C * C, st);
    if (rc) return rc;
    return CWT_OK;
}
