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
#include "iou_stream.cuh"
#include <cstdlib>
#include <cstring>

namespace cwt {

int pack_label_cells(const void* labels, int label_kind, int n_img, int h, int w, int H, int W, int ignore_index,
                     uint4* cells, int32_t* counts, cudaStream_t st);

constexpr int IOU_THREADS = 256;
constexpr int IOU_MAXW = 128;
constexpr int IOU_MAXV = 8;           // weight sets (variants) per episode
constexpr int IOU_R = 4;              // low-res cell rows per CTA band

// ------------------------------------------------------------------------------------------------
// Column-threaded upsample / argmax / confusion counting of one band of low-res rows.
//   lsm[v][ch][row][w]  low-res logits of the band (rows a0 .. a0+nlo-1) in shared memory
//   thread <-> hi-res column X: the horizontal lerp t(row) = fma(l[b0], w0, l[b1]*w1) is computed once per
//   low-res row and reused by the 8 hi-res rows below it; vertically u = fma(t0, h0, t1*h1) — exactly ATen's
//   CPU rounding order (verified bit-exact against F.interpolate), 2 FP ops per channel per pixel.
//   cnt[v][pred][code] (code in {0,1}) per thread -> I = n[c][c], T = n[0][c]+n[1][c], U = row_c + col_c - I.
// ------------------------------------------------------------------------------------------------
template <bool I64, int V>
__device__ __forceinline__ void iou_band(const float* __restrict__ lsm, int lrows, const void* __restrict__ lab, size_t lbase,
                                         int a0, int nlo, int h, int w, int H, int W, int ignore_index, unsigned ce_mask,
                                         int (&cnt)[V][4], float (&loss)[V], int& nvalid) {
    const int tid = threadIdx.x;
    for (int X = tid; X < W; X += IOU_THREADS) {
        const int b0 = X >> 3, b1 = min(b0 + 1, w - 1);
        const float w1 = (X & 7) * 0.125f, w0 = 1.f - w1;
        float t0[V][2], t1[V][2];
#pragma unroll
        for (int v = 0; v < V; ++v)
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                const float* row = lsm + ((v * 2 + ch) * lrows) * w;
                t1[v][ch] = __fmaf_rn(row[b0], w0, __fmul_rn(row[b1], w1));
            }
        for (int ar = 0; ar < nlo; ++ar) {                 // cell row a0+ar: hi-res rows 8(a0+ar) .. +7
            const int a = a0 + ar;
            const int an = min(ar + 1, lrows - 1);         // next low-res row inside the band buffer
#pragma unroll
            for (int v = 0; v < V; ++v)
#pragma unroll
                for (int ch = 0; ch < 2; ++ch) {
                    t0[v][ch] = t1[v][ch];
                    const float* row = lsm + ((v * 2 + ch) * lrows + an) * w;
                    t1[v][ch] = (a + 1 < h) ? __fmaf_rn(row[b0], w0, __fmul_rn(row[b1], w1)) : t0[v][ch];
                }
            const int nr = min(8, H - 8 * a);
            // the 8 label bytes of this column's cell row are requested together (one dependent-load latency, not eight)
            int codes[8];
#pragma unroll
            for (int r = 0; r < 8; ++r)
                codes[r] = (r < nr) ? load_label_code<I64>(lab, lbase + (size_t)(8 * a + r) * W + X, ignore_index) : 2;
            unsigned packed[V];                            // four 8-bit counters [pred*2+code] per variant (<= 8 per cell row)
#pragma unroll
            for (int v = 0; v < V; ++v) packed[v] = 0u;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const float h1 = r * 0.125f, h0 = 1.f - h1;
                const int code = codes[r];
                const bool valid = code < 2;
                nvalid += valid;
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    const float u0 = __fmaf_rn(t0[v][0], h0, __fmul_rn(t1[v][0], h1));
                    const float u1 = __fmaf_rn(t0[v][1], h0, __fmul_rn(t1[v][1], h1));
                    const unsigned sh = (u1 > u0 ? 16u : 0u) + (unsigned)code * 8u;      // torch.argmax: first index wins ties
                    packed[v] += valid ? (1u << sh) : 0u;
                    if ((ce_mask >> v) & 1u) {
                        const float t = (code == 1) ? (u0 - u1) : (u1 - u0);             // -log softmax(u)[y] = softplus(t)
                        const float sp = fmaxf(t, 0.f) + __logf(1.f + __expf(-fabsf(t)));
                        loss[v] += valid ? sp : 0.f;
                    }
                }
            }
#pragma unroll
            for (int v = 0; v < V; ++v) {
                cnt[v][0] += packed[v] & 0xffu; cnt[v][1] += (packed[v] >> 8) & 0xffu;
                cnt[v][2] += (packed[v] >> 16) & 0xffu; cnt[v][3] += packed[v] >> 24;
            }
        }
    }
}

// block-level reduction of the per-thread counters into counts[n][class][I,U,T] / ce[n][sum,count]
template <int V>
__device__ __forceinline__ void iou_flush(int (&cnt)[V][4], float (&loss)[V], int nvalid, size_t map0,
                                          unsigned long long* __restrict__ counts, double* __restrict__ ce, int* sred,
                                          float* fred) {
    const int tid = threadIdx.x;
    for (int i = tid; i < V * 4 + 1; i += IOU_THREADS) sred[i] = 0;
    for (int i = tid; i < V; i += IOU_THREADS) fred[i] = 0.f;
    __syncthreads();
#pragma unroll
    for (int v = 0; v < V; ++v) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int s = __reduce_add_sync(0xffffffffu, cnt[v][k]);
            if ((tid & 31) == 0 && s) atomicAdd(&sred[v * 4 + k], s);
        }
        const float l = warp_sum(loss[v]);
        if ((tid & 31) == 0 && l != 0.f) atomicAdd(&fred[v], l);      // (CE is a logged metric: order-insensitive to ~1e-7)
    }
    const int nv = __reduce_add_sync(0xffffffffu, nvalid);
    if ((tid & 31) == 0 && nv) atomicAdd(&sred[V * 4], nv);
    __syncthreads();
    if (tid < V) {
        const int v = tid;
        const int n00 = sred[v * 4 + 0], n01 = sred[v * 4 + 1], n10 = sred[v * 4 + 2], n11 = sred[v * 4 + 3];   // [pred][code]
        unsigned long long* c = counts + (map0 + v) * 6;
        const int I0 = n00, I1 = n11, T0 = n00 + n10, T1 = n01 + n11;
        const int U0 = n00 + n01 + n10, U1 = n11 + n10 + n01;
        if (I0) atomicAdd(&c[0], (unsigned long long)I0);
        if (U0) atomicAdd(&c[1], (unsigned long long)U0);
        if (T0) atomicAdd(&c[2], (unsigned long long)T0);
        if (I1) atomicAdd(&c[3], (unsigned long long)I1);
        if (U1) atomicAdd(&c[4], (unsigned long long)U1);
        if (T1) atomicAdd(&c[5], (unsigned long long)T1);
        if (ce) {
            atomicAdd(&ce[(map0 + v) * 2], (double)fred[v]);
            atomicAdd(&ce[(map0 + v) * 2 + 1], (double)sred[V * 4]);
        }
    }
}

// ready-made low-res logits: grid (bands, n_maps); one 2-channel map per blockIdx.y
template <bool I64>
__global__ void __launch_bounds__(IOU_THREADS, 4)
k_upsample_argmax_iou(const float* __restrict__ l60, const void* __restrict__ lab, int maps_per_label,
                      unsigned long long* __restrict__ counts, double* __restrict__ ce,
                      int h, int w, int H, int W, int ignore_index) {
    __shared__ float lsm[2 * (IOU_R + 1) * IOU_MAXW];
    __shared__ int sred[8];
    __shared__ float fred[1];
    const int n = blockIdx.y, a0 = blockIdx.x * IOU_R;
    const int nlo = min(IOU_R, h - a0), lrows = min(IOU_R + 1, h - a0);
    const float* lm = l60 + (size_t)n * 2 * h * w;
    for (int i = threadIdx.x; i < 2 * lrows * w; i += IOU_THREADS) {
        const int ch = i / (lrows * w), rem = i - ch * lrows * w;
        lsm[i] = lm[(size_t)ch * h * w + a0 * w + rem];
    }
    __syncthreads();
    int cnt[1][4] = {{0, 0, 0, 0}};
    float loss[1] = {0.f};
    int nvalid = 0;
    iou_band<I64, 1>(lsm, lrows, lab, (size_t)(n / maps_per_label) * H * W, a0, nlo, h, w, H, W, ignore_index, ce ? 1u : 0u,
                     cnt, loss, nvalid);
    iou_flush<1>(cnt, loss, nvalid, (size_t)n, counts, ce, sred, fred);
}

static int launch_upsample_iou(const float* l60, const void* lab, int label_kind, int maps_per_label,
                               long long* counts, double* ce, int n, int h, int w, int H, int W,
                               int ignore_index, cudaStream_t st) {
    CWT_REQUIRE(H == 8 * (h - 1) + 1 && W == 8 * (w - 1) + 1, CWT_ERR_UNSUPPORTED,
                "iou: label size %dx%d is not 8*(%dx%d - 1) + 1", H, W, h, w);
    CWT_REQUIRE(w <= IOU_MAXW, CWT_ERR_UNSUPPORTED, "iou: w=%d > %d", w, IOU_MAXW);
    CWT_CUDA(cudaMemsetAsync(counts, 0, sizeof(long long) * 6 * (size_t)n, st));
    if (ce) CWT_CUDA(cudaMemsetAsync(ce, 0, sizeof(double) * 2 * (size_t)n, st));
    dim3 grid((h + IOU_R - 1) / IOU_R, n);
    auto* cnt = reinterpret_cast<unsigned long long*>(counts);
    if (label_kind == CWT_LABEL_I64)
        k_upsample_argmax_iou<true><<<grid, IOU_THREADS, 0, st>>>(l60, lab, maps_per_label, cnt, ce, h, w, H, W, ignore_index);
    else
        k_upsample_argmax_iou<false><<<grid, IOU_THREADS, 0, st>>>(l60, lab, maps_per_label, cnt, ce, h, w, H, W, ignore_index);
    CWT_LAUNCHED("upsample_argmax_iou");
    return CWT_OK;
}

// ------------------------------------------------------------------------------------------------
// (c) fused: query logits of V weight sets -> up to H x W -> argmax -> I/U/T (+CE), ONE pass over f_q.
// CTA = (band of IOU_R cell rows, episode). Phase 1 streams the (IOU_R+1) low-res rows x C channels of f_q
// that the band's cells interpolate from (the shared halo row is re-read from L2 by the next band) and keeps
// the 2V logit rows (+ |f_p|^2 for F.normalize) in shared memory; phase 2 is iou_band.
// ------------------------------------------------------------------------------------------------
template <bool I64, int V>
__global__ void __launch_bounds__(IOU_THREADS, 4)
k_logits_iou_fused(const float* __restrict__ wts, const float* __restrict__ f_q, const void* __restrict__ lab,
                   int normalize_mask, unsigned ce_mask, unsigned long long* __restrict__ counts, float* __restrict__ logits_out,
                   double* __restrict__ ce, int C, int h, int w, int H, int W, int ignore_index,
                   int nbands, int n_items, unsigned* __restrict__ work_counter) {
    extern __shared__ __align__(16) float fsm[];
    constexpr int R2 = 2 * V;                          // logit rows
    __shared__ int s_item;
    const int tid = threadIdx.x;
  // persistent CTAs pull (band, episode) items from a queue: after the first item the CTAs of an SM are in
  // different phases, so one CTA's up-sample/count phase overlaps the others' streaming phase
  for (;;) {
    __syncthreads();
    if (tid == 0) s_item = (int)atomicAdd(work_counter, 1u);
    __syncthreads();
    const int item = s_item;
    if (item >= n_items) break;
    const int e = item / nbands, a0 = (item - e * nbands) * IOU_R;
    const int nlo = min(IOU_R, h - a0), lrows = min(IOU_R + 1, h - a0);
    const int HW = h * w, npx = lrows * w, nq = (npx + 3) >> 2;          // band pixels, float4 quads (HW % 4 == 0)
    const int ng = IOU_THREADS / nq;                   // channel groups
    float* Ms = fsm;                                   // [C][R2]
    float* red = Ms + (size_t)C * R2;                  // [ng][R2+1][nq*4]
    float* lsm = red + (size_t)IOU_THREADS * (R2 + 1) * 4; // [R2][lrows][w]   (ng * nq <= IOU_THREADS)
    __shared__ int sred[IOU_MAXV * 4 + 1];
    __shared__ float fred[IOU_MAXV];
    const float* Mg = wts + (size_t)e * R2 * C;
    for (int i = tid; i < C * R2; i += IOU_THREADS) { const int c = i / R2, r = i - c * R2; Ms[i] = Mg[(size_t)r * C + c]; }
    __syncthreads();
    // ---- phase 1: logits of the band
    const int q = tid % nq, grp = tid / nq;
    if (grp < ng) {
        float acc[R2][4], n2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int r = 0; r < R2; ++r) { acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f; }
        const float* fp = f_q + ((size_t)e * C) * HW + (size_t)a0 * w + 4 * q;
        const bool inb = 4 * q < npx;
        if (inb) {
            // 8 independent 16-byte loads in flight per thread before the first use (memory-level parallelism)
            int c = grp;
            for (; c + 7 * ng < C; c += 8 * ng) {
                float4 t[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) t[u] = ldg_stream4(fp + (size_t)(c + u * ng) * HW);
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const float v[4] = {t[u].x, t[u].y, t[u].z, t[u].w};
                    const float* m = Ms + (c + u * ng) * R2;
#pragma unroll
                    for (int r = 0; r < R2; ++r)
#pragma unroll
                        for (int k = 0; k < 4; ++k) acc[r][k] = fmaf(m[r], v[k], acc[r][k]);
#pragma unroll
                    for (int k = 0; k < 4; ++k) n2[k] = fmaf(v[k], v[k], n2[k]);
                }
            }
            for (; c < C; c += ng) {
                const float4 t = ldg_stream4(fp + (size_t)c * HW);
                const float v[4] = {t.x, t.y, t.z, t.w};
                const float* m = Ms + c * R2;
#pragma unroll
                for (int r = 0; r < R2; ++r)
#pragma unroll
                    for (int k = 0; k < 4; ++k) acc[r][k] = fmaf(m[r], v[k], acc[r][k]);
#pragma unroll
                for (int k = 0; k < 4; ++k) n2[k] = fmaf(v[k], v[k], n2[k]);
            }
        }
        float* rg = red + (size_t)grp * (R2 + 1) * nq * 4;
#pragma unroll
        for (int r = 0; r < R2; ++r)
#pragma unroll
            for (int k = 0; k < 4; ++k) rg[r * nq * 4 + 4 * q + k] = acc[r][k];
#pragma unroll
        for (int k = 0; k < 4; ++k) rg[R2 * nq * 4 + 4 * q + k] = n2[k];
    }
    __syncthreads();
    for (int i = tid; i < R2 * npx; i += IOU_THREADS) {
        const int r = i / npx, pp = i - r * npx;
        float s = 0.f, nn = 0.f;
        for (int g2 = 0; g2 < ng; ++g2) {
            s += red[((size_t)g2 * (R2 + 1) + r) * nq * 4 + pp];
            nn += red[((size_t)g2 * (R2 + 1) + R2) * nq * 4 + pp];
        }
        const int v = r >> 1;
        if ((normalize_mask >> v) & 1) s = s / fmaxf(sqrtf(nn), 1e-12f);          // F.normalize(f_q, dim=1)
        lsm[r * npx + pp] = s;
        if (logits_out && pp < nlo * w) logits_out[((size_t)e * R2 + r) * HW + (size_t)a0 * w + pp] = s;   // own rows only
    }
    __syncthreads();
    // ---- phase 2: up-sample, argmax, count
    int cnt[V][4];
    float loss[V];
#pragma unroll
    for (int v = 0; v < V; ++v) { cnt[v][0] = cnt[v][1] = cnt[v][2] = cnt[v][3] = 0; loss[v] = 0.f; }
    int nvalid = 0;
    iou_band<I64, V>(lsm, lrows, lab, (size_t)e * H * W, a0, nlo, h, w, H, W, ignore_index, ce ? ce_mask : 0u, cnt, loss, nvalid);
    iou_flush<V>(cnt, loss, nvalid, (size_t)e * V, counts, ce, sred, fred);
  }
}

template <bool I64, int V>
static int launch_fused_t(const float* wts, const float* f_q, const void* lab, int normalize_mask, unsigned ce_mask,
                          unsigned long long* counts, float* logits_out, double* ce, int E, int C, int h, int w, int H, int W,
                          int ignore_index, unsigned* work_counter, cudaStream_t st) {
    const int npx = (IOU_R + 1) * w;
    const size_t sm = sizeof(float) * ((size_t)C * 2 * V + (size_t)IOU_THREADS * (2 * V + 1) * 4 + (size_t)2 * V * npx);
    CWT_REQUIRE(sm <= 200 * 1024, CWT_ERR_UNSUPPORTED, "logits_iou: C=%d too large for the fused kernel", C);
    auto kern = k_logits_iou_fused<I64, V>;
    if (sm > 48 * 1024) CWT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    const int nbands = (h + IOU_R - 1) / IOU_R, n_items = nbands * E;
    int dev = 0, n_sm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    const int grid = n_items < n_sm * 4 ? n_items : n_sm * 4;
    CWT_CUDA(cudaMemsetAsync(work_counter, 0, sizeof(unsigned), st));
    kern<<<grid, IOU_THREADS, sm, st>>>(wts, f_q, lab, normalize_mask, ce_mask, counts, logits_out, ce, C, h, w, H, W, ignore_index,
                                        nbands, n_items, work_counter);
    CWT_LAUNCHED("logits_iou_fused");
    return CWT_OK;
}

__global__ void __launch_bounds__(256) k_zero_outputs(unsigned long long* a, size_t na, unsigned long long* b, size_t nb) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");      // the dependent kernel waits before its first atomic
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < na) a[i] = 0ull;
    else if (b && i - na < nb) b[i - na] = 0ull;         // 0.0 as a double
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

struct LogitsIouWs { float* l60; float* n2; unsigned* work; };
static size_t carve_logits_iou(Carver& cv, LogitsIouWs& ws, int E, int V, int HW) {
    ws.l60 = cv.take<float>((size_t)E * V * 2 * HW);
    ws.n2 = cv.take<float>((size_t)E * HW);
    ws.work = cv.take<unsigned>(64);
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
    // streaming single pass (north-star (c)): persistent warp-specialised CTAs fed by a bulk-TMA ring (iou_stream.cuh).
    // CWT_LOGITS_IOU=band selects the older band-per-CTA fused kernel (kept for V > 2 and odd shapes).
    const char* algo_env = getenv("CWT_LOGITS_IOU");
    const bool want_band = algo_env && strcmp(algo_env, "band") == 0;
    if (!want_band && logits_iou_stream_ok(f_q, V, C, h, w, H, W)) {
        {   // counters and CE sums start at zero: one launch for both output buffers
            const size_t n64 = (size_t)6 * E * V + (ce_or_null ? (size_t)2 * E * V : 0);
            k_zero_outputs<<<(unsigned)((n64 + 255) / 256), 256, 0, st>>>(reinterpret_cast<unsigned long long*>(iu_counts), (size_t)6 * E * V,
                                                                      reinterpret_cast<unsigned long long*>(ce_or_null), (size_t)2 * E * V);
            CWT_LAUNCHED("zero_outputs");
        }
        auto* cnt = reinterpret_cast<unsigned long long*>(iu_counts);
        const unsigned ce_mask = (1u << V) - 1u;
        const bool i64 = label_kind == CWT_LABEL_I64;
        int rc;
        if (V == 1) rc = i64 ? launch_logits_iou_stream<true, 1>(wts, f_q, q_label, normalize_mask, ce_mask, cnt, logits60_or_null, ce_or_null, E, C, h, w, H, W, ignore_index, st)
                             : launch_logits_iou_stream<false, 1>(wts, f_q, q_label, normalize_mask, ce_mask, cnt, logits60_or_null, ce_or_null, E, C, h, w, H, W, ignore_index, st);
        else        rc = i64 ? launch_logits_iou_stream<true, 2>(wts, f_q, q_label, normalize_mask, ce_mask, cnt, logits60_or_null, ce_or_null, E, C, h, w, H, W, ignore_index, st)
                             : launch_logits_iou_stream<false, 2>(wts, f_q, q_label, normalize_mask, ce_mask, cnt, logits60_or_null, ce_or_null, E, C, h, w, H, W, ignore_index, st);
        if (rc != CWT_ERR_UNSUPPORTED) return rc;
    }
    // fused band kernel whenever the band loads can be vectorised; otherwise logits pass + histogram pass
    if (HWl % 4 == 0 && V <= 4 && w <= IOU_MAXW && H == 8 * (h - 1) + 1 && W == 8 * (w - 1) + 1 &&
        (size_t)C * 2 * V * 4 <= 96 * 1024) {
        CWT_CUDA(cudaMemsetAsync(iu_counts, 0, sizeof(long long) * 6 * (size_t)E * V, st));
        if (ce_or_null) CWT_CUDA(cudaMemsetAsync(ce_or_null, 0, sizeof(double) * 2 * (size_t)E * V, st));
        auto* cnt = reinterpret_cast<unsigned long long*>(iu_counts);
        const unsigned ce_mask = (1u << V) - 1u;
        const bool i64 = label_kind == CWT_LABEL_I64;
#define CWT_FUSED(VV)                                                                                                       \
        return i64 ? launch_fused_t<true, VV>(wts, f_q, q_label, normalize_mask, ce_mask, cnt, logits60_or_null, ce_or_null, \
                                              E, C, h, w, H, W, ignore_index, ws.work, st)                                 \
                   : launch_fused_t<false, VV>(wts, f_q, q_label, normalize_mask, ce_mask, cnt, logits60_or_null, ce_or_null, \
                                               E, C, h, w, H, W, ignore_index, ws.work, st)
        switch (V) {
            case 1: CWT_FUSED(1);
            case 2: CWT_FUSED(2);
            case 3: CWT_FUSED(3);
            default: CWT_FUSED(4);
        }
#undef CWT_FUSED
    }
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
