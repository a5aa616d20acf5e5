// CWT_FIT_RESIDENT — the support-classifier fit with the feature map kept ON CHIP for all n_iter steps.
//
// The streaming fit re-reads the [C, h*w] feature map from HBM twice per SGD step ((2T+1) F bytes per
// episode). Here one episode is spread over a GROUP of CPG = h*w / NP co-resident CTAs (one per SM);
// CTA k stages its strip of NP consecutive low-res pixels x all C channels (C*NP*4 B, 204.8 KB for
// 512 x 100) into shared memory ONCE with bulk-TMA copies (cp.async.bulk + mbarrier complete_tx) and
// then runs every step out of shared memory:
//
//   P1  z[p]   = sum_c Wd[c] F[c][p]          linear conflict-free sweep of the strip, own NP pixels
//   X1  publish z strip, read the (w+1)-pixel halos of strips k-1 / k+1        (neighbour flags)
//   HR  full-resolution stage on the NP + w + 1 cells that touch the strip      (hires_cell, gather form)
//   P3  dW[c]  = sum_p g[p] F[c][p]           thread = channel, second sweep of the strip
//   X2  group all-reduce of dW (fp32 reductions at L2 + one group barrier), SGD update of W0/W1/Wd
//
// HBM traffic per episode drops from (2T+1) F to F (+ labels); the per-step bound becomes the shared
// memory sweep (2 x C*NP*4 B at 128 B/clk/SM) plus two group synchronisations. floor(#SM / CPG) groups
// run concurrently (4 x 36 = 144 of 148 SMs for 60x60x512), each looping over its share of the batch.
//
// Launch: cooperative (all CTAs must be co-resident: the groups spin on each other's flags).
// Every spin loop has a watchdog: on timeout the kernel raises an abort flag, stops waiting and
// writes NaN results rather than hanging the GPU.
#include "common.cuh"
#include "hires.cuh"
#include <cstdlib>

namespace cwt {

constexpr int RES_THREADS = 512;
constexpr unsigned RES_SPIN_LIMIT = 1u << 24;

struct ResidentParams {
    const float* f_s;        // [E][C][HW]
    const uint4* cells;      // [E][HW]
    const float2* cw;        // [E]
    float* w;                // [E][2][C]  in: W0, out: fitted
    float* zbuf;             // [G][2][HW]
    float* acc;              // [G][3][C]
    unsigned* zflag;         // [G][CPG]
    unsigned* bar;           // [G][32]   (one counter per 128 B)
    unsigned* abort_flag;    // [1]
    long long* prof;         // [grid][8] or null
    int E, C, HW, h, w_lo, NP, CPG, G, T;
    float lr;
};

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_release_add(unsigned* p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_add_f32(float* p, float v) {
    asm volatile("red.relaxed.gpu.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
// spin until *p >= target; returns false on watchdog / abort
__device__ __forceinline__ bool spin_until(const unsigned* p, unsigned target, unsigned* abort_flag) {
    unsigned it = 0;
    while (ld_acquire(p) < target) {
        if ((++it & 0x3ffu) == 0u) {
            if (*reinterpret_cast<volatile unsigned*>(abort_flag) != 0u) return false;
            if (it > RES_SPIN_LIMIT) { atomicExch(abort_flag, 1u); return false; }
        }
    }
    return true;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// bulk-TMA copy global -> shared, completion signalled on the mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// half a cell (4 of its 8 rows): same arithmetic as hires_cell, used to spread the NP+w+1 cells of a
// strip over twice as many threads
__device__ __forceinline__ void hires_half_cell(float z00, float z01, float z10, float z11, uint32_t w0, uint32_t w1,
                                                int r0, float c0, float c1, float& o00, float& o01, float& o10,
                                                float& o11) {
    const uint32_t words[2] = {w0, w1};
    const float dl = (z10 - z00) * 0.125f, dr = (z11 - z01) * 0.125f;
    float a00 = 0.f, a01 = 0.f, a10 = 0.f, a11 = 0.f;
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
        const float r = (float)(r0 + rr);
        const uint32_t rb = (words[rr >> 1] >> ((rr & 1) * 16)) & 0xffffu;
        const float left = fmaf(r, dl, z00), right = fmaf(r, dr, z01);
        const float slope = (right - left) * 0.125f;
        float gs = 0.f, gr = 0.f;
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            const uint32_t code = (rb >> (2 * s)) & 3u;
            const float d = fmaf((float)s, slope, left);
            const float p = __fdividef(1.f, 1.f + __expf(-d));
            const float coef = (code == 0u) ? c0 : ((code == 1u) ? c1 : 0.f);
            const float g = coef * (p - (float)(code & 1u));
            gs += g;
            gr = fmaf((float)s, g, gr);
        }
        gr *= 0.125f;
        const float gl = gs - gr;
        const float h1 = r * 0.125f, h0 = 1.f - h1;
        a00 = fmaf(h0, gl, a00); a01 = fmaf(h0, gr, a01);
        a10 = fmaf(h1, gl, a10); a11 = fmaf(h1, gr, a11);
    }
    o00 = a00; o01 = a01; o10 = a10; o11 = a11;
}

struct ResSmem {
    size_t F, W0, W1, Wd, zfull, g, scratch, cellbits, mbar, total;
};
static __host__ __device__ inline ResSmem res_smem_layout(int C, int NP, int w_lo) {
    ResSmem s;
    const int NQ = NP / 4, NG = RES_THREADS / NQ, NCELL = NP + w_lo + 1;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o = (o + bytes + 127) / 128 * 128; return r; };
    s.F = take((size_t)C * NP * 4);
    s.W0 = take((size_t)C * 4);
    s.W1 = take((size_t)C * 4);
    s.Wd = take((size_t)C * 4);
    s.zfull = take((size_t)(NP + 2 * (w_lo + 1)) * 4);
    s.g = take((size_t)NP * 4);
    const size_t sc1 = (size_t)NG * NP * 4, sc2 = (size_t)8 * NCELL * 4;
    s.scratch = take(sc1 > sc2 ? sc1 : sc2);
    s.cellbits = take((size_t)NCELL * 16);
    s.mbar = take(8);
    s.total = o;
    return s;
}

template <bool PROF>
__global__ void __launch_bounds__(RES_THREADS, 1) k_fit_resident(ResidentParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int C = p.C, NP = p.NP, HW = p.HW, wl = p.w_lo, h = p.h;
    const ResSmem L = res_smem_layout(C, NP, wl);
    float* F = reinterpret_cast<float*>(smem_raw + L.F);
    float* W0 = reinterpret_cast<float*>(smem_raw + L.W0);
    float* W1 = reinterpret_cast<float*>(smem_raw + L.W1);
    float* Wd = reinterpret_cast<float*>(smem_raw + L.Wd);
    float* zfull = reinterpret_cast<float*>(smem_raw + L.zfull);
    float* gsm = reinterpret_cast<float*>(smem_raw + L.g);
    float* scratch = reinterpret_cast<float*>(smem_raw + L.scratch);
    uint4* cellbits = reinterpret_cast<uint4*>(smem_raw + L.cellbits);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw + L.mbar);

    const int tid = threadIdx.x;
    const int group = blockIdx.x / p.CPG, k = blockIdx.x - group * p.CPG;
    const int NQ = NP / 4, NG = RES_THREADS / NQ, NCELL = NP + wl + 1, HALO = wl + 1;
    const bool p1_active = tid < NQ * NG;
    const int v = tid % NQ, grp = tid / NQ;
    const int strip0 = k * NP;                       // first own pixel
    const int cell0 = strip0 - HALO;                 // pixel / cell index of zfull[0] / cell 0

    float* zbuf = p.zbuf + (size_t)group * 2 * HW;
    float* acc = p.acc + (size_t)group * 3 * C;
    unsigned* zflag = p.zflag + (size_t)group * p.CPG;
    unsigned* bar = p.bar + (size_t)group * 32;

    if (tid == 0) mbar_init(mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    unsigned gstep = 0, tma_parity = 0;
    bool ok = true;
    long long t_acc[6] = {0, 0, 0, 0, 0, 0};

    for (int e = group; e < p.E; e += p.G) {
        // ---------------- stage the episode: strip of F via bulk-TMA, weights, label cells ----------------
        __syncthreads();                                           // previous episode is done with shared memory
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy reads before async-proxy writes
        if (tid == 0) mbar_expect_tx(mbar, (unsigned)(C * NP * 4));
        const float* fsrc = p.f_s + (size_t)e * C * HW + strip0;
        for (int c = tid; c < C; c += RES_THREADS) bulk_g2s(F + (size_t)c * NP, fsrc + (size_t)c * HW, (unsigned)(NP * 4), mbar);
        for (int c = tid; c < C; c += RES_THREADS) {
            const float a = p.w[((size_t)e * 2) * C + c], b = p.w[((size_t)e * 2 + 1) * C + c];
            W0[c] = a; W1[c] = b; Wd[c] = b - a;
        }
        for (int ci = tid; ci < NCELL; ci += RES_THREADS) {
            const int q = cell0 + ci;
            cellbits[ci] = (q >= 0) ? p.cells[(size_t)e * HW + q] : make_uint4(0xAAAAAAAAu, 0xAAAAAAAAu, 0xAAAAAAAAu, 0xAAAAAAAAu);
        }
        const float2 c01 = p.cw[e];
        {
            unsigned it = 0;
            while (!mbar_try_wait(mbar, tma_parity)) { if (++it > RES_SPIN_LIMIT) { ok = false; break; } }
            tma_parity ^= 1u;
        }
        __syncthreads();

        for (int t = 0; t < p.T; ++t, ++gstep) {
            long long tk0 = 0;
            if (PROF && tid == 0) tk0 = clock64();
            // ---------------- P1: z = Wd . F over the own strip ----------------
            {
                float4 za = make_float4(0.f, 0.f, 0.f, 0.f);
                if (p1_active) {
                    const float4* F4 = reinterpret_cast<const float4*>(F) + tid;
                    const int stride = NQ * NG;
#pragma unroll 4
                    for (int c = grp; c < C; c += NG) {
                        const float4 f = *F4;
                        F4 += stride;
                        const float wd = Wd[c];
                        za.x = fmaf(wd, f.x, za.x); za.y = fmaf(wd, f.y, za.y);
                        za.z = fmaf(wd, f.z, za.z); za.w = fmaf(wd, f.w, za.w);
                    }
                    *reinterpret_cast<float4*>(scratch + grp * NP + 4 * v) = za;
                }
                __syncthreads();
                if (tid < NP) {
                    float z = 0.f;
                    for (int g2 = 0; g2 < NG; ++g2) z += scratch[g2 * NP + tid];
                    zfull[HALO + tid] = z;
                    zbuf[(size_t)(gstep & 1u) * HW + strip0 + tid] = z;
                }
            }
            if (PROF && tid == 0) { long long n = clock64(); t_acc[0] += n - tk0; tk0 = n; }
            // ---------------- X1: publish the strip, fetch both halos ----------------
            __threadfence();
            __syncthreads();
            if (tid == 0) {
                st_release(&zflag[k], gstep + 1u);
                if (k > 0 && ok) ok = spin_until(&zflag[k - 1], gstep + 1u, p.abort_flag);
            } else if (tid == 32) {
                if (k + 1 < p.CPG && ok) ok = spin_until(&zflag[k + 1], gstep + 1u, p.abort_flag);
            }
            __syncthreads();
            if (tid < 2 * HALO) {
                const bool before = tid < HALO;
                const int j = before ? tid : tid - HALO;
                const int q = before ? cell0 + j : strip0 + NP + j;
                float z = 0.f;
                if (q >= 0 && q < HW) z = __ldcg(&zbuf[(size_t)(gstep & 1u) * HW + q]);
                zfull[before ? j : HALO + NP + j] = z;
            }
            __syncthreads();
            if (PROF && tid == 0) { long long n = clock64(); t_acc[1] += n - tk0; tk0 = n; }
            // ---------------- HR: full-resolution stage on the cells touching the strip ----------------
            for (int task = tid; task < 2 * NCELL; task += RES_THREADS) {
                const int ci = task >> 1, half = task & 1;
                const int q = cell0 + ci;
                float o00 = 0.f, o01 = 0.f, o10 = 0.f, o11 = 0.f;
                if (q >= 0) {
                    const int a = q / wl, b = q - a * wl;
                    const int dx = (b + 1 < wl) ? 1 : 0, dy = (a + 1 < h) ? wl : 0;
                    const uint4 bits = cellbits[ci];
                    hires_half_cell(zfull[ci], zfull[ci + dx], zfull[ci + dy], zfull[ci + dy + dx],
                                    half ? bits.z : bits.x, half ? bits.w : bits.y, half * 4, c01.x, c01.y,
                                    o00, o01, o10, o11);
                }
                float* cc = scratch + (size_t)half * 4 * NCELL;
                cc[0 * NCELL + ci] = o00; cc[1 * NCELL + ci] = o01;
                cc[2 * NCELL + ci] = o10; cc[3 * NCELL + ci] = o11;
            }
            __syncthreads();
            if (tid < NP) {
                const int q = strip0 + tid, ci = HALO + tid;
                const int a = q / wl, b = q - a * wl;
                float s = 0.f;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const float* cc = scratch + (size_t)half * 4 * NCELL;
                    s += cc[0 * NCELL + ci];
                    if (b > 0) s += cc[1 * NCELL + ci - 1];
                    if (a > 0) s += cc[2 * NCELL + ci - wl];
                    if (a > 0 && b > 0) s += cc[3 * NCELL + ci - wl - 1];
                }
                gsm[tid] = s;
            }
            __syncthreads();
            if (PROF && tid == 0) { long long n = clock64(); t_acc[2] += n - tk0; tk0 = n; }
            // ---------------- P3: dW = g . F^T (thread = channel), L2 reductions into the group accumulator ----------------
            {
                float* acc_cur = acc + (size_t)(gstep % 3u) * C;
                float* acc_nxt = acc + (size_t)((gstep + 1u) % 3u) * C;
                for (int c = tid; c < C; c += RES_THREADS) {
                    const float4* row = reinterpret_cast<const float4*>(F + (size_t)c * NP);
                    const float4* g4 = reinterpret_cast<const float4*>(gsm);
                    float d0 = 0.f, d1 = 0.f;
#pragma unroll 5
                    for (int j = 0; j < NQ; ++j) {
                        const float4 f = row[j];
                        const float4 gg = g4[j];
                        d0 = fmaf(gg.x, f.x, d0); d1 = fmaf(gg.y, f.y, d1);
                        d0 = fmaf(gg.z, f.z, d0); d1 = fmaf(gg.w, f.w, d1);
                    }
                    red_add_f32(&acc_cur[c], d0 + d1);
                    if (c % p.CPG == k) acc_nxt[c] = 0.f;          // recycled two barriers from now
                }
            }
            if (PROF && tid == 0) { long long n = clock64(); t_acc[3] += n - tk0; tk0 = n; }
            // ---------------- X2: group barrier, then everyone applies the same SGD update ----------------
            __threadfence();
            __syncthreads();
            if (tid == 0) {
                red_release_add(bar, 1u);
                if (ok) ok = spin_until(bar, (gstep + 1u) * (unsigned)p.CPG, p.abort_flag);
            }
            __syncthreads();
            for (int c = tid; c < C; c += RES_THREADS) {
                const float dw = __ldcg(&acc[(size_t)(gstep % 3u) * C + c]);
                const float n0 = fmaf(p.lr, dw, W0[c]), n1 = fmaf(-p.lr, dw, W1[c]);
                W0[c] = n0; W1[c] = n1; Wd[c] = n1 - n0;
            }
            __syncthreads();
            if (PROF && tid == 0) { long long n = clock64(); t_acc[4] += n - tk0; tk0 = n; }
        }
        // ---------------- result ----------------
        if (k == 0) {
            const bool bad = *reinterpret_cast<volatile unsigned*>(p.abort_flag) != 0u;
            for (int c = tid; c < C; c += RES_THREADS) {
                p.w[((size_t)e * 2) * C + c] = bad ? __int_as_float(0x7fc00000) : W0[c];
                p.w[((size_t)e * 2 + 1) * C + c] = bad ? __int_as_float(0x7fc00000) : W1[c];
            }
        }
    }
    if (PROF && tid == 0 && p.prof) {
        for (int i = 0; i < 5; ++i) p.prof[(size_t)blockIdx.x * 8 + i] = t_acc[i];
    }
    if (tid == 0 && !ok) atomicExch(p.abort_flag, 1u);
}

// ---- host side ---------------------------------------------------------------------------------
struct ResidentPlan { int NP, CPG, G; size_t smem; bool ok; };

static ResidentPlan plan_resident(int E, int C, int h, int w, int n_sm, size_t smem_cap) {
    ResidentPlan best{0, 0, 0, 0, false};
    const int HW = h * w;
    int forced = 0;
    if (const char* s = getenv("CWT_RESIDENT_NP")) forced = atoi(s);
    double best_score = -1.0;
    for (int NP = 4; NP <= HW; NP += 4) {
        if (HW % NP) continue;
        if (forced && NP != forced) continue;
        if (NP < w + 1 && HW != NP) continue;          // halos must come from the adjacent strips only
        if (NP / 4 > RES_THREADS) continue;
        const size_t sm = res_smem_layout(C, NP, w).total;
        if (sm > smem_cap) continue;
        const int CPG = HW / NP;
        if (CPG > n_sm) continue;
        int G = n_sm / CPG;
        if (G > E) G = E;
        const double score = (double)G / NP + 1e-9 * NP;
        if (score > best_score) { best_score = score; best = ResidentPlan{NP, CPG, G, sm, true}; }
    }
    return best;
}

size_t fit_resident_workspace_bytes(int C, int h, int w) {
    const int HW = h * w;
    const int maxG = 148;
    return align_up(sizeof(float) * 2 * HW * maxG) + align_up(sizeof(float) * 3 * C * maxG) +
           align_up(sizeof(unsigned) * (size_t)maxG * 160) + align_up(sizeof(unsigned) * 32 * maxG) + 256 +
           align_up(sizeof(long long) * 8 * 160);
}

// returns CWT_ERR_UNSUPPORTED (without error text side effects mattering) when the shape does not fit
int fit_resident(const float* f_s, const uint4* cells, const float2* cw, float* w_io, int E, int C, int h, int w,
                 int n_iter, float lr, void* ws, size_t ws_bytes, long long* prof_out, cudaStream_t st) {
    int dev = 0, n_sm = 0, smem_cap = 0, coop = 0;
    CWT_CUDA(cudaGetDevice(&dev));
    CWT_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    CWT_CUDA(cudaDeviceGetAttribute(&smem_cap, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    CWT_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    CWT_REQUIRE(coop, CWT_ERR_UNSUPPORTED, "fit_resident: device lacks cooperative launch");
    CWT_REQUIRE(C <= RES_THREADS * 4, CWT_ERR_UNSUPPORTED, "fit_resident: C=%d too large", C);
    if (n_sm > 148) n_sm = 148;
    const ResidentPlan pl = plan_resident(E, C, h, w, n_sm, (size_t)smem_cap);
    CWT_REQUIRE(pl.ok, CWT_ERR_UNSUPPORTED, "fit_resident: no strip size fits C=%d, %dx%d in %d B of shared memory", C, h, w, smem_cap);
    const int HW = h * w;
    Carver cv(ws, ws_bytes);
    ResidentParams p{};
    p.zbuf = cv.take<float>((size_t)2 * HW * pl.G);
    p.acc = cv.take<float>((size_t)3 * C * pl.G);
    p.zflag = cv.take<unsigned>((size_t)pl.G * pl.CPG);
    p.bar = cv.take<unsigned>((size_t)32 * pl.G);
    p.abort_flag = cv.take<unsigned>(64);
    p.prof = prof_out ? cv.take<long long>((size_t)8 * pl.G * pl.CPG) : nullptr;
    CWT_REQUIRE(ws && cv.ok(), CWT_ERR_WORKSPACE, "fit_resident: workspace too small");
    // the sync area (accumulators, flags, counters) must start at zero
    const size_t sync_bytes = (size_t)(reinterpret_cast<char*>(p.abort_flag + 64) - reinterpret_cast<char*>(p.acc));
    CWT_CUDA(cudaMemsetAsync(p.acc, 0, sync_bytes, st));
    p.f_s = f_s; p.cells = cells; p.cw = cw; p.w = w_io;
    p.E = E; p.C = C; p.HW = HW; p.h = h; p.w_lo = w; p.NP = pl.NP; p.CPG = pl.CPG; p.G = pl.G; p.T = n_iter; p.lr = lr;
    void* args[] = {&p};
    dim3 grid(pl.G * pl.CPG), block(RES_THREADS);
    if (prof_out) {
        CWT_CUDA(cudaFuncSetAttribute(k_fit_resident<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
        CWT_CUDA(cudaLaunchCooperativeKernel((void*)k_fit_resident<true>, grid, block, args, pl.smem, st));
        count_launch();
        CWT_CUDA(cudaMemcpyAsync(prof_out, p.prof, sizeof(long long) * 8 * pl.G * pl.CPG, cudaMemcpyDeviceToDevice, st));
    } else {
        CWT_CUDA(cudaFuncSetAttribute(k_fit_resident<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
        CWT_CUDA(cudaLaunchCooperativeKernel((void*)k_fit_resident<false>, grid, block, args, pl.smem, st));
        count_launch();
    }
    return CWT_OK;
}

}  // namespace cwt
