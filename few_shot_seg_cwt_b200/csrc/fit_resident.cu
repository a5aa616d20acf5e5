// CWT_FIT_RESIDENT — the support-classifier fit with the feature map kept ON CHIP for all n_iter steps.
//
// The streaming fit re-reads the [C, h*w] feature map from HBM twice per SGD step ((2T+1) F bytes per
// episode). Here one episode is spread over a GROUP of CPG = h*w / NP co-resident CTAs (one per SM);
// CTA k stages its strip of NP consecutive low-res pixels x all C channels (C*NP*4 B, 204.8 KB for
// 512 x 100) into shared memory ONCE with bulk-TMA copies (cp.async.bulk + mbarrier complete_tx) and
// then runs every step out of shared memory:
//
//   P1  z[p]   = sum_c Wd[c] F[c][p]          linear conflict-free sweep of the strip, own NP pixels
//   X1  halo exchange of z with strips k-1 / k+1 (w+1 pixels each side)
//   HR  full-resolution stage on the NP + w + 1 cells that touch the strip (one task = one cell row)
//   P3  dW[c]  = sum_p g[p] F[c][p]           second sweep of the strip, 4 threads per channel
//   AR  group all-reduce of dW, SGD update of W0 / W1 / Wd
//
// All cross-CTA traffic uses SEQUENCE-TAGGED 8-byte words {fp32 value, step number}: 64-bit stores are
// single-copy atomic, so a reader simply polls the data word until its tag matches — no fences, no
// atomics, no barriers, and a stale value can never be consumed. The all-reduce is two such hops,
// pipelined in RES_KCH channel chunks underneath P3 / P1: every CTA owns a slice of the channels; its
// REDUCER WARP (warp 16, never computes) collects the CPG partials of the slice, adds them in a fixed
// order (deterministic), and republishes the sums; the APPLIER WARP (warp 17) picks the sums up, applies
// the SGD update to W0 / W1 / Wd in shared memory and releases the compute warps chunk by chunk through a
// shared-memory mbarrier. The compute warps never wait on global memory for the all-reduce.
//
// HBM traffic per episode drops from (2T+1) F to F (+ labels); the per-step bound becomes the shared
// memory sweep (2 x C*NP*4 B at 128 B/clk/SM). floor(#SM / CPG) groups run concurrently
// (4 x 36 = 144 of 148 SMs for 60x60x512), each looping over its share of the batch.
//
// The kernel is compiled twice: specialised for the PSPNet head geometry (C=512, 60x60, NP=100: all
// loop bounds, divisions and predicates become compile-time) and generic (run-time shapes).
//
// Launch: cooperative (all CTAs must be co-resident: they poll each other's words). Every poll loop
// has a watchdog: on timeout the kernel raises an abort flag, stops waiting and writes NaN results
// rather than hanging the GPU.
#include "common.cuh"
#include "hires.cuh"
#include <cstdlib>

namespace cwt {

constexpr int RES_CTHREADS = 512;               // 16 compute warps
constexpr int RES_THREADS = RES_CTHREADS + 64;  // + the reducer warp (16) and the applier warp (17)
constexpr unsigned RES_SPIN_LIMIT = 1u << 24;
constexpr int RES_KCH = 4;        // channel chunks of the pipelined all-reduce
constexpr int RES_MAXQ = 8;       // float4 pixel-quads per P3 thread (NP <= 128)
constexpr int RES_RW = 5;         // tagged words a reducer lane keeps in flight per chunk
constexpr int RES_AW = 4;         // tagged words an applier lane keeps in flight
constexpr int RES_MAXTASK = 4;    // HR row-tasks per compute thread (8*(NP+w+1) <= 4*512)

struct ResidentParams {
    const float* f_s;          // [E][C][HW]
    const uint4* cells;        // [E][HW]
    const float2* cw;          // [E]
    float* w;                  // [E][2][C]  in: W0, out: fitted
    unsigned long long* zll;   // [G][2][HW]                            {z, step} words of the halo exchange
    unsigned long long* inbox; // [G][2][KCH][CPG owner][CPG src][SLS]  {partial dW, step} words (hop 1)
    unsigned long long* sums;  // [G][2][C]                             {all-reduced dW, step} words (hop 2)
    unsigned* abort_flag;      // [1]
    long long* prof;           // [grid][8] or null
    int E, C, HW, h, w_lo, NP, CPG, G, T, SLS;
    float lr;
};

// ---- small PTX helpers ---------------------------------------------------------------------------
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
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// compute warps only (the helper warps never join): named barrier 1
__device__ __forceinline__ void compute_sync() { asm volatile("bar.sync 1, %0;" ::"n"(RES_CTHREADS) : "memory"); }

// tagged words: {payload bits, step number} in one 64-bit access (single-copy atomic)
__device__ __forceinline__ void st_tagged(unsigned long long* p, float v, unsigned seq) {
    asm volatile("{\n .reg .b64 t;\n mov.b64 t, {%1, %2};\n st.relaxed.gpu.global.u64 [%0], t;\n}"
                 ::"l"(p), "r"(__float_as_uint(v)), "r"(seq) : "memory");
}
__device__ __forceinline__ void ld_tagged(const unsigned long long* p, unsigned& bits, unsigned& seq) {
    asm volatile("{\n .reg .b64 t;\n ld.relaxed.gpu.global.u64 t, [%2];\n mov.b64 {%0, %1}, t;\n}"
                 : "=r"(bits), "=r"(seq) : "l"(p) : "memory");
}
// poll a tagged word until it carries step number `seq`; returns the payload (0 on abort)
__device__ __noinline__ float poll_word(const unsigned long long* src, unsigned seq, unsigned* abort_flag) {
    unsigned bits, tag, it = 0;
    ld_tagged(src, bits, tag);
    while (tag != seq) {
        if ((++it & 0x3ffu) == 0u) {
            if (*reinterpret_cast<volatile unsigned*>(abort_flag) != 0u) return 0.f;
            if (it > RES_SPIN_LIMIT) { atomicExch(abort_flag, 1u); return 0.f; }
        }
        ld_tagged(src, bits, tag);
    }
    return __uint_as_float(bits);
}

// one row (8 pixels) of a cell: gradient mass sent to the left / right low-res column of that row
__device__ __forceinline__ void hires_row(float left, float right, uint32_t rb, float c0, float c1,
                                          float& gl_out, float& gr_out) {
    const float nleft = NEG_LOG2E * left, nslope = NEG_LOG2E * (right - left) * 0.125f;
    float gs = 0.f, gr = 0.f;
#pragma unroll
    for (int s = 0; s < 8; ++s) {
        const float g = ce_grad_pixel(fmaf((float)s, nslope, nleft), rb & (3u << (2 * s)), 1u << (2 * s), c0, c1);
        gs += g;
        gr = fmaf((float)s, g, gr);
    }
    gr *= 0.125f;
    gl_out = gs - gr;
    gr_out = gr;
}

struct ResSmem {
    size_t F, W0, W1, Wd, zfull, g, scratch, mbar, total;
};
static __host__ __device__ inline ResSmem res_smem_layout(int C, int NP, int w_lo) {
    ResSmem s;
    const int NQ = NP / 4, NG = RES_CTHREADS / NQ, NCELL = NP + w_lo + 1;
    size_t o = 0;
    size_t sz[8];
    sz[0] = (size_t)C * NP * 4; sz[1] = sz[2] = sz[3] = (size_t)C * 4;
    sz[4] = (size_t)(NP + 2 * (w_lo + 1)) * 4; sz[5] = (size_t)NP * 4;
    const size_t sc1 = (size_t)NG * NP * 4, sc2 = (size_t)16 * NCELL * 4;
    sz[6] = sc1 > sc2 ? sc1 : sc2; sz[7] = 8 * (1 + RES_KCH);
    size_t off[8];
    for (int i = 0; i < 8; ++i) { off[i] = o; o = (o + sz[i] + 127) / 128 * 128; }
    s.F = off[0]; s.W0 = off[1]; s.W1 = off[2]; s.Wd = off[3]; s.zfull = off[4]; s.g = off[5];
    s.scratch = off[6]; s.mbar = off[7]; s.total = o;
    return s;
}

// TC/TNP/TWL/THL > 0: compile-time shape (C, strip pixels, low-res width / height); 0: run-time shape.
template <int TC, int TNP, int TWL, int THL, bool PROF>
__global__ void __launch_bounds__(RES_THREADS, 1) k_fit_resident(ResidentParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int C = TC ? TC : p.C, NP = TNP ? TNP : p.NP, wl = TWL ? TWL : p.w_lo, h = THL ? THL : p.h;
    const int HW = (TWL && THL) ? TWL * THL : p.HW;
    const int CPG = (TNP && TWL && THL) ? (TWL * THL) / (TNP ? TNP : 1) : p.CPG;
    const int CCH = C / RES_KCH;                     // channels per chunk (C % RES_KCH == 0)
    const int SLS = (CCH + CPG - 1) / CPG;           // channels of a chunk owned by one CTA
    constexpr bool kStatic = (TC == 512 && TNP == 100);          // PSPNet head geometry: fully unrolled sweeps
    const int NQ = NP / 4, NG = kStatic ? 16 : RES_CTHREADS / NQ, NCELL = NP + wl + 1, HALO = wl + 1;
    const ResSmem L = res_smem_layout(C, NP, wl);
    float* F = reinterpret_cast<float*>(smem_raw + L.F);
    float* W0 = reinterpret_cast<float*>(smem_raw + L.W0);
    float* W1 = reinterpret_cast<float*>(smem_raw + L.W1);
    float* Wd = reinterpret_cast<float*>(smem_raw + L.Wd);
    float* zfull = reinterpret_cast<float*>(smem_raw + L.zfull);
    float* gsm = reinterpret_cast<float*>(smem_raw + L.g);
    float* scratch = reinterpret_cast<float*>(smem_raw + L.scratch);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw + L.mbar);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool is_reducer = warp == RES_CTHREADS / 32, is_applier = warp == RES_CTHREADS / 32 + 1;
    const int group = blockIdx.x / CPG, k = blockIdx.x - group * CPG;
    const bool p1_active = tid < NQ * NG;
    const int v = tid % NQ, grp = tid / NQ;
    const int strip0 = k * NP;                       // first own pixel
    const int cell0 = strip0 - HALO;                 // pixel / cell index of zfull[0] / cell 0

    // group-private exchange areas (all words are {value, step}; step numbers start at 1)
    unsigned long long* zll = p.zll + (size_t)group * 2 * HW;
    unsigned long long* inbox = p.inbox + (size_t)group * 2 * RES_KCH * CPG * CPG * SLS;
    unsigned long long* sums = p.sums + (size_t)group * 2 * C;
    const unsigned inbox_chunk = (unsigned)(CPG * CPG * SLS);          // words per (parity, chunk)

    uint64_t* applied = mbar + 1;                     // [RES_KCH] chunk j of the previous step has been applied to Wd
    if (tid == 0) { mbar_init(mbar, 1); for (int j = 0; j < RES_KCH; ++j) mbar_init(&applied[j], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    unsigned gstep = 0, tma_parity = 0;
    bool ok = true;
    long long t_acc[6] = {0, 0, 0, 0, 0, 0};

    // P3 thread mapping: 4 threads per channel (adjacent lanes), interleaved pixel-quads -> conflict-free LDS.128
    const int p3_cl = tid >> 2, p3_part = tid & 3;
    // where this thread's partial of a chunk goes: owner CTA's inbox row of this CTA, slot of the channel
    const unsigned p3_inbox_off = (unsigned)(((p3_cl / SLS) * CPG + k) * SLS + (p3_cl % SLS));

    // compute warps: chunk j of global step gs has been folded into W0 / W1 / Wd by the applier warp
    long long t_store[RES_KCH] = {0, 0, 0, 0};        // PROF: when chunk j's partials left this CTA
    auto wait_applied = [&](int j, unsigned gs) {
        unsigned it = 0;
        if (PROF && tid == 0 && !mbar_try_wait(&applied[j], gs & 1u)) {      // had to wait: latency = now - store time
            while (!mbar_try_wait(&applied[j], gs & 1u)) { if (++it > RES_SPIN_LIMIT) break; }
            t_acc[5] += clock64() - t_store[j];
            t_acc[4] += 1;
            return;
        }
        while (!mbar_try_wait(&applied[j], gs & 1u)) {
            if ((++it & 0xfffu) == 0u) {
                if (*reinterpret_cast<volatile unsigned*>(p.abort_flag) != 0u) break;
                if (it > RES_SPIN_LIMIT) { atomicExch(p.abort_flag, 1u); break; }
            }
        }
    };

    for (int e = group; e < p.E; e += p.G) {
        // ---------------- stage the episode: strip of F via bulk-TMA, weights, HR task descriptors ----------------
        __syncthreads();                                           // previous episode is done with shared memory
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy reads before async-proxy writes
        if (tid == 0) mbar_expect_tx(mbar, (unsigned)(C * NP * 4));
        const float* fsrc = p.f_s + (size_t)e * C * HW + strip0;
        for (int c = tid; c < C; c += RES_THREADS) bulk_g2s(F + (size_t)c * NP, fsrc + (size_t)c * HW, (unsigned)(NP * 4), mbar);
        for (int c = tid; c < C; c += RES_THREADS) {
            const float a = p.w[((size_t)e * 2) * C + c], b = p.w[((size_t)e * 2 + 1) * C + c];
            W0[c] = a; W1[c] = b; Wd[c] = b - a;
        }
        // HR task descriptors (static for the episode): task = (cell ci, row r); 8 adjacent lanes = one cell.
        // desc = row label bits | dx << 16 | dy_flag << 17 | live << 18
        unsigned hr_desc[RES_MAXTASK];
#pragma unroll
        for (int m = 0; m < RES_MAXTASK; ++m) {
            const int task = m * RES_CTHREADS + tid;
            const int ci = task >> 3, r = task & 7;
            unsigned d = 0u;
            if (tid < RES_CTHREADS && ci < NCELL) {
                const int q = cell0 + ci;
                if (q >= 0) {
                    const uint4 bits = p.cells[(size_t)e * HW + q];
                    const uint32_t wsel = (r < 4) ? ((r < 2) ? bits.x : bits.y) : ((r < 6) ? bits.z : bits.w);
                    const uint32_t rb = (wsel >> ((r & 1) * 16)) & 0xffffu;
                    const int a = q / wl, b = q - a * wl;
                    if (rb != 0xAAAAu)                              // rows with only ignored pixels send nothing
                        d = rb | ((b + 1 < wl) ? (1u << 16) : 0u) | ((a + 1 < h) ? (1u << 17) : 0u) | (1u << 18);
                }
            }
            hr_desc[m] = d;
        }
        const float2 c01 = p.cw[e];
        {
            unsigned it = 0;
            while (!mbar_try_wait(mbar, tma_parity)) { if (++it > RES_SPIN_LIMIT) { ok = false; break; } }
            tma_parity ^= 1u;
        }
        __syncthreads();

        if (is_reducer) {
            // ================= reducer warp: hop 1 (collect + add the CPG partials of the own slice), hop 2 (republish) =================
            const int n_own = max(0, min(SLS, CCH - k * SLS));       // channels of every chunk owned by this CTA
            int slp = 1;
            while (slp < n_own) slp <<= 1;                            // lanes per source-CTA row (power of two)
            const int sl = lane % slp, kg = lane / slp, nkg = 32 / slp;
            for (int t = 0; t < p.T; ++t) {
                const unsigned gs = gstep + (unsigned)t;
                if (n_own > 0) {
                    for (int j = 0; j < RES_KCH; ++j) {
                        const unsigned long long* ib = inbox + ((gs & 1u) * RES_KCH + j) * inbox_chunk + (unsigned)(k * CPG * SLS);
                        float acc = 0.f;
                        if (sl < n_own) {
                            // every load of this lane is issued before any tag is looked at (independent L2 round
                            // trips); words that are not there yet are re-requested together, again as one batch
                            unsigned wb[RES_RW], wt[RES_RW];
                            unsigned pending = 0u, it = 0u;
#pragma unroll
                            for (int m = 0; m < RES_RW; ++m) { wb[m] = 0u; if (kg + m * nkg < CPG) pending |= 1u << m; }
                            while (pending) {
#pragma unroll
                                for (int m = 0; m < RES_RW; ++m)
                                    if (pending & (1u << m)) ld_tagged(&ib[(kg + m * nkg) * SLS + sl], wb[m], wt[m]);
#pragma unroll
                                for (int m = 0; m < RES_RW; ++m)
                                    if ((pending & (1u << m)) && wt[m] == gs + 1u) pending &= ~(1u << m);
                                if (pending && (++it & 0xffu) == 0u) {
                                    if (*reinterpret_cast<volatile unsigned*>(p.abort_flag) != 0u) break;
                                    if (it > (RES_SPIN_LIMIT >> 2)) { atomicExch(p.abort_flag, 1u); break; }
                                }
                            }
#pragma unroll
                            for (int m = 0; m < RES_RW; ++m) acc += __uint_as_float(wb[m]);       // fixed order: deterministic
                            for (int kk = kg + RES_RW * nkg; kk < CPG; kk += nkg)                  // (unusual shapes only)
                                acc += poll_word(&ib[kk * SLS + sl], gs + 1u, p.abort_flag);
                        }
                        for (int o = slp; o < 32; o <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                        if (lane < n_own) st_tagged(&sums[(gs & 1u) * (unsigned)C + j * CCH + k * SLS + lane], acc, gs + 1u);
                    }
                }
            }
            gstep += (unsigned)p.T;
        } else if (is_applier) {
            // ================= applier warp: sums -> SGD update in shared memory -> release the compute warps =================
            for (int t = 0; t < p.T; ++t) {
                const unsigned gs = gstep + (unsigned)t;
                for (int j = 0; j < RES_KCH; ++j) {
                    const unsigned long long* sw = sums + (gs & 1u) * (unsigned)C + j * CCH;
                    for (int c0 = 0; c0 < CCH; c0 += 32 * RES_AW) {
                        unsigned wb[RES_AW], wt[RES_AW];
                        unsigned pending = 0u, it = 0u;
#pragma unroll
                        for (int m = 0; m < RES_AW; ++m) { wb[m] = 0u; if (c0 + m * 32 + lane < CCH) pending |= 1u << m; }
                        while (pending) {
#pragma unroll
                            for (int m = 0; m < RES_AW; ++m)
                                if (pending & (1u << m)) ld_tagged(&sw[c0 + m * 32 + lane], wb[m], wt[m]);
#pragma unroll
                            for (int m = 0; m < RES_AW; ++m)
                                if ((pending & (1u << m)) && wt[m] == gs + 1u) pending &= ~(1u << m);
                            if (pending && (++it & 0xffu) == 0u) {
                                if (*reinterpret_cast<volatile unsigned*>(p.abort_flag) != 0u) break;
                                if (it > (RES_SPIN_LIMIT >> 2)) { atomicExch(p.abort_flag, 1u); break; }
                            }
                        }
#pragma unroll
                        for (int m = 0; m < RES_AW; ++m) {
                            const int cl = c0 + m * 32 + lane;
                            if (cl < CCH) {
                                const int c = j * CCH + cl;
                                const float dw = __uint_as_float(wb[m]);
                                const float n0 = fmaf(p.lr, dw, W0[c]), n1 = fmaf(-p.lr, dw, W1[c]);
                                W0[c] = n0; W1[c] = n1; Wd[c] = n1 - n0;
                            }
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&applied[j]);          // release.cta: the Wd stores above are visible to the waiters
                }
            }
            gstep += (unsigned)p.T;
        } else {
            // ================= compute warps =================
            for (int t = 0; t < p.T; ++t, ++gstep) {
                long long tk0 = 0;
                if (PROF && tid == 0) tk0 = clock64();
                // ------------ P1 (each chunk first picks up the previous step's all-reduced dW) ------------
                {
                    float4 za = make_float4(0.f, 0.f, 0.f, 0.f);
                    const float4* F4 = reinterpret_cast<const float4*>(F) + tid;
                    const int stride = NQ * NG;
                    int c = grp;
#pragma unroll
                    for (int j = 0; j < RES_KCH; ++j) {
                        if (t > 0) wait_applied(j, gstep - 1u);
                        if (p1_active) {
                            if constexpr (kStatic) {
                                // chunk j = channels grp + 16*(8j .. 8j+7): 8 independent LDS.128 + 8 broadcast LDS.32, immediate offsets
                                const float4* Fj = reinterpret_cast<const float4*>(F) + tid + j * (8 * 400);
                                const float* Wj = Wd + grp + j * 128;
                                float4 f[8];
                                float wd[8];
#pragma unroll
                                for (int u = 0; u < 8; ++u) { f[u] = Fj[u * 400]; wd[u] = Wj[u * 16]; }
#pragma unroll
                                for (int u = 0; u < 8; ++u) {
                                    za.x = fmaf(wd[u], f[u].x, za.x); za.y = fmaf(wd[u], f[u].y, za.y);
                                    za.z = fmaf(wd[u], f[u].z, za.z); za.w = fmaf(wd[u], f[u].w, za.w);
                                }
                            } else {
                                const int cend = (j + 1) * CCH;
#pragma unroll 8
                                for (; c < cend; c += NG) {
                                    const float4 f = *F4;
                                    F4 += stride;
                                    const float wd = Wd[c];
                                    za.x = fmaf(wd, f.x, za.x); za.y = fmaf(wd, f.y, za.y);
                                    za.z = fmaf(wd, f.z, za.z); za.w = fmaf(wd, f.w, za.w);
                                }
                            }
                        }
                    }
                    if (p1_active) *reinterpret_cast<float4*>(scratch + grp * NP + 4 * v) = za;
                    compute_sync();
                    if (tid < NP) {
                        float z = 0.f;
#pragma unroll 4
                        for (int g2 = 0; g2 < NG; ++g2) z += scratch[g2 * NP + tid];
                        zfull[HALO + tid] = z;
                        st_tagged(&zll[(gstep & 1u) * (unsigned)HW + strip0 + tid], z, gstep + 1u);
                    }
                }
                if (PROF && tid == 0) { long long n = clock64(); t_acc[0] += n - tk0; tk0 = n; }
                // ------------ X1: fetch both halos from the neighbouring strips ------------
                if (tid < 2 * HALO) {
                    const bool before = tid < HALO;
                    const int j = before ? tid : tid - HALO;
                    const int q = before ? cell0 + j : strip0 + NP + j;
                    float z = 0.f;
                    if (q >= 0 && q < HW) z = poll_word(&zll[(gstep & 1u) * (unsigned)HW + q], gstep + 1u, p.abort_flag);
                    zfull[before ? j : HALO + NP + j] = z;
                }
                compute_sync();
                if (PROF && tid == 0) { long long n = clock64(); t_acc[1] += n - tk0; tk0 = n; }
                // ------------ HR: one task = one row of one cell -> (gl, gr) of that row in shared memory ------------
#pragma unroll
                for (int m = 0; m < RES_MAXTASK; ++m) {
                    if (m * RES_CTHREADS < 8 * NCELL) {                     // uniform: does this round have tasks at all
                        const int task = m * RES_CTHREADS + tid;
                        const int ci = task >> 3, r = tid & 7;
                        const unsigned d = hr_desc[m];
                        float gl = 0.f, gr = 0.f;
                        if (d & (1u << 18)) {
                            const int dx = (d >> 16) & 1, dy = (d & (1u << 17)) ? wl : 0;
                            const float z00 = zfull[ci], z01 = zfull[ci + dx], z10 = zfull[ci + dy], z11 = zfull[ci + dy + dx];
                            const float fr = (float)r * 0.125f;
                            hires_row(fmaf(fr, z10 - z00, z00), fmaf(fr, z11 - z01, z01), d & 0xffffu, c01.x, c01.y, gl, gr);
                        }
                        if (ci < NCELL) *reinterpret_cast<float2*>(scratch + 2 * task) = make_float2(gl, gr);   // [cell][row][gl,gr]
                    }
                }
                compute_sync();
                // g(q) = sum_r (1-r/8) [gl(q,r) + gr(q-1,r)] + (r/8) [gl(q-w,r) + gr(q-w-1,r)]; 4 threads per own pixel, 2 rows each
                {
                    const int pl = tid >> 2, rq = tid & 3;
                    float s2 = 0.f;
                    if (pl < NP) {
                        const int ci = HALO + pl;
                        const int q = strip0 + pl;
                        const int a = q / wl, b = q - a * wl;
#pragma unroll
                        for (int rr = 0; rr < 2; ++rr) {
                            const int r = 2 * rq + rr;
                            const float h1 = (float)r * 0.125f, h0 = 1.f - h1;
                            float top = scratch[(ci * 8 + r) * 2];                               // gl(q, r)
                            if (b > 0) top += scratch[((ci - 1) * 8 + r) * 2 + 1];               // gr(q-1, r)
                            float bot = 0.f;
                            if (a > 0) {
                                bot = scratch[((ci - wl) * 8 + r) * 2];                          // gl(q-w, r)
                                if (b > 0) bot += scratch[((ci - wl - 1) * 8 + r) * 2 + 1];      // gr(q-w-1, r)
                            }
                            s2 = fmaf(h0, top, fmaf(h1, bot, s2));
                        }
                    }
                    s2 += __shfl_xor_sync(0xffffffffu, s2, 1);
                    s2 += __shfl_xor_sync(0xffffffffu, s2, 2);
                    if (pl < NP && rq == 0) gsm[pl] = s2;
                }
                compute_sync();
                if (PROF && tid == 0) { long long n = clock64(); t_acc[2] += n - tk0; tk0 = n; }
                // ------------ P3: dW = g . F^T chunk by chunk; every partial goes straight to its owner's inbox ------------
                {
                    // this thread's pixel-quads of g are the same for every channel: keep them in registers
                    float4 gq[RES_MAXQ];
#pragma unroll
                    for (int i = 0; i < RES_MAXQ / 2; ++i) {
                        const int q0 = 2 * p3_part + 8 * i;
                        gq[2 * i] = (q0 < NQ) ? reinterpret_cast<const float4*>(gsm)[q0] : make_float4(0.f, 0.f, 0.f, 0.f);
                        gq[2 * i + 1] = (q0 + 1 < NQ) ? reinterpret_cast<const float4*>(gsm)[q0 + 1] : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    unsigned long long* ibw = inbox + (gstep & 1u) * RES_KCH * inbox_chunk + p3_inbox_off;
#pragma unroll
                    for (int j = 0; j < RES_KCH; ++j) {
                        for (int cb = 0; cb < CCH; cb += RES_CTHREADS / 4) {
                            const int cl = cb + p3_cl;
                            float d0 = 0.f, d1 = 0.f;
                            if (cl < CCH) {
                                const float4* row = reinterpret_cast<const float4*>(F + (size_t)(j * CCH + cl) * NP) + 2 * p3_part;
                                if constexpr (kStatic) {
                                    // 25 quads per channel row: parts own quads {2p, 2p+1} + 8i for i = 0..2; quad 24 belongs to part 0
                                    float4 f[6];
#pragma unroll
                                    for (int i = 0; i < 3; ++i) { f[2 * i] = row[8 * i]; f[2 * i + 1] = row[8 * i + 1]; }
                                    float4 f6 = make_float4(0.f, 0.f, 0.f, 0.f);
                                    if (p3_part == 0) f6 = row[24];
#pragma unroll
                                    for (int i = 0; i < 6; ++i) {
                                        d0 = fmaf(gq[i].x, f[i].x, d0); d1 = fmaf(gq[i].y, f[i].y, d1);
                                        d0 = fmaf(gq[i].z, f[i].z, d0); d1 = fmaf(gq[i].w, f[i].w, d1);
                                    }
                                    d0 = fmaf(gq[6].x, f6.x, d0); d1 = fmaf(gq[6].y, f6.y, d1);
                                    d0 = fmaf(gq[6].z, f6.z, d0); d1 = fmaf(gq[6].w, f6.w, d1);
                                } else {
#pragma unroll
                                    for (int i = 0; i < RES_MAXQ / 2; ++i) {
                                        const int q0 = 2 * p3_part + 8 * i;
                                        if (q0 < NQ) {
                                            const float4 f = row[8 * i];
                                            d0 = fmaf(gq[2 * i].x, f.x, d0); d1 = fmaf(gq[2 * i].y, f.y, d1);
                                            d0 = fmaf(gq[2 * i].z, f.z, d0); d1 = fmaf(gq[2 * i].w, f.w, d1);
                                        }
                                        if (q0 + 1 < NQ) {
                                            const float4 f = row[8 * i + 1];
                                            d0 = fmaf(gq[2 * i + 1].x, f.x, d0); d1 = fmaf(gq[2 * i + 1].y, f.y, d1);
                                            d0 = fmaf(gq[2 * i + 1].z, f.z, d0); d1 = fmaf(gq[2 * i + 1].w, f.w, d1);
                                        }
                                    }
                                }
                            }
                            float d = d0 + d1;
                            d += __shfl_xor_sync(0xffffffffu, d, 1);
                            d += __shfl_xor_sync(0xffffffffu, d, 2);
                            if (cl < CCH && p3_part == 0) {
                                const int off = (cb == 0) ? 0 : (((cl / SLS) * CPG + k) * SLS + (cl % SLS)) - (int)p3_inbox_off;
                                st_tagged(ibw + (int)(j * inbox_chunk) + off, d, gstep + 1u);
                            }
                        }
                        if (PROF && tid == 0) t_store[j] = clock64();
                    }
                }
                if (PROF && tid == 0) { long long n = clock64(); t_acc[3] += n - tk0; tk0 = n; }
            }
            // ------------ drain the last step's all-reduce ------------
            if (p.T > 0) {
                long long tk0 = 0;
                if (PROF && tid == 0) tk0 = clock64();
#pragma unroll
                for (int j = 0; j < RES_KCH; ++j) wait_applied(j, gstep - 1u);
            }
        }
        __syncthreads();
        if (k == 0) {
            const bool bad = *reinterpret_cast<volatile unsigned*>(p.abort_flag) != 0u;
            for (int c = tid; c < C; c += RES_THREADS) {
                p.w[((size_t)e * 2) * C + c] = bad ? __int_as_float(0x7fc00000) : W0[c];
                p.w[((size_t)e * 2 + 1) * C + c] = bad ? __int_as_float(0x7fc00000) : W1[c];
            }
        }
    }
    if (PROF && tid == 0 && p.prof) {
        for (int i = 0; i < 6; ++i) p.prof[(size_t)blockIdx.x * 8 + i] = t_acc[i];
    }
    if (tid == 0 && !ok) atomicExch(p.abort_flag, 1u);
}

// ---- host side ---------------------------------------------------------------------------------
struct ResidentPlan { int NP, CPG, G; size_t smem; bool ok; int SLS; size_t inbox_words; };

static ResidentPlan plan_resident(int E, int C, int h, int w, int n_sm, size_t smem_cap) {
    ResidentPlan best{0, 0, 0, 0, false, 0, 0};
    const int HW = h * w;
    int forced = 0;
    if (const char* s = getenv("CWT_RESIDENT_NP")) forced = atoi(s);
    double best_score = -1.0;
    for (int NP = 4; NP <= HW; NP += 4) {
        if (HW % NP) continue;
        if (forced && NP != forced) continue;
        if (NP < w + 1 && HW != NP) continue;          // halos must come from the adjacent strips only
        if (NP > 16 * RES_MAXQ) continue;              // P3 register tiling
        if (8 * (NP + w + 1) > RES_MAXTASK * RES_CTHREADS) continue;   // HR task descriptors
        if (C % RES_KCH) continue;
        const size_t sm = res_smem_layout(C, NP, w).total;
        if (sm > smem_cap) continue;
        const int CPG = HW / NP;
        if (CPG > n_sm) continue;
        int G = n_sm / CPG;
        if (G > E) G = E;
        const double score = (double)G / NP + 1e-9 * NP;
        const int CCH = C / RES_KCH, SLS = (CCH + CPG - 1) / CPG;
        if (score > best_score) { best_score = score; best = ResidentPlan{NP, CPG, G, sm, true, SLS, (size_t)2 * RES_KCH * CPG * CPG * SLS}; }
    }
    return best;
}

size_t fit_resident_workspace_bytes(int C, int h, int w) {
    const int HW = h * w;
    const int maxG = 148;
    // zll: 2*HW words per group; inbox: 2*KCH*CPG*CPG*SLS words per group (CPG*SLS < C/KCH + CPG, G*CPG <= 148);
    // sums: 2*C words per group; abort flag; profile counters
    const size_t inbox_words_all = (size_t)2 * RES_KCH * 148 * ((size_t)C / RES_KCH + 148);
    return align_up(sizeof(unsigned long long) * 2 * HW * maxG) + align_up(sizeof(unsigned long long) * inbox_words_all) +
           align_up(sizeof(unsigned long long) * 2 * (size_t)C * maxG) + 512 + align_up(sizeof(long long) * 8 * 160);
}

template <int TC, int TNP, int TWL, int THL>
static int launch_resident(const ResidentParams& p, const ResidentPlan& pl, bool prof, cudaStream_t st) {
    void* args[] = {const_cast<ResidentParams*>(&p)};
    dim3 grid(pl.G * pl.CPG), block(RES_THREADS);
    const void* fn = prof ? (const void*)k_fit_resident<TC, TNP, TWL, THL, true> : (const void*)k_fit_resident<TC, TNP, TWL, THL, false>;
    CWT_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    CWT_CUDA(cudaLaunchCooperativeKernel(fn, grid, block, args, pl.smem, st));
    count_launch();
    return CWT_OK;
}

// returns CWT_ERR_UNSUPPORTED when the shape does not fit on chip
int fit_resident(const float* f_s, const uint4* cells, const float2* cw, float* w_io, int E, int C, int h, int w,
                 int n_iter, float lr, void* ws, size_t ws_bytes, long long* prof_out, cudaStream_t st) {
    int dev = 0, n_sm = 0, smem_cap = 0, coop = 0;
    CWT_CUDA(cudaGetDevice(&dev));
    CWT_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    CWT_CUDA(cudaDeviceGetAttribute(&smem_cap, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    CWT_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    CWT_REQUIRE(coop, CWT_ERR_UNSUPPORTED, "fit_resident: device lacks cooperative launch");
    CWT_REQUIRE(C <= RES_CTHREADS * RES_KCH, CWT_ERR_UNSUPPORTED, "fit_resident: C=%d too large", C);
    if (n_sm > 148) n_sm = 148;
    const ResidentPlan pl = plan_resident(E, C, h, w, n_sm, (size_t)smem_cap);
    CWT_REQUIRE(pl.ok, CWT_ERR_UNSUPPORTED, "fit_resident: no strip size fits C=%d, %dx%d in %d B of shared memory", C, h, w, smem_cap);
    const int HW = h * w;
    Carver cv(ws, ws_bytes);
    ResidentParams p{};
    p.zll = cv.take<unsigned long long>((size_t)2 * HW * pl.G);
    p.inbox = cv.take<unsigned long long>(pl.inbox_words * pl.G);
    p.sums = cv.take<unsigned long long>((size_t)2 * C * pl.G);
    p.abort_flag = cv.take<unsigned>(64);
    p.prof = prof_out ? cv.take<long long>((size_t)8 * pl.G * pl.CPG) : nullptr;
    CWT_REQUIRE(ws && cv.ok(), CWT_ERR_WORKSPACE, "fit_resident: workspace too small");
    // every tagged word and the abort flag start at zero (step numbers start at 1)
    const size_t sync_bytes = (size_t)(reinterpret_cast<char*>(p.abort_flag + 64) - reinterpret_cast<char*>(p.zll));
    CWT_CUDA(cudaMemsetAsync(p.zll, 0, sync_bytes, st));
    p.f_s = f_s; p.cells = cells; p.cw = cw; p.w = w_io;
    p.E = E; p.C = C; p.HW = HW; p.h = h; p.w_lo = w; p.NP = pl.NP; p.CPG = pl.CPG; p.G = pl.G; p.T = n_iter; p.lr = lr; p.SLS = pl.SLS;
    int rc;
    if (C == 512 && h == 60 && w == 60 && pl.NP == 100) rc = launch_resident<512, 100, 60, 60>(p, pl, prof_out != nullptr, st);
    else rc = launch_resident<0, 0, 0, 0>(p, pl, prof_out != nullptr, st);
    if (rc != CWT_OK) return rc;
    if (prof_out) CWT_CUDA(cudaMemcpyAsync(prof_out, p.prof, sizeof(long long) * 8 * pl.G * pl.CPG, cudaMemcpyDeviceToDevice, st));
    return CWT_OK;
}

}  // namespace cwt
