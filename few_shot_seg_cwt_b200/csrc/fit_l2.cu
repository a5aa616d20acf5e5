// CWT_FIT_L2 — the support-classifier fit for S > 1 shots as ONE persistent cooperative kernel whose feature maps stay in L2.
//
// (a-3 with `shot 5`: src/test.py:177-187 — f_s is [n_shots, c, h, w], the classifier is fitted on all shots at once, the
//  loss is the weighted mean over the pixels of all shots; BASELINE.json configs[1].)
//
// A 5-shot episode holds 5 x 7.37 MB of fp32 features: too much for the shared memory of a group of SMs (fit_resident.cu needs
// one SM per 204.8 KB tile), but two episodes (73.7 MB) fit the 126 MB L2, whose read bandwidth (17.9 TB/s measured) is 2.7 x
// that of HBM. The streaming algorithm (three launches per SGD step over the whole batch) re-reads every map from HBM twice per
// step; here G episodes at a time are fitted to completion by G groups of CPG co-resident CTAs, so that their maps are fetched
// from HBM once and then served by L2 for the 2 T sweeps.
//
// The step is fit_resident.cu's, with the tile loop and the data movement changed:
//   * an episode's S images are cut into 36 S tiles of 20 x 5 low-res pixels; CTA k of the group owns tiles k, k + CPG, ...
//     (NT per CTA; 3 for 5 shots with CPG = 60);
//   * a PRODUCER WARP streams the tiles' features through a 3-stage shared-memory ring, one `cp.async.bulk.tensor.3d` per
//     (tile, 128-channel chunk) = 51.2 KB (tensor map over f_s as [E S C][60][60], box 20 x 5 x 128; SASS UTMALDG.3D), in the
//     order the sweeps consume them; full / empty mbarriers pace it, and it runs ahead of the compute warps across phases and
//     steps (the features never change), so the ring is full again when a sweep starts;
//   * P1 (z = Wd . F) goes chunk-major over the CTA's tiles — chunk j of all tiles, then chunk j + 1 — so that, exactly as in
//     the resident kernel, chunk j of the previous step's all-reduce is only needed when its first sweep starts;
//   * the halo exchange (sequence-tagged 8-byte words), the full-resolution stage on the 126 cells around every tile, the
//     adjoint gather and the applier warp (64-bit L2 accumulator words {fixed-point sum | arrival count}, two parity sets) are
//     the resident kernel's, looped over the NT tiles;
//   * P3 (dW = g . F^T): warp quad j owns channel chunk j of every tile, accumulates over the tiles in registers and sends its
//     128 partials with one `red.add.u64` each when its last tile is done.
// The max |F| that fixes the fixed-point unit of the all-reduce comes from one extra sweep at the start of the episode.
//
// Specialised for the PSPNet head geometry (C = 512, 60 x 60 features, tile 20 x 5); other shapes use the streaming algorithm.
#include "common.cuh"
#include "hires.cuh"
#include "resident_common.cuh"
#include "tma_pipe.cuh"
#include <cstdio>
#include <cstdlib>

namespace cwt {

constexpr int L2_C = 512, L2_TW = 20, L2_TH = 5, L2_NP = 100, L2_WL = 60, L2_HL = 60, L2_HW = 3600;
constexpr int L2_TPR = 3, L2_TPI = 36;              // tiles per row of tiles / per image
constexpr int L2_KCH = 4, L2_CCH = 128;             // channel chunks
constexpr int L2_NSTAGE = 3;
constexpr unsigned L2_STAGE_BYTES = L2_CCH * L2_NP * 4;      // 51 200
constexpr int L2_CT = 512;                          // compute threads
constexpr int L2_THREADS = L2_CT + 96;              // + applier, halo, producer warps
constexpr int L2_ZW = L2_TW + 2, L2_ZH = L2_TH + 2, L2_CW = L2_TW + 1, L2_NCELL = L2_CW * (L2_TH + 1);    // 22, 7, 21, 126
constexpr int L2_NINT = (L2_TW - 1) * (L2_TH - 1);  // 76 interior cells
constexpr int L2_NRING = 2 * L2_ZW + 2 * (L2_ZH - 2);        // 54
constexpr int L2_MAXNT = 4;
// L2_TM: the FIRST tile of every CTA lives in the SM's tensor memory for the steps (tcgen05.st once per episode, tcgen05.ld every
// sweep: lane = channel mod 128, column = 100 (channel / 128) + pixel, exactly as in fit_resident.cu) and is never streamed again:
// with 3 tiles per CTA (5 shots) a third of the per-step L2 traffic — the resource this kernel is bound by — disappears.
#ifndef L2_TM
#define L2_TM 1
#endif
// the halo warp sleeps on named barrier 5 until its CTA has published the step's z (fit_resident.cu's RES_HALO_GATE): the
// neighbours publish about then, earlier polls of the ring are only traffic on lines that are being written
#ifndef L2_HALO_GATE
#define L2_HALO_GATE 1
#endif

struct L2Params {
    const uint4* cells;        // [E][S][HW]
    const float2* cw;          // [E]
    float* w;                  // [E][2][C]  in: W0, out: fitted
    unsigned long long* zll;   // [G][2][S*HW]        {z, step} words of the halo exchange
    unsigned long long* sums;  // [G][SPL][2][C]      accumulator words
    unsigned long long* fmaxw; // [G][CPG]            {max|F| bits, episode + 1}
    unsigned* abort_flag;
    float* ftile;              // [G][NTILE][KCH][128][100]  the episode in flight, re-laid out tile by tile (see below)
    int E, S, CPG, G, T, KBITS, SPL, NTILE;
    float lr;
};

struct L2Smem { size_t ring, W0, W1, Wd, zt, g, scratch, cum, hrd, mbar, total; };
static __host__ __device__ inline L2Smem l2_smem_layout(int NT) {
    L2Smem s;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o = (o + bytes + 127) / 128 * 128; return at; };
    s.ring = take((size_t)L2_NSTAGE * L2_STAGE_BYTES);
    s.W0 = take(L2_C * 4); s.W1 = take(L2_C * 4); s.Wd = take(L2_C * 4);
    s.zt = take((size_t)NT * 640);                              // NT x [7][22] floats (616 B)
    s.g = take((size_t)NT * 512);                               // NT x [100] floats
    s.scratch = take((size_t)NT * 16 * L2_NCELL * 4);           // NT x {16 x 100 P1 partials | 126 cells x 8 rows x (gl, gr)}
    s.cum = take((size_t)L2_C * 16);
    s.hrd = take((size_t)NT * 2 * L2_CT * 4);                   // HR task descriptors [NT][2][512] (static per episode)
    s.mbar = take(8 * (2 * L2_NSTAGE + L2_KCH + 2) + 16);
    s.total = o;
    return s;
}

// CTA-wide barrier of the compute, applier and halo warps (the producer warp free-runs: it is paced by the ring alone)
__device__ __forceinline__ void l2_role_sync() { asm volatile("bar.sync 6, %0;" ::"n"(L2_CT + 64) : "memory"); }

template <int NT>
__global__ void __launch_bounds__(L2_THREADS, 1) k_fit_l2(const __grid_constant__ CUtensorMap fmap, L2Params p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const L2Smem L = l2_smem_layout(NT);
    float* ring = reinterpret_cast<float*>(smem_raw + L.ring);
    float* W0 = reinterpret_cast<float*>(smem_raw + L.W0);
    float* W1 = reinterpret_cast<float*>(smem_raw + L.W1);
    float* Wd = reinterpret_cast<float*>(smem_raw + L.Wd);
    float* zt = reinterpret_cast<float*>(smem_raw + L.zt);            // [NT][160]
    float* gsm = reinterpret_cast<float*>(smem_raw + L.g);            // [NT][128]
    float* scratch = reinterpret_cast<float*>(smem_raw + L.scratch);  // [NT][2016]
    long long* cum = reinterpret_cast<long long*>(smem_raw + L.cum);
    unsigned* hrd = reinterpret_cast<unsigned*>(smem_raw + L.hrd);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + L.mbar);
    uint64_t* empty = full + L2_NSTAGE;
    uint64_t* applied = empty + L2_NSTAGE;                             // [KCH]
    uint64_t* halo_ready = applied + L2_KCH;
    uint64_t* retiled = halo_ready + 1;                                 // this CTA's tiles of the episode are in p.ftile
    unsigned* smax = reinterpret_cast<unsigned*>(retiled + 1);         // [0] CTA max|F| bits, [1] episode max|F| bits
    constexpr int ZTS = 160, GS = 128, SCS = 16 * L2_NCELL;           // per-tile strides (floats)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int group = blockIdx.x / p.CPG, k = blockIdx.x - group * p.CPG;
    const int S = p.S, KB = p.KBITS;
    // this CTA's tiles: global tile gt = k + i * CPG -> (shot, tile row, tile column)
    int nta = 0;
#pragma unroll
    for (int i = 0; i < NT; ++i) nta += (k + i * p.CPG < p.NTILE) ? 1 : 0;
    auto tile_shot = [&](int i) { return (k + i * p.CPG) / L2_TPI; };
    auto tile_y0 = [&](int i) { return (((k + i * p.CPG) % L2_TPI) / L2_TPR) * L2_TH; };
    auto tile_x0 = [&](int i) { return (((k + i * p.CPG) % L2_TPI) % L2_TPR) * L2_TW; };

    unsigned long long* zll = p.zll + (size_t)group * 2 * S * L2_HW;
    unsigned long long* fmaxw = p.fmaxw + (size_t)group * p.CPG;

    __shared__ uint32_t tmem_slot;
    if (tid == 0) {
        for (int s = 0; s < L2_NSTAGE; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 16); }
        for (int j = 0; j < L2_KCH; ++j) mbar_init(&applied[j], 1);
        mbar_init(halo_ready, 1);
        mbar_init(retiled, 1);
    }
    if (L2_TM && warp == 0) tmem_alloc_512(&tmem_slot);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (L2_TM) tmem_fence_before_sync();
    __syncthreads();
    if (L2_TM) tmem_fence_after_sync();
    const uint32_t tq = L2_TM ? tmem_slot + ((uint32_t)((warp & 3) * 32) << 16) : 0u;     // this warp's lane quarter of tensor memory
    constexpr int I0 = L2_TM ? 1 : 0;                  // first tile that is streamed in the steps (tile 0: tensor memory)

    const int units_per_sweep = L2_KCH * nta;          // max|F| sweep: every tile
    const int units_per_step_sweep = L2_KCH * (nta - I0);
    // every role walks the same unit sequence: per episode one max|F| sweep, then per step the P1 sweep and the P3 sweep,
    // each of KCH * nta units in chunk-major order (unit u of a sweep = chunk u / nta, tile u % nta)

    if (warp < L2_CT / 32) {
        // =====================================================================================================
        // compute warps
        // =====================================================================================================
        unsigned gstep = 0, uc = 0;                    // global step counter (tags / mbarrier phases), consumed units
        const bool p1_active = tid < 400;
        const int v = tid % 25, grp = tid / 25;
        auto wait_full = [&](unsigned u) {
            const unsigned st = u % L2_NSTAGE, ph = (u / L2_NSTAGE) & 1u;
            unsigned it = 0;
            while (!mbar_try_wait(&full[st], ph)) {
                if ((++it & 0xfffu) == 0u) {
                    if (*reinterpret_cast<volatile unsigned*>(p.abort_flag) != 0u) break;
                    if (it > RES_SPIN_LIMIT) { atomicExch(p.abort_flag, 1u); break; }
                }
            }
            return st;
        };
        auto release = [&](unsigned st, unsigned count) {     // one arrival per warp (count 1) or per quad warp (count 4)
            __syncwarp();
            if (lane == 0)
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&empty[st])), "r"(count) : "memory");
        };
        auto wait_applied = [&](int j, unsigned gs) {
            unsigned it = 0;
            while (!mbar_try_wait(&applied[j], gs & 1u)) {
                if ((++it & 0xfffu) == 0u) {
                    if (*reinterpret_cast<volatile unsigned*>(p.abort_flag) != 0u) break;
                    if (it > RES_SPIN_LIMIT) { atomicExch(p.abort_flag, 1u); break; }
                }
            }
        };

        for (int e = group; e < p.E; e += p.G) {
            l2_role_sync();                                           // S1: previous episode is done with shared memory
            for (int c = tid; c < L2_C; c += L2_CT) {
                const float a = p.w[((size_t)e * 2) * L2_C + c], b = p.w[((size_t)e * 2 + 1) * L2_C + c];
                W0[c] = a; W1[c] = b; Wd[c] = b - a;
            }
            if (tid < 2) smax[tid] = 0u;
            // HR task descriptors of every tile (static for the episode): task = (cell, row r); 8 adjacent lanes = one cell;
            // interior cells first. desc = row label bits | dx << 16 | dy_flag << 17 | live << 18 | cell << 19 | cell_row << 27 | valid << 31
            unsigned hr_uniform = 0u;
#pragma unroll
            for (int i = 0; i < NT; ++i) {
                const int sh = tile_shot(i), y0 = tile_y0(i), x0 = tile_x0(i);
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    const int task = m * L2_CT + tid;
                    const int o = task >> 3, r = task & 7;
                    unsigned d = 0u;
                    if (i < nta && o < L2_NCELL) {
                        int cy, cx;
                        if (o < L2_NINT) { cy = 1 + o / (L2_TW - 1); cx = 1 + o % (L2_TW - 1); }
                        else {
                            const int bo = o - L2_NINT;
                            if (bo < L2_CW) { cy = 0; cx = bo; }
                            else if (bo < 2 * L2_CW) { cy = L2_TH; cx = bo - L2_CW; }
                            else if (bo < 2 * L2_CW + L2_TH - 1) { cy = 1 + (bo - 2 * L2_CW); cx = 0; }
                            else { cy = 1 + (bo - 2 * L2_CW - (L2_TH - 1)); cx = L2_TW; }
                        }
                        d = (1u << 31) | ((unsigned)(cy * L2_CW + cx) << 19) | ((unsigned)cy << 27);
                        const int a = y0 - 1 + cy, b = x0 - 1 + cx;
                        if (a >= 0 && b >= 0) {
                            const uint4 bits = p.cells[((size_t)e * S + sh) * L2_HW + a * L2_WL + b];
                            const uint32_t wsel = (r < 4) ? ((r < 2) ? bits.x : bits.y) : ((r < 6) ? bits.z : bits.w);
                            const uint32_t rb = (wsel >> ((r & 1) * 16)) & 0xffffu;
                            if (rb != 0xAAAAu)                              // rows with only ignored pixels send nothing
                                d |= rb | ((b + 1 < L2_WL) ? (1u << 16) : 0u) | ((a + 1 < L2_HL) ? (1u << 17) : 0u) | (1u << 18);
                        }
                    }
                    hrd[(i * 2 + m) * L2_CT + tid] = d;              // (read back by the same thread only)
                    const uint32_t rbits = d & 0xffffu;
                    if (__all_sync(0xffffffffu, !(d & (1u << 18)) || rbits == 0u || rbits == 0x5555u)) hr_uniform |= 1u << (2 * i + m);
                }
            }
            const float2 c01 = p.cw[e];
            l2_role_sync();                                           // S2
            // ---------------- max|F| sweep: fixed-point unit of the episode ----------------
            {
                unsigned mb = 0u;
                for (int u = 0; u < units_per_sweep; ++u, ++uc) {
                    const unsigned st = wait_full(uc);
                    if (L2_TM && u % nta == 0 && (warp >> 2) == u / nta) {
                        // tile 0, chunk j = u / nta: the quad of warps j moves its 128 channel rows into tensor memory (thread = channel
                        // row, 100 columns); rows are 400 B apart: conflict-free LDS.128
                        const uint32_t row = smem_u32(ring + (size_t)st * (L2_STAGE_BYTES / 4)) + (uint32_t)(tid & 127) * 400u;
                        const uint32_t tcol = tq + 100u * (uint32_t)(warp >> 2);
                        float v[32];
#pragma unroll
                        for (int b = 0; b < 4; ++b) {
#pragma unroll
                            for (int q4 = 0; q4 < (b < 3 ? 8 : 1); ++q4) {
                                const float4 f = lds128_v(row + (uint32_t)(b * 128 + q4 * 16));
                                v[4 * q4] = f.x; v[4 * q4 + 1] = f.y; v[4 * q4 + 2] = f.z; v[4 * q4 + 3] = f.w;
                            }
                            if (b < 3) tmem_st32(tcol + 32u * b, v); else tmem_st4(tcol + 96u, v);
                        }
                    }
                    if (tid == 0 && !(L2_TM && u % nta == 0)) {
                        // first sweep of the episode: the unit came through the 3-D tensor map (640 rows of 80 B: slow); write it
                        // back CONTIGUOUSLY, so that the 2 T sweeps that follow fetch it with one linear 51.2 KB bulk copy
                        const int j = u / nta, i = u - j * nta;
                        float* dst = p.ftile + (((size_t)group * p.NTILE + (k + i * p.CPG)) * L2_KCH + j) * (L2_STAGE_BYTES / 4);
                        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                                     ::"l"(dst), "r"(smem_u32(ring + (size_t)st * (L2_STAGE_BYTES / 4))), "r"(L2_STAGE_BYTES) : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // the stage has been read: it may be refilled
                    }
                    const uint4* F4u = reinterpret_cast<const uint4*>(ring + (size_t)st * (L2_STAGE_BYTES / 4));
                    for (int i = tid; i < L2_CCH * 25; i += L2_CT) {
                        const uint4 f = F4u[i];
                        mb = max(max(mb, f.x & 0x7fffffffu), max(max(f.y & 0x7fffffffu, f.z & 0x7fffffffu), f.w & 0x7fffffffu));
                    }
                    release(st, 1u);
                }
                mb = __reduce_max_sync(0xffffffffu, mb);
                if (lane == 0) atomicMax(&smax[0], mb);
                if (tid == 0) {
                    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");               // ... and written: the producer may fetch it
                    mbar_arrive(retiled);
                }
                if (L2_TM) { tmem_wait_st(); tmem_fence_before_sync(); }
            }
            l2_role_sync();                                           // S3
            if (L2_TM) tmem_fence_after_sync();
            if (tid == 0) st_tagged(&fmaxw[k], __uint_as_float(smax[0]), (unsigned)e + 1u);
            for (int kk = tid; kk < p.CPG; kk += L2_CT)
                atomicMax(&smax[1], __float_as_uint(poll_word(&fmaxw[kk], (unsigned)e + 1u, p.abort_flag)));
            l2_role_sync();                                           // S4
            const bool ep_finite = smax[1] < 0x7f800000u;
            float fx_inv;
            { float unit; fixed_point_unit(smax[1], p.T, KB, unit, fx_inv); }
            unsigned long long* acc_ep = p.sums + acc_offset(e, group, p.G, p.SPL, L2_C);

            for (int t = 0; t < p.T; ++t, ++gstep) {
                // ------------ P1: z_i = Wd . F_i, chunk-major over the tiles ------------
                float4 za[NT];
#pragma unroll
                for (int i = 0; i < NT; ++i) za[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                // tile 0 from tensor memory (L2_TM): thread = lane 32 (warp % 4) + lane of the quarter, pixel block warp / 4 (25 pixels);
                // in-thread FFMA2 over the chunks, then a butterfly reduce-scatter over the lanes (fit_resident.cu's P1)
                f32x2 z2[13];
#pragma unroll
                for (int i = 0; i < 13; ++i) z2[i] = 0ull;
#pragma unroll
                for (int j = 0; j < L2_KCH; ++j) {
                    float fj[26];
                    if (L2_TM) {                           // the columns do not depend on the all-reduce: requested before the wait
                        const uint32_t tcol = tq + 100u * j + 25u * (uint32_t)(warp >> 2);
                        tmem_ld16(tcol, fj); tmem_ld8(tcol + 16u, fj + 16); tmem_ld2(tcol + 24u, fj + 24);
                    }
                    if (t > 0) wait_applied(j, gstep - 1u);
                    if (L2_TM) {
                        const float wl = Wd[j * 128 + (warp & 3) * 32 + lane];
                        const f32x2 wl2 = pk2(wl, wl);
                        tmem_wait_ld();
#pragma unroll
                        for (int i = 0; i < 26; ++i) asm volatile("" : "+f"(fj[i]));      // keeps the uses below the tcgen05.wait::ld
#pragma unroll
                        for (int i = 0; i < 13; ++i) z2[i] = fma2(wl2, pk2(fj[2 * i], fj[2 * i + 1]), z2[i]);
                    }
                    float wd[8];
                    if (p1_active) {
#pragma unroll
                        for (int u = 0; u < 8; ++u) wd[u] = Wd[grp + j * 128 + u * 16];
                    }
#pragma unroll
                    for (int i = I0; i < NT; ++i) {
                        if (i < nta) {
                            const unsigned st = wait_full(uc);
                            ++uc;
                            if (p1_active) {
                                const float4* Fj = reinterpret_cast<const float4*>(ring + (size_t)st * (L2_STAGE_BYTES / 4)) + tid;
#pragma unroll
                                for (int hh = 0; hh < 2; ++hh) {         // two batches of 4 independent 128-bit loads
                                    float4 f[4];
#pragma unroll
                                    for (int u = 0; u < 4; ++u) f[u] = Fj[(hh * 4 + u) * 400];
#pragma unroll
                                    for (int u = 0; u < 4; ++u) {
                                        const float wv = wd[hh * 4 + u];
                                        za[i].x = fmaf(wv, f[u].x, za[i].x); za[i].y = fmaf(wv, f[u].y, za[i].y);
                                        za[i].z = fmaf(wv, f[u].z, za[i].z); za[i].w = fmaf(wv, f[u].w, za[i].w);
                                    }
                                }
                            }
                            release(st, 1u);
                        }
                    }
                }
#pragma unroll
                for (int i = I0; i < NT; ++i)
                    if (i < nta && p1_active) *reinterpret_cast<float4*>(scratch + i * SCS + grp * L2_NP + 4 * v) = za[i];
                if (L2_TM) {
                    float zb[32];
#pragma unroll
                    for (int i = 0; i < 13; ++i) upk2(z2[i], zb[2 * i], zb[2 * i + 1]);
#pragma unroll
                    for (int i = 25; i < 32; ++i) zb[i] = 0.f;
#pragma unroll
                    for (int sft = 16; sft >= 1; sft >>= 1) {
                        const bool hi = (lane & sft) != 0;
#pragma unroll
                        for (int i = 0; i < sft; ++i) {
                            const float send = hi ? zb[i] : zb[i + sft], keep = hi ? zb[i + sft] : zb[i];
                            zb[i] = keep + __shfl_xor_sync(0xffffffffu, send, sft);
                        }
                    }
                    if (lane < 25) scratch[(warp & 3) * L2_NP + 25 * (warp >> 2) + lane] = zb[0];      // tile 0: four lane-quarter partials
                }
                compute_sync<L2_CT>();
                for (int q = tid; q < nta * L2_NP; q += L2_CT) {
                    const int i = q / L2_NP, pp = q - i * L2_NP;
                    float z = 0.f;
                    if (L2_TM && i == 0) {
                        z = (scratch[pp] + scratch[L2_NP + pp]) + (scratch[2 * L2_NP + pp] + scratch[3 * L2_NP + pp]);
                    } else {
#pragma unroll 4
                    for (int g2 = 0; g2 < 16; ++g2) z += scratch[i * SCS + g2 * L2_NP + pp];
                    }
                    const int py = pp / L2_TW, px = pp - py * L2_TW;
                    zt[i * ZTS + (py + 1) * L2_ZW + px + 1] = z;
                    st_tagged(&zll[(size_t)(gstep & 1u) * S * L2_HW + (size_t)tile_shot(i) * L2_HW + (tile_y0(i) + py) * L2_WL + tile_x0(i) + px],
                              z, gstep + 1u);
                }
                if (L2_HALO_GATE) asm volatile("bar.arrive 5, %0;" ::"n"(L2_CT + 32) : "memory");     // wake the halo warp (non-blocking)
                compute_sync<L2_CT>();
                // ------------ HR: interior cells of every tile, then (halo) the cells on the tile borders ------------
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    if (m == 1) {
                        unsigned it = 0;
                        while (!mbar_try_wait(halo_ready, gstep & 1u)) {
                            if ((++it & 0xfffu) == 0u) {
                                if (*reinterpret_cast<volatile unsigned*>(p.abort_flag) != 0u) break;
                                if (it > RES_SPIN_LIMIT) { atomicExch(p.abort_flag, 1u); break; }
                            }
                        }
                    }
#pragma unroll
                    for (int i = 0; i < NT; ++i) {
                        const unsigned d = hrd[(i * 2 + m) * L2_CT + tid];
                        if (d >> 31) {
                            const int ci = (d >> 19) & 0xff, r = tid & 7;
                            float gl = 0.f, gr = 0.f;
                            if (d & (1u << 18)) {
                                const float* z = zt + i * ZTS;
                                const int zi = ci + (int)((d >> 27) & 0xfu);     // cy * ZW + cx = ci + cy
                                const int dx = (d >> 16) & 1, dy = (d & (1u << 17)) ? L2_ZW : 0;
                                const float z00 = z[zi], z01 = z[zi + dx], z10 = z[zi + dy], z11 = z[zi + dy + dx];
                                const float fr = (float)r * 0.125f;
                                const float left = fmaf(fr, z10 - z00, z00), right = fmaf(fr, z11 - z01, z01);
                                if (hr_uniform & (1u << (2 * i + m))) hires_row_uniform(left, right, (d & 1u) != 0u, c01.x, c01.y, gl, gr);
                                else hires_row(left, right, d & 0xffffu, c01.x, c01.y, gl, gr);
                            }
                            *reinterpret_cast<float2*>(scratch + i * SCS + 2 * (ci * 8 + r)) = make_float2(gl, gr);
                        }
                    }
                }
                compute_sync<L2_CT>();
                // adjoint gather: g(q) = sum_r (1-r/8) [gl(q,r) + gr(q-1,r)] + (r/8) [gl(q-w,r) + gr(q-w-1,r)], 4 threads per pixel
#pragma unroll
                for (int i = 0; i < NT; ++i) {
                    if (i < nta) {
                        const int pl = tid >> 2, rq = tid & 3;
                        float s2 = 0.f;
                        if (pl < L2_NP) {
                            const int py = pl / L2_TW, px = pl - py * L2_TW;
                            const int ci = (py + 1) * L2_CW + px + 1;
                            const int a = tile_y0(i) + py, b = tile_x0(i) + px;
                            const float4* sc4 = reinterpret_cast<const float4*>(scratch + i * SCS) + rq;
                            const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
                            const float4 own = sc4[ci * 4];
                            const float4 lft = (b > 0) ? sc4[(ci - 1) * 4] : zero4;
                            const float4 up = (a > 0) ? sc4[(ci - L2_CW) * 4] : zero4;
                            const float4 ul = (a > 0 && b > 0) ? sc4[(ci - L2_CW - 1) * 4] : zero4;
                            const float ha = (float)(2 * rq) * 0.125f, hb = (float)(2 * rq + 1) * 0.125f;
                            s2 = fmaf(1.f - ha, own.x + lft.y, ha * (up.x + ul.y)) + fmaf(1.f - hb, own.z + lft.w, hb * (up.z + ul.w));
                        }
                        s2 += __shfl_xor_sync(0xffffffffu, s2, 1);
                        s2 += __shfl_xor_sync(0xffffffffu, s2, 2);
                        if (pl < L2_NP && rq == 0) gsm[i * GS + pl] = s2;
                    }
                }
                compute_sync<L2_CT>();
                // ------------ P3: quad j owns channel chunk j of every tile; one thread per channel ------------
                {
                    const int j = warp >> 2, cl = tid & 127;
                    float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
                    if (L2_TM) {
                        // tile 0 from tensor memory: thread = channel 128 j + cl = lane (warp % 4) * 32 + lane of its quarter, columns
                        // 100 j .. 100 j + 99; no ring unit involved, so all quads do it at once (same order of the sums as below)
                        f32x2 da = 0ull, db = 0ull;
                        const uint32_t tcol = tq + 100u * (uint32_t)j, gad = smem_u32(gsm);
                        tm_p3_block<32>(tcol, gad, da, db);
                        tm_p3_block<32>(tcol + 32u, gad + 128u, da, db);
                        tm_p3_block<32>(tcol + 64u, gad + 256u, da, db);
                        tm_p3_block<4>(tcol + 96u, gad + 384u, da, db);
                        upk2(da, d0, d1); upk2(db, d2, d3);
                    }
                    // The quads take their units in ring order. A parity wait is only meaningful for the current or the next
                    // phase of a stage's barrier, so quad j starts waiting for its first unit only when quad j - 1 has SEEN its
                    // last one (named barriers 2..4): by then every earlier phase of every stage is complete.
                    if (nta > I0 && j > 0) asm volatile("bar.sync %0, 256;" ::"r"(1 + j) : "memory");      // (no chain without streamed tiles)
#pragma unroll
                    for (int i = I0; i < NT; ++i) {
                        if (i < nta) {
                            const unsigned st = wait_full(uc + (unsigned)(j * (nta - I0) + (i - I0)));
                            if (i == nta - 1 && j < L2_KCH - 1) asm volatile("bar.arrive %0, 256;" ::"r"(2 + j) : "memory");
                            const uint32_t row = smem_u32(ring + (size_t)st * (L2_STAGE_BYTES / 4)) + (uint32_t)cl * 400u;
                            const uint32_t gad = smem_u32(gsm + i * GS);
                            float4 fb[4], gb[4];
#pragma unroll
                            for (int q = 0; q < 4; ++q) { fb[q] = lds128_v(row + 16 * q); gb[q] = lds128_v(gad + 16 * q); }
#pragma unroll
                            for (int q = 0; q < 25; ++q) {
                                const float4 f = fb[q % 4], g = gb[q % 4];
                                if (q + 4 < 25) { fb[q % 4] = lds128_v(row + 16 * (q + 4)); gb[q % 4] = lds128_v(gad + 16 * (q + 4)); }
                                d0 = fmaf(g.x, f.x, d0); d1 = fmaf(g.y, f.y, d1); d2 = fmaf(g.z, f.z, d2); d3 = fmaf(g.w, f.w, d3);
                            }
                            release(st, 4u);
                        }
                    }
                    uc += (unsigned)units_per_step_sweep;
                    const float d = (d0 + d1) + (d2 + d3);
                    unsigned long long* acc_t = acc_ep + (t & 1) * L2_C;
                    red_add_u64(acc_t + j * 128 + cl, ((unsigned long long)__float2ll_rn(d * fx_inv) << KB) + 1ull);
                }
                // the next sweep's units are more than a ring ahead of a quad that finished early: nobody may wait for them
                // before every unit of this sweep has been seen (same parity argument)
                compute_sync<L2_CT>();
            }
            if (p.T > 0) {
#pragma unroll
                for (int j = 0; j < L2_KCH; ++j) wait_applied(j, gstep - 1u);
            }
            l2_role_sync();                                           // S5
            if (k == 0) {
                const bool bad = *reinterpret_cast<volatile unsigned*>(p.abort_flag) != 0u || !ep_finite;
                for (int c = tid; c < L2_C; c += L2_CT) {
                    p.w[((size_t)e * 2) * L2_C + c] = bad ? __int_as_float(0x7fc00000) : W0[c];
                    p.w[((size_t)e * 2 + 1) * L2_C + c] = bad ? __int_as_float(0x7fc00000) : W1[c];
                }
            }
        }
        if (L2_TM && warp == 0) { tmem_fence_after_sync(); tmem_dealloc_512(tmem_slot); }      // every warp's last tensor-memory load precedes S5
    } else if (warp == L2_CT / 32) {
        // =====================================================================================================
        // applier warp: accumulator words -> SGD update in shared memory -> release the compute warps (fit_resident.cu)
        // =====================================================================================================
        const unsigned long long cnt_mask = (1ull << KB) - 1ull;
        for (int e = group; e < p.E; e += p.G) {
            l2_role_sync();                                           // S1
            for (int c = lane; c < 2 * L2_C; c += 32) cum[c] = 0ll;
            l2_role_sync();                                           // S2
            l2_role_sync();                                           // S3
            l2_role_sync();                                           // S4
            float fx_unit;
            { float inv; fixed_point_unit(smax[1], p.T, KB, fx_unit, inv); }
            const unsigned long long* acc_ep = p.sums + acc_offset(e, group, p.G, p.SPL, L2_C);
            for (int t = 0; t < p.T; ++t) {
                const unsigned long long expect = (unsigned long long)p.CPG * (unsigned)(t / 2 + 1);
                long long* cum_t = cum + (t & 1) * L2_C;
                for (int j = 0; j < L2_KCH; ++j) {
                    const unsigned long long* sw = acc_ep + (t & 1) * L2_C + j * L2_CCH;
                    unsigned long long wv[4];
                    unsigned pending = 0xfu, it = 0u;
#pragma unroll
                    for (int m = 0; m < 4; ++m) wv[m] = 0ull;
                    while (pending) {
#pragma unroll
                        for (int m = 0; m < 4; ++m)
                            if (pending & (1u << m)) wv[m] = ld_relaxed_u64(&sw[m * 32 + lane]);
#pragma unroll
                        for (int m = 0; m < 4; ++m)
                            if ((pending & (1u << m)) && (wv[m] & cnt_mask) == expect) pending &= ~(1u << m);
                        if (pending && (++it & 0xffu) == 0u) {
                            if (*reinterpret_cast<volatile unsigned*>(p.abort_flag) != 0u) break;
                            if (it > (RES_SPIN_LIMIT >> 2)) { atomicExch(p.abort_flag, 1u); break; }
                        }
                    }
#pragma unroll
                    for (int m = 0; m < 4; ++m) {
                        const int c = j * L2_CCH + m * 32 + lane;
                        const long long cur = (long long)wv[m] >> KB;
                        const float dw = __ll2float_rn(cur - cum_t[c]) * fx_unit;
                        cum_t[c] = cur;
                        const float n0 = fmaf(p.lr, dw, W0[c]), n1 = fmaf(-p.lr, dw, W1[c]);
                        W0[c] = n0; W1[c] = n1; Wd[c] = n1 - n0;
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&applied[j]);
                }
            }
            l2_role_sync();                                           // S5
        }
    } else if (warp == L2_CT / 32 + 1) {
        // =====================================================================================================
        // halo warp: the ring of z around every tile, step after step
        // =====================================================================================================
        int ring_zi[NT][2], ring_q[NT][2];
#pragma unroll
        for (int i = 0; i < NT; ++i)
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                const int r = lane + 32 * m;
                ring_zi[i][m] = -1; ring_q[i][m] = -1;
                if (i < nta && r < L2_NRING) {
                    int zy, zx;
                    if (r < L2_ZW) { zy = 0; zx = r; }
                    else if (r < 2 * L2_ZW) { zy = L2_ZH - 1; zx = r - L2_ZW; }
                    else if (r < 2 * L2_ZW + (L2_ZH - 2)) { zy = 1 + (r - 2 * L2_ZW); zx = 0; }
                    else { zy = 1 + (r - 2 * L2_ZW - (L2_ZH - 2)); zx = L2_ZW - 1; }
                    const int a = tile_y0(i) - 1 + zy, b = tile_x0(i) - 1 + zx;
                    ring_zi[i][m] = i * ZTS + zy * L2_ZW + zx;
                    if (a >= 0 && a < L2_HL && b >= 0 && b < L2_WL) ring_q[i][m] = tile_shot(i) * L2_HW + a * L2_WL + b;
                }
            }
        unsigned gstep = 0;
        for (int e = group; e < p.E; e += p.G) {
            l2_role_sync();                                           // S1
            l2_role_sync();                                           // S2
            l2_role_sync();                                           // S3
            l2_role_sync();                                           // S4
            for (int t = 0; t < p.T; ++t, ++gstep) {
                const unsigned long long* zsrc = zll + (size_t)(gstep & 1u) * S * L2_HW;
                unsigned bits[NT][2], tag[NT][2];
                unsigned pending = 0u, it = 0u;
                if (L2_HALO_GATE) asm volatile("bar.sync 5, %0;" ::"n"(L2_CT + 32) : "memory");       // this CTA's z of the step is out
#pragma unroll
                for (int i = 0; i < NT; ++i)
#pragma unroll
                    for (int m = 0; m < 2; ++m) { bits[i][m] = 0u; tag[i][m] = 0u; if (ring_q[i][m] >= 0) pending |= 1u << (2 * i + m); }
                while (pending) {
#pragma unroll
                    for (int i = 0; i < NT; ++i)
#pragma unroll
                        for (int m = 0; m < 2; ++m)
                            if (pending & (1u << (2 * i + m))) ld_tagged(&zsrc[ring_q[i][m]], bits[i][m], tag[i][m]);
#pragma unroll
                    for (int i = 0; i < NT; ++i)
#pragma unroll
                        for (int m = 0; m < 2; ++m)
                            if ((pending & (1u << (2 * i + m))) && tag[i][m] == gstep + 1u) pending &= ~(1u << (2 * i + m));
                    if (pending && (++it & 0xffu) == 0u) {
                        if (*reinterpret_cast<volatile unsigned*>(p.abort_flag) != 0u) break;
                        if (it > (RES_SPIN_LIMIT >> 2)) { atomicExch(p.abort_flag, 1u); break; }
                    }
                }
#pragma unroll
                for (int i = 0; i < NT; ++i)
#pragma unroll
                    for (int m = 0; m < 2; ++m)
                        if (ring_zi[i][m] >= 0) zt[ring_zi[i][m]] = (ring_q[i][m] >= 0) ? __uint_as_float(bits[i][m]) : 0.f;
                // pace this warp on the CTA's own z (a tile whose whole ring lies outside the image has nothing to poll): it may
                // never run a whole mbarrier phase ahead of the compute warps
                if (lane == 0 && nta > 0)
                    (void)poll_word(&zsrc[(size_t)tile_shot(0) * L2_HW + tile_y0(0) * L2_WL + tile_x0(0)], gstep + 1u, p.abort_flag);
                __syncwarp();
                if (lane == 0) mbar_arrive(halo_ready);
            }
            l2_role_sync();                                           // S5
        }
    } else {
        // =====================================================================================================
        // producer warp: one TMA tile copy per (tile, channel chunk) unit, in consumption order
        // =====================================================================================================
        // (free-running over all episodes: the ring's full / empty barriers are its only synchronisation — the features are
        // read-only, and nothing else ever writes the ring)
        if (lane == 0) {
            unsigned uc = 0;
            unsigned long long pol_first, pol_last;                // L2 eviction priorities of the two kinds of loads
            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));
            asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_last));
            unsigned ep = 0;
            for (int e = group; e < p.E; e += p.G, ++ep) {
                const int n_sweeps = 1 + 2 * p.T;
                for (int sw = 0; sw < n_sweeps; ++sw) {
                    if (sw == 1) {                       // the linear copy of this episode's tiles is complete
                        unsigned it = 0;
                        while (!mbar_try_wait(retiled, ep & 1u)) {
                            if ((++it & 0xfffu) == 0u) {
                                if (*reinterpret_cast<volatile unsigned*>(p.abort_flag) != 0u) return;
                                if (it > RES_SPIN_LIMIT) { atomicExch(p.abort_flag, 1u); return; }
                            }
                        }
                    }
                    for (int j = 0; j < L2_KCH; ++j)
                        for (int i = (sw == 0 ? 0 : I0); i < nta; ++i, ++uc) {
                            const unsigned st = uc % L2_NSTAGE, ph = (uc / L2_NSTAGE) & 1u;
                            unsigned it = 0;
                            while (!mbar_try_wait(&empty[st], ph ^ 1u)) {       // first pass over the ring: returns at once
                                if ((++it & 0xfffu) == 0u) {
                                    if (*reinterpret_cast<volatile unsigned*>(p.abort_flag) != 0u) return;
                                    if (it > RES_SPIN_LIMIT) { atomicExch(p.abort_flag, 1u); return; }
                                }
                            }
                            mbar_expect_tx(&full[st], L2_STAGE_BYTES);
                            if (sw == 0) {
                                // read once: evict-first, so that the original maps do not push the re-laid-out copies out of L2
                                const int z = ((e * S + tile_shot(i)) * L2_C) + j * L2_CCH;
                                asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;"
                                             ::"r"(smem_u32(ring + (size_t)st * (L2_STAGE_BYTES / 4))), "l"(&fmap), "r"(smem_u32(&full[st])),
                                               "r"(tile_x0(i)), "r"(tile_y0(i)), "r"(z), "l"(pol_first) : "memory");
                            } else {
                                const float* src = p.ftile + (((size_t)group * p.NTILE + (k + i * p.CPG)) * L2_KCH + j) * (L2_STAGE_BYTES / 4);
                                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                                             ::"r"(smem_u32(ring + (size_t)st * (L2_STAGE_BYTES / 4))), "l"(src), "r"(L2_STAGE_BYTES),
                                               "r"(smem_u32(&full[st])), "l"(pol_last) : "memory");
                            }
                        }
                }
            }
        }
    }
}

// ---- host side ---------------------------------------------------------------------------------
struct L2Plan { int NT, CPG, G; bool ok; };

// NT tiles per CTA, CPG CTAs per episode, G episodes at a time: as many concurrent episodes as the L2 budget and the SMs
// allow, then the fewest tiles per CTA
static L2Plan plan_l2(int E, int S, int n_sm) {
    L2Plan best{0, 0, 0, false};
    const int ntile = L2_TPI * S;
    double budget_mb = 100.0;                                   // of the 126 MB L2 (two 5-shot episodes = 73.7 MB)
    if (const char* s = getenv("CWT_FIT_L2_MB")) budget_mb = atof(s);
    const double ep_mb = (double)S * L2_C * L2_HW * 4 / 1e6;
    double best_score = -1.0;
    // Score = episodes in flight / time of one SGD step of a group. Step times measured on B200 with the first tile in tensor
    // memory (profiles/r2g_fit_l2_tmem_plans.txt, us): 6.2 / 8.9 / 13.0 / 17.5 for 1 / 2 / 3 / 4 tiles per CTA, + 0.019 per CTA of a
    // group beyond 72 (exchange latencies), + 0.25 per group beyond 4. (The round-2f score — episodes per tile slot — tied all
    // plans that fill the SMs and took the first: 1 615 instead of 2 255 episodes/s at 2 shots, 663 instead of 1 171 at 4.)
    static const double step_us[L2_MAXNT + 1] = {0.0, 6.2, 8.9, 13.0, 17.5};
    for (int NT = 1; NT <= L2_MAXNT; ++NT) {
        const int CPG = (ntile + NT - 1) / NT;
        if (CPG > n_sm) continue;
        int G = n_sm / CPG;
        // what has to stay in L2 for the steps: with the first tile of every CTA in tensor memory only (NT - 1) / NT of an
        // episode is ever re-read (its tile-major copy is not even written). 5 shots: 4 tiles per CTA x 3 groups = 83 MB
        // resident, 858 episodes/s, against 3 tiles x 2 groups = 49 MB, 807 (profiles/r2i_fit_l2_5shot_plans.txt)
        const double resident_mb = ep_mb * (L2_TM ? (double)(NT - 1) / NT : 1.0);
        const int g_l2 = resident_mb > 0.0 ? (int)(budget_mb / resident_mb) : G;
        if (G > g_l2) G = g_l2;
        if (G < 1) G = 1;                                        // a single episode larger than the budget still runs (from HBM)
        const double t_us = step_us[NT] + 0.019 * (CPG > 72 ? CPG - 72 : 0) + 0.25 * (G > 4 ? G - 4 : 0);
        const double score = (double)G / t_us * ((double)ntile / (NT * CPG));
        if (G > E) G = E;
        if (score > best_score + 1e-9) { best_score = score; best = L2Plan{NT, CPG, G, true}; }
    }
    if (const char* s = getenv("CWT_FIT_L2_NT")) {
        const int NT = atoi(s);
        if (NT >= 1 && NT <= L2_MAXNT && (ntile + NT - 1) / NT <= n_sm) {
            const int CPG = (ntile + NT - 1) / NT;
            int G = n_sm / CPG;
            const double resident_mb = ep_mb * (L2_TM ? (double)(NT - 1) / NT : 1.0);
            const int g_l2 = resident_mb > 0.0 ? (int)(budget_mb / resident_mb) : G;
            if (G > g_l2) G = g_l2 < 1 ? 1 : g_l2;
            if (G > E) G = E;
            best = L2Plan{NT, CPG, G, true};
        }
    }
    return best;
}

bool fit_l2_supported(int S, int C, int h, int w) { return C == L2_C && h == L2_HL && w == L2_WL && S >= 1 && L2_TPI * S <= 148 * L2_MAXNT; }

size_t fit_l2_workspace_bytes(int E, int S, int C, int h, int w) {
    if (!fit_l2_supported(S, C, h, w)) return 0;
    const L2Plan pl = plan_l2(E, S, 148);
    if (!pl.ok) return 0;
    const int nslots = (E + pl.G - 1) / pl.G;
    return align_up(sizeof(unsigned long long) * 2 * S * L2_HW * pl.G) + align_up(sizeof(unsigned long long) * (size_t)pl.G * nslots * 2 * L2_C) +
           align_up(sizeof(unsigned long long) * 320) + 512 + align_up((size_t)pl.G * S * L2_C * L2_HW * 4);
}

template <int NT>
static int launch_l2(const CUtensorMap& map, const L2Params& p, const L2Plan& pl, cudaStream_t st) {
    const size_t smem = l2_smem_layout(NT).total;
    const void* fn = (const void*)k_fit_l2<NT>;
    CWT_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int resident_ctas = 0;
    CWT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident_ctas, fn, L2_THREADS, smem));
    CWT_REQUIRE(resident_ctas >= 1, CWT_ERR_UNSUPPORTED, "fit_l2: the kernel does not fit on an SM (%zu B of shared memory)", smem);
    void* args[] = {const_cast<CUtensorMap*>(&map), const_cast<L2Params*>(&p)};
    CWT_CUDA(cudaLaunchCooperativeKernel(fn, dim3(pl.G * pl.CPG), dim3(L2_THREADS), args, smem, st));
    count_launch();
    return CWT_OK;
}

// returns CWT_ERR_UNSUPPORTED when the shape is not the head geometry
int fit_l2(const float* f_s, const uint4* cells, const float2* cw, float* w_io, int E, int S, int C, int h, int w,
           int n_iter, float lr, void* ws, size_t ws_bytes, cudaStream_t st) {
    CWT_REQUIRE(fit_l2_supported(S, C, h, w), CWT_ERR_UNSUPPORTED, "fit_l2: C=%d %dx%d S=%d is not the head geometry (512, 60x60)", C, h, w, S);
    int dev = 0, n_sm = 0, coop = 0;
    CWT_CUDA(cudaGetDevice(&dev));
    CWT_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    CWT_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    CWT_REQUIRE(coop, CWT_ERR_UNSUPPORTED, "fit_l2: device lacks cooperative launch");
    if (n_sm > 148) n_sm = 148;
    const L2Plan pl = plan_l2(E, S, n_sm);
    CWT_REQUIRE(pl.ok, CWT_ERR_UNSUPPORTED, "fit_l2: no plan for S=%d on %d SMs", S, n_sm);
    Carver cv(ws, ws_bytes);
    L2Params p{};
    p.zll = cv.take<unsigned long long>((size_t)2 * S * L2_HW * pl.G);
    p.SPL = (E + pl.G - 1) / pl.G;
    p.sums = cv.take<unsigned long long>((size_t)pl.G * p.SPL * 2 * L2_C);
    p.fmaxw = cv.take<unsigned long long>(320);
    p.abort_flag = cv.take<unsigned>(64);
    p.ftile = cv.take<float>((size_t)pl.G * S * L2_C * L2_HW);
    CWT_REQUIRE(ws && cv.ok(), CWT_ERR_WORKSPACE, "fit_l2: workspace too small");
    const size_t sync_bytes = (size_t)(reinterpret_cast<char*>(p.abort_flag + 64) - reinterpret_cast<char*>(p.zll));
    CWT_CUDA(cudaMemsetAsync(p.zll, 0, sync_bytes, st));
    p.cells = cells; p.cw = cw; p.w = w_io;
    p.E = E; p.S = S; p.CPG = pl.CPG; p.G = pl.G; p.T = n_iter; p.lr = lr; p.NTILE = L2_TPI * S;
    p.KBITS = 1;
    while ((1ll << p.KBITS) <= (long long)pl.CPG * n_iter) ++p.KBITS;
    CWT_REQUIRE(p.KBITS <= 24, CWT_ERR_UNSUPPORTED, "fit_l2: n_iter=%d too large for the on-chip all-reduce", n_iter);
    // tensor map over f_s as [E*S*C][60][60] fp32, box 20 x 5 x 128 channels
    LsEncodeFn enc = ls_encode_fn();
    CWT_REQUIRE(enc, CWT_ERR_CUDA, "fit_l2: cuTensorMapEncodeTiled is not available from this driver");
    CUtensorMap map;
    cuuint64_t dims[3] = {(cuuint64_t)L2_WL, (cuuint64_t)L2_HL, (cuuint64_t)E * S * L2_C};
    cuuint64_t strides[2] = {(cuuint64_t)L2_WL * 4, (cuuint64_t)L2_HW * 4};
    cuuint32_t box[3] = {L2_TW, L2_TH, L2_CCH};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(f_s), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CWT_REQUIRE(r == CUDA_SUCCESS, CWT_ERR_CUDA, "fit_l2: cuTensorMapEncodeTiled failed (%d)", (int)r);
    switch (pl.NT) {
        case 1: return launch_l2<1>(map, p, pl, st);
        case 2: return launch_l2<2>(map, p, pl, st);
        case 3: return launch_l2<3>(map, p, pl, st);
        default: return launch_l2<4>(map, p, pl, st);
    }
}

}  // namespace cwt
