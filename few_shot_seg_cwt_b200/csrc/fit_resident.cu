// CWT_FIT_RESIDENT — the support-classifier fit with the feature map kept ON CHIP for all n_iter steps.
//
// The streaming fit re-reads the [C, h*w] feature map from HBM twice per SGD step ((2T+1) F bytes per
// episode). Here one episode is spread over a GROUP of CPG = h*w / NP co-resident CTAs (one per SM);
// CTA k stages a rectangular TILE of TW x TH low-res pixels x all C channels (C*NP*4 B, 204.8 KB for
// 512 x (20 x 5)) into shared memory ONCE with bulk-TMA copies (cp.async.bulk + mbarrier complete_tx)
// and then runs every step out of shared memory:
//
//   P1  z[p]   = sum_c Wd[c] F[c][p]          linear conflict-free sweep of the tile, own NP pixels
//   X1  halo exchange of z: the ring of pixels around the tile (2 TW + 2 TH + 4 values), fetched by the
//       HALO WARP while the compute warps already work on the interior cells
//   HR  full-resolution stage on the (TW+1) x (TH+1) cells that touch the tile (one task = one row of a
//       cell; 126 cells = 1008 tasks = two full rounds of the 512 compute threads for 20 x 5)
//   P3  dW[c]  = sum_p g[p] F[c][p]           second sweep, one thread per channel (no cross-thread reduction);
//       the four warp quads take turns, so channel chunk j leaves at (j+1)/4 of the sweep
//   AR  group all-reduce of dW, SGD update of W0 / W1 / Wd
//
// The halo exchange uses SEQUENCE-TAGGED 8-byte words {fp32 value, step number}: 64-bit stores are
// single-copy atomic, so a reader simply polls the data word until its tag matches — no fences, no
// barriers, and a stale value can never be consumed.
//
// The all-reduce is ONE hop through the L2 atomic units, pipelined in RES_KCH channel chunks underneath
// P3 / P1: every channel of every episode has one 64-bit accumulator word {fixed-point sum of dW : V bits,
// arrival count : K bits}. A CTA adds its partial with a single fire-and-forget `red.add.u64` of
// (fix(d) << K) + 1 — value and arrival land in the same atomic, so no fence is needed, and integer
// addition makes the sum exact, order-independent and therefore deterministic. Words are cumulative (never
// reset) and there are two sets, for the even and the odd steps: a CTA that is not a neighbour of this one
// may run a whole step ahead, but it cannot add to step t+2 before every CTA has contributed to step t+1,
// i.e. has consumed step t — so the count of a set can never run past the value a reader is waiting for.
// The APPLIER WARP polls the chunk's words until the count field reads CPG*(t/2+1), takes the difference to
// the previous cumulative value of that set (exact), applies the SGD update to W0 / W1 / Wd in shared
// memory and releases the compute warps chunk by chunk through a shared-memory mbarrier. The compute warps
// never wait on global memory.
// The fixed-point unit is a power of two chosen per episode from max|F| (|dW_c| <= max|F| because the
// gradient w.r.t. the logits has l1 norm <= 1 under the weighted-mean CE), exchanged once per episode with
// tagged words; a partial loses nothing unless it is < 2^-26 of that bound.
//
// HBM traffic per episode drops from (2T+1) F to F (+ labels); the per-step bound becomes the shared
// memory sweep (2 x C*NP*4 B at 128 B/clk/SM). floor(#SM / CPG) groups run concurrently
// (4 x 36 = 144 of 148 SMs for 60x60x512), each looping over its share of the batch.
//
// TENSOR-MEMORY variant (default for the PSPNet head geometry, C = 512, 60 x 60, tile 20 x 5; round 2): shared memory only
// STAGES the tile; for the 200 steps it lives in the SM's 256 KB of tensor memory (tcgen05.alloc / st / ld, no MMA involved):
// lane = channel mod 128, column = 100 (channel / 128) + pi(pixel), 400 of the 512 columns (pi: tm_col2pix below — the 44 pixels
// that the first round of the full-resolution stage completes come first). tcgen05.ld streams ~380 B/clk/SM with 16
// warps (tools/micro/tmem_bench.cu) against the 128 B/clk of the shared-memory pipe, and it leaves the LSU to the exchanges:
//   P1  thread = lane L of its quarter, pixel block warp / 4 (25 pixels): in-thread FFMA2 over the 4 channel chunks, butterfly
//       reduce-scatter over the 32 lanes (31 shuffles), the four lane quarters combined through shared memory;
//   P3  thread = channel: its columns in blocks of 32 / 16 / 8 / 4, g broadcast from shared memory, no cross-thread reduction,
//       all chunks at once; the gather and the sweep over the first 44 columns run BEFORE the compute warps wait for the ring
//       (in the shadow of the halo exchange), the other 56 after the second round of the full-resolution stage;
//   HR  a per-episode coefficient table (it reuses the staging buffer) makes every live cell row cost the same 16 FMAs;
//   AR  no applier warp: after its RED every compute thread waits RES_POLL_DELAY clocks, polls the accumulator word of ITS
//       channel (one 8-byte load per lane; one warp polling all 512 words needs ~1 600 clk per round), applies the SGD step to
//       the weights it keeps in registers and publishes Wd; the halo warp sleeps on a named barrier until z is published.
//   staging: four tensor-map copies per tile (cp.async.bulk.tensor.3d, box 20 x 5 pixels x 128 channels of f_s seen as
//       [E C][60][60]) instead of one 80-byte bulk copy per channel and tile row.
// Measured (E = 64, 200 steps): 11.28 ms = 6 820 clk per step + 17 us per episode against 10 400 clk for the shared-memory kernel
// (DESIGN.md 4.4 lists the step-by-step measurements and the variants that were slower).
//
// The shared-memory kernel is compiled for 512 compute threads / one CTA per SM and for 256 / two CTAs per SM
// (CWT_RESIDENT_BPS=2: measured slower, kept for comparison), each specialised for the PSPNet head geometry
// (C=512, 60x60, tile 20x5 resp. 4x10: all loop bounds, divisions and predicates become compile-time) and
// generic (run-time shapes); CWT_RESIDENT_TMEM=0 selects it for the head geometry too.
//
// Launch: cooperative (all CTAs must be co-resident: they poll each other's words). Every poll loop
// has a watchdog: on timeout the kernel raises an abort flag, stops waiting and writes NaN results
// rather than hanging the GPU.
#include "common.cuh"
#include "hires.cuh"
#include "resident_common.cuh"
#include "tma_pipe.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>

// Developer builds only (tools/build_variants.py): bit mask of ABLATIONS that are timed against the product kernel
// (RES_VARIANT == 0; they give wrong results on purpose: they measure what a phase costs). 0x01 compute warps do not wait for
// the halo, 0x02 no full-resolution arithmetic, 0x04 compute warps do not wait for the all-reduce, 0x08 no P3 sweep, 0x10 no P1
// sweep, 0x20 shared-memory kernel: P3 with all quads at once, 0x100 / 0x200 the applier / halo warps do not poll.
// (The alternative implementations measured in round 2 — P3 with lane = pixel quad, per-chunk applier warps, bulk-copy polls,
// Wd in consumption order: profiles/r2_resident_ablation_*.txt — lost and were removed from the source.)
#ifndef RES_VARIANT
#define RES_VARIANT 0
#endif

namespace cwt {

// A CTA has CT compute threads + the applier warp + the halo warp. Two configurations are built:
//   CT = 512, one CTA per SM   (tile 20 x 5 for the PSPNet head: the whole shared memory holds one tile)
//   CT = 256, two CTAs per SM  (tile 4 x 10: while one CTA waits for the halo / the all-reduce or runs the ALU-bound
//                               full-resolution stage, the other one could keep the shared-memory pipe busy — measured
//                               25.4 ms vs 17 ms at E = 64: the exchange latencies grow with the 90 CTAs per episode)
// helper warps: the applier warp (shared-memory kernels) + the halo warp.
// Tensor-memory kernel, RES_SPLIT: the step's work that does not need the halo — the full-resolution stage of the first 64
// interior cells, the gather of the 44 pixels those cells complete and the P3 sweep over them — runs BEFORE the compute warps
// wait for the ring, i.e. in the shadow of the halo exchange. For that the 100 pixel columns of a channel chunk are kept in
// this order: the 45 pixels (rows 1..3, x = 1..15) whose four cells are the cells cy = 1..4, cx = 1..16 first, the others after.
#ifndef RES_SPLIT
#define RES_SPLIT 1
#endif
constexpr int TM_EARLY = 44;            // columns of the early P3 part (a multiple of 4: g is read in quads)
__host__ __device__ constexpr int tm_col2pix(int c) {
    if (!RES_SPLIT) return c;
    if (c < 45) return 20 * (1 + c / 15) + 1 + c % 15;
    if (c < 65) return c - 45;
    if (c < 80) return 20 * (1 + (c - 65) / 5) + (((c - 65) % 5) ? 15 + (c - 65) % 5 : 0);
    return c;
}
__host__ __device__ constexpr int tm_pix2col(int p) {
    if (!RES_SPLIT) return p;
    const int py = p / 20, px = p % 20;
    if (py == 0) return 45 + px;
    if (py == 4) return 80 + px;
    if (px >= 1 && px <= 15) return 15 * (py - 1) + (px - 1);
    return 65 + 5 * (py - 1) + (px ? px - 15 : 0);
}
// compile-time checks of the column order: a permutation, and the early columns are exactly pixels whose four cells belong to
// HR round 0 (cells cy = 1..4, cx = 1..16 <=> pixel rows 1..3, x = 1..15)
constexpr bool tm_order_ok() {
    bool seen[100] = {};
    for (int c = 0; c < 100; ++c) {
        const int p = tm_col2pix(c);
        if (p < 0 || p >= 100 || seen[p] || tm_pix2col(p) != c) return false;
        seen[p] = true;
        const int py = p / 20, px = p % 20;
        if (RES_SPLIT && c < TM_EARLY && !(py >= 1 && py <= 3 && px >= 1 && px <= 15)) return false;
    }
    return true;
}
static_assert(tm_order_ok(), "tm_col2pix / tm_pix2col must be inverse permutations with the round-0 pixels first");
constexpr int res_helper_threads(int NA) { return 32 * (NA + 1); }     // NA applier warps + the halo warp
constexpr int RES_KCH = 4;        // channel chunks of the pipelined all-reduce
constexpr int RES_AW = 4;         // accumulator words an applier lane keeps in flight
constexpr int RES_HWORDS = 4;     // halo words per lane of the halo warp (ring <= 128 pixels)
constexpr int RES_MAXTASK = 4;    // HR row-tasks per compute thread (8 * cells <= 4 * 512)
constexpr int RES_NPROF = 12;     // profile counters per CTA
constexpr int RES_P3_DEPTH = 4;   // pixel-quads a P3 thread keeps in flight per operand
// polling of the accumulator words by the compute threads (tensor-memory kernel): clocks between the RED and the first load
#ifndef RES_POLL_DELAY
#define RES_POLL_DELAY 600
#endif
#ifndef RES_HALO_GATE
#define RES_HALO_GATE 1            // the halo warp sleeps until its CTA has published z (tensor-memory kernel)
#endif
// accumulator words of the all-reduce are 8 << RES_ACC_SHIFT bytes apart (0: 16 words per 128-byte line, 2: one word per 32-byte
// sector, 4: one word per line). Measured in round 2 (profiles/r2i_acc_stride.txt): see DESIGN.md 4.4.
#ifndef RES_ACC_SHIFT
#define RES_ACC_SHIFT 0
#endif
#define RES_ACC(i) ((size_t)(i) << RES_ACC_SHIFT)
// Tensor-memory kernel: the tile is staged with 4 tensor-map copies (box 20 x 5 pixels x 128 channels of f_s seen as
// [E C][60][60]; SASS UTMALDG.3D) instead of 2 560 bulk copies of 80 bytes (one per channel and tile row), whose issue rate
// bounded the staging of an episode.
#ifndef RES_TMA_STAGE
#define RES_TMA_STAGE 1
#endif

struct ResidentParams {
    const float* f_s;          // [E][C][HW]
    const uint4* cells;        // [E][HW]
    const float2* cw;          // [E]
    float* w;                  // [E][2][C]  in: W0, out: fitted
    unsigned long long* zll;   // [G][2][HW]                            {z, step} words of the halo exchange
    unsigned long long* sums;  // [G][SPL][2][C]                        {fixed-point cumulative dW : 64-K bits, arrivals : K bits},
                               //   one set for the even and one for the odd steps of every episode
    unsigned long long* fmaxw; // [G][CPG]                              {max|F| of the tile (bits), episode+1} words
    unsigned* abort_flag;      // [1]
    long long* prof;           // [grid][RES_NPROF] or null
    int E, C, HW, h, w_lo, TW, TH, CPG, G, T, KBITS, SPL;
    float lr;
};

struct ResSmem {
    size_t F, W0, W1, Wd, zt, g, scratch, mbar, cum, poll, total;
};
static __host__ __device__ inline ResSmem res_smem_layout(int C, int TW, int TH, int CT) {
    ResSmem s;
    const int NP = TW * TH, NQ = NP / 4, NG = CT / NQ, NCELL = (TW + 1) * (TH + 1);
    size_t o = 0;
    size_t sz[10];
    sz[0] = (size_t)C * NP * 4; sz[1] = sz[2] = sz[3] = (size_t)C * 4;
    sz[4] = (size_t)(TW + 2) * (TH + 2) * 4; sz[5] = (size_t)NP * 4;
    const size_t sc1 = (size_t)NG * NP * 4, sc2 = (size_t)16 * NCELL * 4;
    sz[6] = sc1 > sc2 ? sc1 : sc2;
    sz[7] = 8 * (2 + RES_KCH) + 16 + 16;                             // mbarriers + two words of the max|F| exchange + poll mbarrier
    sz[8] = (size_t)C * 16;                                          // previous cumulative dW per channel and step parity (applier warp)
    sz[9] = 0;
    size_t off[10];
    for (int i = 0; i < 10; ++i) { off[i] = o; o = (o + sz[i] + 127) / 128 * 128; }
    s.F = off[0]; s.W0 = off[1]; s.W1 = off[2]; s.Wd = off[3]; s.zt = off[4]; s.g = off[5];
    s.scratch = off[6]; s.mbar = off[7]; s.cum = off[8]; s.poll = off[9]; s.total = o;
    return s;
}

// CT: compute threads, MINB: CTAs per SM. TC / TTW / TTH / TWL / THL > 0: compile-time shape (C, tile width / height,
// low-res width / height); 0: run-time shape. TM: the tile lives in TENSOR MEMORY for the steps (PSPNet head geometry, one CTA
// per SM): lane = channel mod 128, column = 100 (channel / 128) + pixel — shared memory only stages it. NA: applier warps
// (1 for the shared-memory kernels; 0 for the tensor-memory kernel, whose compute threads poll and apply their own channel).
template <int CT, int MINB, int TC, int TTW, int TTH, int TWL, int THL, bool PROF, bool TM, int NA>
__global__ void __launch_bounds__(CT + res_helper_threads(NA), MINB)
k_fit_resident(const __grid_constant__ CUtensorMap fmap, ResidentParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int C = TC ? TC : p.C, TW = TTW ? TTW : p.TW, TH = TTH ? TTH : p.TH, wl = TWL ? TWL : p.w_lo, h = THL ? THL : p.h;
    const int NP = TW * TH, HW = wl * h, TPR = wl / TW, CPG = HW / NP;
    const int CCH = C / RES_KCH;                     // channels per chunk (C % RES_KCH == 0)
    constexpr bool kStatic = (CT == 512 && TC == 512 && TTW * TTH == 100);   // PSPNet head geometry, one CTA per SM: fully unrolled sweeps
    constexpr bool kStatic40 = (CT == 256 && TC == 512 && TTW * TTH == 40);  // ... two CTAs per SM
    static_assert(!TM || kStatic, "the tensor-memory tile is built for C = 512, 100 pixels, 512 compute threads");
    const int NQ = NP / 4, NG = kStatic ? 16 : CT / NQ;
    const int ZW = TW + 2, ZH = TH + 2, CW = TW + 1, NCELL = CW * (TH + 1);
    const int NINT = (TW - 1) * (TH - 1);            // cells whose four corners are own pixels: need no halo
    const bool round0_interior = 8 * NINT >= CT;       // HR round 0 can run before the halo has arrived
    const int NRING = 2 * ZW + 2 * (ZH - 2);         // halo pixels around the tile
    const ResSmem L = res_smem_layout(C, TW, TH, CT);
    float* F = reinterpret_cast<float*>(smem_raw + L.F);
    float* W0 = reinterpret_cast<float*>(smem_raw + L.W0);
    float* W1 = reinterpret_cast<float*>(smem_raw + L.W1);
    float* Wd = reinterpret_cast<float*>(smem_raw + L.Wd);
    float* zt = reinterpret_cast<float*>(smem_raw + L.zt);       // [ZH][ZW]: the tile's z plus the ring around it
    float* gsm = reinterpret_cast<float*>(smem_raw + L.g);
    float* scratch = reinterpret_cast<float*>(smem_raw + L.scratch);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw + L.mbar);
    uint64_t* applied = mbar + 1;                     // [RES_KCH] chunk j of the previous step has been applied to Wd
    uint64_t* halo_ready = mbar + 1 + RES_KCH;        // the ring of this step is in zt
    unsigned* smax = reinterpret_cast<unsigned*>(mbar + 2 + RES_KCH);   // [0] tile max|F| bits, [1] episode max|F| bits
    long long* cum = reinterpret_cast<long long*>(smem_raw + L.cum);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int group = blockIdx.x / CPG, k = blockIdx.x - group * CPG;
    const int ty = k / TPR, tx = k - ty * TPR, y0 = ty * TH, x0 = tx * TW;     // tile origin in the low-res map

    // group-private exchange areas (halo words are {value, step}; step numbers start at 1)
    unsigned long long* zll = p.zll + (size_t)group * 2 * HW;
    unsigned long long* fmaxw = p.fmaxw + (size_t)group * CPG;
    const int KB = p.KBITS;                                            // arrival-count bits of an accumulator word

    __shared__ uint32_t tmem_slot;
    if (tid == 0) {
        mbar_init(mbar, 1);
        for (int j = 0; j < RES_KCH; ++j) mbar_init(&applied[j], 1);
        mbar_init(halo_ready, 1);
    }
    if (TM && warp == 0) tmem_alloc_512(&tmem_slot);       // all 512 columns: this CTA has the SM to itself
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (TM) tmem_fence_before_sync();
    __syncthreads();
    if (TM) tmem_fence_after_sync();
    // this warp's window of tensor memory: lane quarter 32 (warp % 4), bits 31..16 of a TMEM address are the lane
    const uint32_t tq = TM ? tmem_slot + ((uint32_t)((warp & 3) * 32) << 16) : 0u;

    __shared__ long long tstore_sm[RES_KCH];          // PROF: when chunk j's partials left this CTA

    // Every role has its own episode loop (so that the state of one role never occupies registers of another); all of
    // them execute the same five CTA-wide barriers per episode: S1 previous episode done, S2 tile staged, S3 tile max
    // known, S4 episode max known, S5 steps done.
    if (warp < CT / 32) {
        // =====================================================================================================
        // compute warps
        // =====================================================================================================
        unsigned gstep = 0, tma_parity = 0;
        bool ok = true;
        long long t_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const bool p1_active = tid < NQ * NG;
        const int v = tid % NQ, grp = tid / NQ;

        // chunk j of global step gs has been folded into W0 / W1 / Wd by the applier warp
        auto wait_applied = [&](int j, unsigned gs) {
            if (RES_VARIANT & 0x04) return;
            unsigned it = 0;
            if (PROF && tid == 0 && !mbar_try_wait(&applied[j], gs & 1u)) {      // had to wait: latency = now - store time
                while (!mbar_try_wait(&applied[j], gs & 1u)) { if (++it > RES_SPIN_LIMIT) break; }
                t_acc[5] += clock64() - *reinterpret_cast<volatile long long*>(&tstore_sm[j]);
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
            // ---------------- stage the episode: tile of F via bulk-TMA, weights, HR task descriptors ----------------
            __syncthreads();                                           // S1: previous episode is done with shared memory
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy reads before async-proxy writes
            if (tid == 0) mbar_expect_tx(mbar, (unsigned)(C * NP * 4));
            if constexpr (TM && RES_TMA_STAGE != 0) {
                // four tensor-map copies of 128 channels x 5 rows x 20 pixels: the box arrives densely packed, i.e. as F[c][py][px]
                if (tid == 0) {
#pragma unroll
                    for (int j = 0; j < RES_KCH; ++j)
                        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                                     ::"r"(smem_u32(F + (size_t)j * 128 * NP)), "l"(&fmap), "r"(smem_u32(mbar)),
                                       "r"(x0), "r"(y0), "r"(e * C + j * 128) : "memory");
                }
            } else {
                const float* fsrc = p.f_s + (size_t)e * C * HW + (size_t)y0 * wl + x0;
                for (int i = tid; i < C * TH; i += CT) {         // one copy per (channel, tile row): TW * 4 bytes
                    const int c = i / TH, py = i - c * TH;
                    bulk_g2s(F + (size_t)c * NP + py * TW, fsrc + (size_t)c * HW + (size_t)py * wl, (unsigned)(TW * 4), mbar);
                }
            }
            for (int c = tid; c < C; c += CT) {
                const float a = p.w[((size_t)e * 2) * C + c], b = p.w[((size_t)e * 2 + 1) * C + c];
                W0[c] = a; W1[c] = b; Wd[c] = b - a;
            }
            if (tid < 2) smax[tid] = 0u;
            // HR task descriptors (static for the episode): task = (cell, row r); 8 adjacent lanes = one cell. Interior cells
            // come first so that the first round needs no halo.
            // desc = row label bits | dx << 16 | dy_flag << 17 | live << 18 | cell << 19 | cell_row << 27 | valid << 31
            unsigned hr_desc[RES_MAXTASK];
            unsigned hr_uniform = 0u;
#pragma unroll
            for (int m = 0; m < RES_MAXTASK; ++m) {
                const int task = m * CT + tid;
                const int o = task >> 3, r = task & 7;
                unsigned d = 0u;
                if (o < NCELL) {
                    int cy, cx;
                    if (TM && RES_SPLIT && o < 64) { cy = 1 + o / 16; cx = 1 + o % 16; }             // the cells of the early part first
                    else if (TM && RES_SPLIT && o < NINT) { cy = 1 + (o - 64) / 3; cx = 17 + (o - 64) % 3; }
                    else if (o < NINT) { cy = 1 + o / (TW - 1); cx = 1 + o % (TW - 1); }
                    else {
                        const int bo = o - NINT;
                        if (bo < CW) { cy = 0; cx = bo; }
                        else if (bo < 2 * CW) { cy = TH; cx = bo - CW; }
                        else if (bo < 2 * CW + TH - 1) { cy = 1 + (bo - 2 * CW); cx = 0; }
                        else { cy = 1 + (bo - 2 * CW - (TH - 1)); cx = TW; }
                    }
                    d = (1u << 31) | ((unsigned)(cy * CW + cx) << 19) | ((unsigned)cy << 27);
                    const int a = y0 - 1 + cy, b = x0 - 1 + cx;
                    if (a >= 0 && b >= 0) {
                        const uint4 bits = p.cells[(size_t)e * HW + a * wl + b];
                        const uint32_t wsel = (r < 4) ? ((r < 2) ? bits.x : bits.y) : ((r < 6) ? bits.z : bits.w);
                        const uint32_t rb = (wsel >> ((r & 1) * 16)) & 0xffffu;
                        if (rb != 0xAAAAu)                              // rows with only ignored pixels send nothing
                            d |= rb | ((b + 1 < wl) ? (1u << 16) : 0u) | ((a + 1 < h) ? (1u << 17) : 0u) | (1u << 18);
                    }
                }
                hr_desc[m] = d;
                // rows without a label boundary take the cheaper uniform form; decided per warp and round (no divergence)
                const uint32_t rbits = d & 0xffffu;
                if (__all_sync(0xffffffffu, !(d & (1u << 18)) || rbits == 0u || rbits == 0x5555u)) hr_uniform |= 1u << m;
            }
            const float2 c01 = p.cw[e];
            {
                unsigned it = 0;
                while (!mbar_try_wait(mbar, tma_parity)) { if (++it > RES_SPIN_LIMIT) { ok = false; break; } }
                tma_parity ^= 1u;
            }
            __syncthreads();                                           // S2
            // ---------------- fixed-point unit of the episode: power of two from the group-wide max|F| ----------------
            if constexpr (TM) {
                // the same scan moves the tile into tensor memory: thread = channel (lane 32 (warp % 4) + lane of its quarter, chunk
                // warp / 4), its 100 pixels = columns 100 (warp / 4) .. + 99 in the order of tm_col2pix. From here on shared memory is
                // free of the tile until the next episode is staged.
                unsigned mb = 0u;
                const float* frow = F + (size_t)tid * 100;
                const uint32_t tcol = tq + 100u * (uint32_t)(warp >> 2);
                float v[32];
#pragma unroll
                for (int b = 0; b < 4; ++b) {
#pragma unroll
                    for (int i = 0; i < (b < 3 ? 32 : 4); ++i) {
                        v[i] = frow[tm_col2pix(32 * b + i)];            // column order of the tile: tm_col2pix
                        mb = max(mb, __float_as_uint(v[i]) & 0x7fffffffu);
                    }
                    if (b < 3) tmem_st32(tcol + 32u * b, v); else tmem_st4(tcol + 96u, v);
                }
                tmem_wait_st();
                mb = __reduce_max_sync(0xffffffffu, mb);
                if (lane == 0) atomicMax(&smax[0], mb);
                tmem_fence_before_sync();
            } else {
                unsigned mb = 0u;
                const uint4* F4u = reinterpret_cast<const uint4*>(F);
                for (int i = tid; i < C * NQ; i += CT) {
                    const uint4 f = F4u[i];
                    mb = max(max(mb, f.x & 0x7fffffffu), max(max(f.y & 0x7fffffffu, f.z & 0x7fffffffu), f.w & 0x7fffffffu));
                }
                mb = __reduce_max_sync(0xffffffffu, mb);      // |x| as bits: ordered like the values; NaN / Inf sort above every finite
                if (lane == 0) atomicMax(&smax[0], mb);
            }
            __syncthreads();                                           // S3
            if (TM) tmem_fence_after_sync();
            if constexpr (TM) {
                // The staging buffer is free now: it takes the coefficient table of the full-resolution stage. The gradient mass a
                // cell row sends to its left / right low-res column is linear in the 8 sigmoids p_s of the row,
                //     gl = sum_s L_s p_s - L0,  gr = sum_s R_s p_s - R0,   L_s = A_s (1 - s/8), R_s = A_s s/8,
                //     A_s = c0 | c1 | 0 (label 0 | 1 | ignored),  L0 / R0 = the same sums over B_s = c1 [label 1],
                // so every live row costs the same 16 FMAs whatever its labels are (no uniform / boundary distinction, no
                // label selects inside the step loop): 20 floats per task, rows 80 B apart (conflict-free LDS.128).
                float* tab = F;
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    const unsigned d = hr_desc[m];
                    if ((d >> 31) && (d & (1u << 18))) {
                        float Ls[8], Rs[8], L0 = 0.f, R0 = 0.f;
#pragma unroll
                        for (int sx = 0; sx < 8; ++sx) {
                            const unsigned code = (d >> (2 * sx)) & 3u;
                            const float A = code == 0u ? c01.x : (code == 1u ? c01.y : 0.f), B = code == 1u ? c01.y : 0.f;
                            const float fr = (float)sx * 0.125f;
                            Ls[sx] = A * (1.f - fr); Rs[sx] = A * fr;
                            L0 = fmaf(B, 1.f - fr, L0); R0 = fmaf(B, fr, R0);
                        }
                        float4* t4 = reinterpret_cast<float4*>(tab + (size_t)(m * CT + tid) * 20);
                        t4[0] = make_float4(Ls[0], Ls[1], Ls[2], Ls[3]); t4[1] = make_float4(Ls[4], Ls[5], Ls[6], Ls[7]);
                        t4[2] = make_float4(Rs[0], Rs[1], Rs[2], Rs[3]); t4[3] = make_float4(Rs[4], Rs[5], Rs[6], Rs[7]);
                        t4[4] = make_float4(L0, R0, 0.f, 0.f);
                    }
                }
            }
            if (tid == 0) st_tagged(&fmaxw[k], __uint_as_float(smax[0]), (unsigned)e + 1u);
            for (int kk = tid; kk < CPG; kk += CT)
                atomicMax(&smax[1], __float_as_uint(poll_word(&fmaxw[kk], (unsigned)e + 1u, p.abort_flag)));
            __syncthreads();                                           // S4
            const bool ep_finite = smax[1] < 0x7f800000u;
            float fx_inv, fx_unit;
            fixed_point_unit(smax[1], p.T, KB, fx_unit, fx_inv);
            unsigned long long* acc_ep = p.sums + RES_ACC(acc_offset(e, group, p.G, p.SPL, C));
            // NA == 0: thread c keeps W0[c] / W1[c] and the previous cumulative sums of its accumulator words in registers
            float w0r = (TM && NA == 0) ? W0[tid & (TC ? TC - 1 : 0)] : 0.f, w1r = (TM && NA == 0) ? W1[tid & (TC ? TC - 1 : 0)] : 0.f;
            long long cum0 = 0ll, cum1 = 0ll, tk_ar = 0ll;
            (void)fx_unit; (void)cum0; (void)cum1; (void)tk_ar;

            for (int t = 0; t < p.T; ++t, ++gstep) {
                long long tk0 = 0, tstep0 = 0;
                if (PROF && tid == 0) { tk0 = clock64(); tstep0 = tk0; }
                // ------------ P1 (each chunk first picks up the previous step's all-reduced dW) ------------
                if constexpr (TM) {
                    // Tile in tensor memory: this thread owns lane L = 32 (warp % 4) + lane, i.e. the channels 128 j + L, and the
                    // four warps of a lane quarter split the pixels (25 each). In-thread sum over the 4 chunks, butterfly
                    // reduce-scatter over the 32 lanes (31 shuffles: lane i ends with pixel i of the block), then the four lane
                    // quarters are combined through shared memory in a fixed order. The loads do not depend on the all-reduce
                    // (the features never change): a chunk's columns are on their way while the thread waits for its Wd.
                    const uint32_t tcol = tq + 25u * (uint32_t)(warp >> 2);
                    const float* Wl = Wd + (warp & 3) * 32 + lane;
                    // (26 columns per chunk: the 26th belongs to the next pixel block and only fills the last packed-fp32 pair)
                    float z[32], fj[26];
                    f32x2 z2[13];
#pragma unroll
                    for (int i = 0; i < 13; ++i) z2[i] = 0ull;
#pragma unroll
                    for (int j = 0; j < RES_KCH; ++j) {
                        tmem_ld16(tcol + 100u * j, fj); tmem_ld8(tcol + 100u * j + 16u, fj + 16); tmem_ld2(tcol + 100u * j + 24u, fj + 24);
                        if (NA > 0 && t > 0) wait_applied(j, gstep - 1u);
                        const float wd = (RES_VARIANT & 0x10) ? 0.f : Wl[j * 128];
                        const f32x2 wd2 = pk2(wd, wd);
                        tmem_wait_ld();
#pragma unroll
                        for (int i = 0; i < 26; ++i) asm volatile("" : "+f"(fj[i]));      // keeps the uses below the tcgen05.wait::ld
#pragma unroll
                        for (int i = 0; i < 13; ++i) z2[i] = fma2(wd2, pk2(fj[2 * i], fj[2 * i + 1]), z2[i]);     // FFMA2: two pixels per issue slot
                    }
#pragma unroll
                    for (int i = 0; i < 13; ++i) upk2(z2[i], z[2 * i], z[2 * i + 1]);
#pragma unroll
                    for (int i = 25; i < 32; ++i) z[i] = 0.f;
#pragma unroll
                    for (int sft = 16; sft >= 1; sft >>= 1) {
                        const bool hi = (lane & sft) != 0;
#pragma unroll
                        for (int i = 0; i < sft; ++i) {
                            const float send = hi ? z[i] : z[i + sft], keep = hi ? z[i + sft] : z[i];
                            z[i] = keep + __shfl_xor_sync(0xffffffffu, send, sft);
                        }
                    }
                    if (lane < 25) scratch[(warp & 3) * NP + 25 * (warp >> 2) + lane] = z[0];
                    compute_sync<CT>();
                    if (tid < NP) {
                        const float zz = (scratch[tid] + scratch[NP + tid]) + (scratch[2 * NP + tid] + scratch[3 * NP + tid]);
                        const int pix = tm_col2pix(tid), py = pix / TW, px = pix - py * TW;      // column tid of the tile
                        zt[(py + 1) * ZW + px + 1] = zz;
                        st_tagged(&zll[(gstep & 1u) * (unsigned)HW + (y0 + py) * wl + x0 + px], zz, gstep + 1u);
                    }
                    // wake the halo warp (named barrier 5, non-blocking on this side): the neighbours publish about now, so every
                    // earlier poll of the ring would only be traffic on lines that are being written
                    if (RES_HALO_GATE) asm volatile("bar.arrive 5, %0;" ::"n"(CT + 32) : "memory");
                    compute_sync<CT>();
                } else {
                    float4 za = make_float4(0.f, 0.f, 0.f, 0.f);
                    const float4* F4 = reinterpret_cast<const float4*>(F) + tid;
                    const int stride = NQ * NG;
                    int c = grp;
#pragma unroll
                    for (int j = 0; j < RES_KCH; ++j) {
                        if (t > 0) wait_applied(j, gstep - 1u);
                        if (p1_active && !(RES_VARIANT & 0x10)) {
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
                    compute_sync<CT>();
                    if (tid < NP) {
                        float z = 0.f;
#pragma unroll 4
                        for (int g2 = 0; g2 < NG; ++g2) z += scratch[g2 * NP + tid];
                        const int py = tid / TW, px = tid - py * TW;
                        zt[(py + 1) * ZW + px + 1] = z;
                        st_tagged(&zll[(gstep & 1u) * (unsigned)HW + (y0 + py) * wl + x0 + px], z, gstep + 1u);
                    }
                    compute_sync<CT>();
                }
                if (PROF && tid == 0) { long long n = clock64(); t_acc[0] += n - tk0; tk0 = n; }
                // g(q) = sum_r (1-r/8) [gl(q,r) + gr(q-1,r)] + (r/8) [gl(q-w,r) + gr(q-w-1,r)]; 4 threads per own pixel, 2 rows each:
                // the (gl, gr) pairs of rows 2k, 2k+1 of a cell are one aligned 16-byte word -> conflict-free LDS.128.
                // Columns [c_lo, c_hi) of the tile (tensor-memory kernel: tm_col2pix; otherwise column = pixel).
                auto gather_part = [&](int c_lo, int c_hi) {
                    const int cc = c_lo + (tid >> 2), rq = tid & 3;
                    const bool act = cc < c_hi;
                    float s2 = 0.f;
                    if (act) {
                        const int pl = TM ? tm_col2pix(cc) : cc;
                        const int py = pl / TW, px = pl - py * TW;
                        const int ci = (py + 1) * CW + px + 1;             // the cell whose top-left corner is this pixel
                        const int a = y0 + py, b = x0 + px;
                        const float4* sc4 = reinterpret_cast<const float4*>(scratch) + rq;     // [cell][4 row pairs]
                        const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
                        const float4 own = sc4[ci * 4];
                        const float4 lft = (b > 0) ? sc4[(ci - 1) * 4] : zero4;
                        const float4 up = (a > 0) ? sc4[(ci - CW) * 4] : zero4;
                        const float4 ul = (a > 0 && b > 0) ? sc4[(ci - CW - 1) * 4] : zero4;
                        const float ha = (float)(2 * rq) * 0.125f, hb = (float)(2 * rq + 1) * 0.125f;
                        s2 = fmaf(1.f - ha, own.x + lft.y, ha * (up.x + ul.y)) + fmaf(1.f - hb, own.z + lft.w, hb * (up.z + ul.w));
                    }
                    s2 += __shfl_xor_sync(0xffffffffu, s2, 1);
                    s2 += __shfl_xor_sync(0xffffffffu, s2, 2);
                    if (act && rq == 0) gsm[cc] = s2;
                };
                constexpr bool kSplit = TM && (RES_SPLIT != 0);
                f32x2 p3a = 0ull, p3b = 0ull;                     // tensor-memory kernel: {d0, d1}, {d2, d3} of the thread's channel
                const uint32_t p3col = tq + 100u * (uint32_t)(warp >> 2), p3g = smem_u32(gsm);
                // ------------ HR: one task = one row of one cell -> (gl, gr) of that row in shared memory ------------
                // round 0 (interior cells) runs while the halo warp is still fetching the ring
#pragma unroll
                for (int m = 0; m < RES_MAXTASK; ++m) {
                    if (m * CT < 8 * NCELL) {                     // uniform: does this round have tasks at all
                        if constexpr (kSplit) {
                            if (m == 1) {
                                // in the shadow of the halo exchange: gather + P3 over the columns round 0 has completed
                                compute_sync<CT>();
                                gather_part(0, TM_EARLY);
                                compute_sync<CT>();
                                if (!(RES_VARIANT & 0x08)) {
                                    tm_p3_block<32>(p3col, p3g, p3a, p3b);
                                    tm_p3_block<8>(p3col + 32u, p3g + 128u, p3a, p3b);
                                    tm_p3_block<4>(p3col + 40u, p3g + 160u, p3a, p3b);
                                }
                            }
                        }
                        if (!(RES_VARIANT & 0x01) && m == (round0_interior ? 1 : 0)) {
                            long long tw0 = 0;
                            if (PROF && tid == 0) tw0 = clock64();
                            unsigned it = 0;
                            while (!mbar_try_wait(halo_ready, gstep & 1u)) {
                                if ((++it & 0xfffu) == 0u) {
                                    if (*reinterpret_cast<volatile unsigned*>(p.abort_flag) != 0u) break;
                                    if (it > RES_SPIN_LIMIT) { atomicExch(p.abort_flag, 1u); break; }
                                }
                            }
                            if (PROF && tid == 0) t_acc[1] += clock64() - tw0;
                        }
                        unsigned d = hr_desc[m];
                        // TM: keep what is derived from the descriptor (shared-memory addresses of the four corners, ...) out of the
                        // loop-invariant registers: they spill beside the sweeps' buffers and local-memory loads cost a long scoreboard
                        if (TM) asm volatile("" : "+r"(d));
                        if (d >> 31) {
                            const int ci = (d >> 19) & 0xff, r = tid & 7;
                            float gl = 0.f, gr = 0.f;
                            if ((RES_VARIANT & 0x02) && (d & (1u << 18))) {
                                gl = zt[ci + (int)((d >> 27) & 0xfu)] * 1e-6f;
                            } else if (TM && (d & (1u << 18))) {
                                const int zi = ci + (int)((d >> 27) & 0xfu);     // cy * ZW + cx = ci + cy
                                const int dx = (d >> 16) & 1, dy = (d & (1u << 17)) ? ZW : 0;
                                const float z00 = zt[zi], z01 = zt[zi + dx], z10 = zt[zi + dy], z11 = zt[zi + dy + dx];
                                const float fr = (float)r * 0.125f;
                                const float left = fmaf(fr, z10 - z00, z00), right = fmaf(fr, z11 - z01, z01);
                                float pr[8];
                                sigmoid_row(left, right, pr);
                                const float4* t4 = reinterpret_cast<const float4*>(F + (size_t)(m * CT + tid) * 20);
                                const float4 la = t4[0], lb = t4[1], ra = t4[2], rb4 = t4[3], k0 = t4[4];
                                gl = (fmaf(la.x, pr[0], fmaf(la.y, pr[1], fmaf(la.z, pr[2], la.w * pr[3]))) +
                                      fmaf(lb.x, pr[4], fmaf(lb.y, pr[5], fmaf(lb.z, pr[6], lb.w * pr[7])))) - k0.x;
                                gr = (fmaf(ra.x, pr[0], fmaf(ra.y, pr[1], fmaf(ra.z, pr[2], ra.w * pr[3]))) +
                                      fmaf(rb4.x, pr[4], fmaf(rb4.y, pr[5], fmaf(rb4.z, pr[6], rb4.w * pr[7])))) - k0.y;
                            } else if (d & (1u << 18)) {
                                const int zi = ci + (int)((d >> 27) & 0xfu);     // cy * ZW + cx = ci + cy
                                const int dx = (d >> 16) & 1, dy = (d & (1u << 17)) ? ZW : 0;
                                const float z00 = zt[zi], z01 = zt[zi + dx], z10 = zt[zi + dy], z11 = zt[zi + dy + dx];
                                const float fr = (float)r * 0.125f;
                                const float left = fmaf(fr, z10 - z00, z00), right = fmaf(fr, z11 - z01, z01);
                                if (hr_uniform & (1u << m)) hires_row_uniform(left, right, (d & 1u) != 0u, c01.x, c01.y, gl, gr);
                                else hires_row(left, right, d & 0xffffu, c01.x, c01.y, gl, gr);
                            }
                            *reinterpret_cast<float2*>(scratch + 2 * (ci * 8 + r)) = make_float2(gl, gr);   // [cell][row][gl,gr]
                        }
                    }
                }
                compute_sync<CT>();
                gather_part(kSplit ? TM_EARLY : 0, NP);
                compute_sync<CT>();
                if (PROF && tid == 0) { long long n = clock64(); t_acc[2] += n - tk0; tk0 = n; }
                // ------------ P3: dW = g . F^T chunk by chunk; every partial goes straight to the channel's accumulator word,
                // so chunk j's all-reduce runs under the rest of P3 and the next P1 ------------
                unsigned long long* acc_t = acc_ep + RES_ACC((t & 1) * C);          // even / odd steps use different words (see header)
                if constexpr (TM) {
                    // Tile in tensor memory: thread = channel tid (lane L of its quarter, chunk warp / 4 = columns 100 (warp / 4) ..),
                    // its 100 pixels come in blocks of 32 / 16 / 8 / 4 columns; g is broadcast from shared memory.
                    // No cross-thread reduction and no hand-over between the quads: all 16 warps sweep at once (the tensor-memory
                    // read path is ~3x the shared-memory pipe) and the four chunks leave together. Same summation order as the
                    // shared-memory sweep (pixel p goes to accumulator p mod 4).
                    if (!(RES_VARIANT & 0x08)) {
                        if constexpr (kSplit) {                       // columns 44 .. 99 (0 .. 43 went before the halo wait)
                            tm_p3_block<32>(p3col + 44u, p3g + 176u, p3a, p3b);
                            tm_p3_block<16>(p3col + 76u, p3g + 304u, p3a, p3b);
                            tm_p3_block<8>(p3col + 92u, p3g + 368u, p3a, p3b);
                        } else {
                            tm_p3_block<32>(p3col, p3g, p3a, p3b);
                            tm_p3_block<32>(p3col + 32u, p3g + 128u, p3a, p3b);
                            tm_p3_block<32>(p3col + 64u, p3g + 256u, p3a, p3b);
                            tm_p3_block<4>(p3col + 96u, p3g + 384u, p3a, p3b);
                        }
                    } else {
                        p3a = pk2(gsm[tid & 63], 0.f);
                    }
                    const f32x2 da = p3a, db = p3b;
                    float d0, d1, d2, d3;
                    upk2(da, d0, d1); upk2(db, d2, d3);
                    const float d = (d0 + d1) + (d2 + d3);
                    red_add_u64(acc_t + RES_ACC(tid), ((unsigned long long)__float2ll_rn(d * fx_inv) << KB) + 1ull);
                    if constexpr (NA == 0) {
                        // No applier warp: every compute thread waits for the all-reduced sum of ITS channel (one 8-byte load per
                        // lane and round: a poll costs the warp ~8 sectors, where one applier warp polling all 512 words needs
                        // ~1 600 clk per round), applies the SGD step to the weights it keeps in registers and publishes Wd.
                        if (PROF && tid == 0) tk_ar = clock64();
                        const unsigned long long expect = (unsigned long long)CPG * (unsigned)(t / 2 + 1), cnt_mask = (1ull << KB) - 1ull;
                        const unsigned long long* wp = acc_t + RES_ACC(tid);
                        unsigned long long wv = 0ull;
                        if (!(RES_VARIANT & 0x04)) {
                            auto done = [&](unsigned long long v) {
                                return (RES_VARIANT & 0x1f) ? (v & cnt_mask) >= expect : (v & cnt_mask) == expect;
                            };
                            auto gap = [&](long long n) { const long long t0g = clock64(); while (clock64() - t0g < n) { } };
                            // (measured and dropped: three staggered loads in flight per thread, polls on a %globaltimer grid, a returning
                            // atomic instead of RED + delay — all 1-8 % slower: more traffic on the lines the atomics are working on)
                            if (RES_POLL_DELAY > 0) gap(RES_POLL_DELAY);      // a load issued with the RED overtakes it and is wasted
                            unsigned it = 0;
                            for (;;) {
                                wv = ld_relaxed_u64(wp);
                                if (PROF && tid == 0) t_acc[7] += 1;
                                if (done(wv)) break;
                                if ((++it & 0xffu) == 0u) {
                                    if (*reinterpret_cast<volatile unsigned*>(p.abort_flag) != 0u) break;
                                    if (it > (RES_SPIN_LIMIT >> 2)) { atomicExch(p.abort_flag, 1u); break; }
                                }
                            }
                        }
                        const long long cur = (long long)wv >> KB;                       // cumulative sum over steps 0..t of this parity (exact)
                        const long long prev = (t & 1) ? cum1 : cum0;
                        if (t & 1) cum1 = cur; else cum0 = cur;
                        const float dw = __ll2float_rn(cur - prev) * fx_unit;            // this step's all-reduced dW
                        w0r = fmaf(p.lr, dw, w0r); w1r = fmaf(-p.lr, dw, w1r);
                        Wd[tid] = w1r - w0r;
                        if (PROF && tid == 0) { t_acc[5] += clock64() - tk_ar; t_acc[4] += 1; }
                        compute_sync<CT>();
                    }
                    if (PROF && lane == 0 && (warp & 3) == 3) *reinterpret_cast<volatile long long*>(&tstore_sm[warp >> 2]) = clock64();
                } else if constexpr (kStatic) {
                    // One thread per channel, no cross-thread reduction; g is broadcast from shared memory. The four warp
                    // quads take turns (quad j = channel chunk j, handed over through named barriers 2..4): the sweep of a
                    // quad is latency-bound anyway, a finished quad moves on to the next step's P1 at once, and chunk j's
                    // partials leave at (j+1)/4 of P3 so that its all-reduce runs under the rest of P3 and the next P1.
                    // (Measured alternatives at E = 64: all quads at once 20.3 ms, 4 threads per channel + shuffles 18.8 ms,
                    // two channels per thread 18.0 ms, this 17.5 ms.)
                    const int j = warp >> 2;
                    if (!(RES_VARIANT & 0x20) && j > 0) asm volatile("bar.sync %0, 256;" ::"r"(1 + j) : "memory");
                    const uint32_t row = smem_u32(F) + (uint32_t)tid * 400u, gad = smem_u32(gsm);
                    constexpr int D = RES_P3_DEPTH;                 // quads in flight per operand
                    float4 fb[D], gb[D];
                    float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
                    if (RES_VARIANT & 0x08) {
                        d0 = gsm[tid & 63];
                    } else {
#pragma unroll
                    for (int i = 0; i < D; ++i) { fb[i] = lds128_v(row + 16 * i); gb[i] = lds128_v(gad + 16 * i); }
#pragma unroll
                    for (int i = 0; i < 25; ++i) {
                        const float4 f = fb[i % D], g = gb[i % D];
                        if (i + D < 25) { fb[i % D] = lds128_v(row + 16 * (i + D)); gb[i % D] = lds128_v(gad + 16 * (i + D)); }
                        d0 = fmaf(g.x, f.x, d0); d1 = fmaf(g.y, f.y, d1); d2 = fmaf(g.z, f.z, d2); d3 = fmaf(g.w, f.w, d3);
                    }
                    }
                    if (!(RES_VARIANT & 0x20) && j < RES_KCH - 1) asm volatile("bar.arrive %0, 256;" ::"r"(2 + j) : "memory");
                    const float d = (d0 + d1) + (d2 + d3);
                    // value and arrival in ONE atomic: (fix(d) << K) + 1
                    red_add_u64(acc_t + RES_ACC(tid), ((unsigned long long)__float2ll_rn(d * fx_inv) << KB) + 1ull);
                    if (PROF && lane == 0 && (warp & 3) == 3) *reinterpret_cast<volatile long long*>(&tstore_sm[j]) = clock64();
                } else if constexpr (kStatic40) {
                    // 2 threads per channel (adjacent lanes), part p owns quads p + 2i (i = 0..4): a quarter-warp reads 4 rows
                    // of 160 B at 32-byte granularity -> conflict-free LDS.128; 128 channels per pass = one chunk
                    const int p3_cl = tid >> 1, p3_part = tid & 1;
                    float4 gq[5];
#pragma unroll
                    for (int i = 0; i < 5; ++i) gq[i] = reinterpret_cast<const float4*>(gsm)[p3_part + 2 * i];
                    const uint32_t row0 = smem_u32(F) + (uint32_t)(p3_cl * 160 + p3_part * 16);
                    float4 f[5];
#pragma unroll
                    for (int i = 0; i < 5; ++i) f[i] = lds128_v(row0 + 32 * i);
#pragma unroll
                    for (int j = 0; j < RES_KCH; ++j) {
                        float d0 = 0.f, d1 = 0.f;
#pragma unroll
                        for (int i = 0; i < 5; ++i) {
                            d0 = fmaf(gq[i].x, f[i].x, d0); d1 = fmaf(gq[i].y, f[i].y, d1);
                            d0 = fmaf(gq[i].z, f[i].z, d0); d1 = fmaf(gq[i].w, f[i].w, d1);
                        }
                        if (j + 1 < RES_KCH) {
                            const uint32_t rown = row0 + (uint32_t)((j + 1) * 128 * 160);
#pragma unroll
                            for (int i = 0; i < 5; ++i) f[i] = lds128_v(rown + 32 * i);
                        }
                        float d = d0 + d1;
                        d += __shfl_xor_sync(0xffffffffu, d, 1);
                        if (p3_part == 0)                          // value and arrival in ONE atomic: (fix(d) << K) + 1
                            red_add_u64(acc_t + RES_ACC(j * 128 + p3_cl),
                                        ((unsigned long long)__float2ll_rn(d * fx_inv) << KB) + 1ull);
                        if (PROF && tid == 0) *reinterpret_cast<volatile long long*>(&tstore_sm[j]) = clock64();
                    }
                } else {
                    const float4* g4 = reinterpret_cast<const float4*>(gsm);
                    for (int c = tid; c < C; c += CT) {
                        const float4* row = reinterpret_cast<const float4*>(F + (size_t)c * NP);
                        float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll 4
                        for (int i = 0; i < NQ; ++i) {
                            const float4 f = row[i], g = g4[i];
                            d0 = fmaf(g.x, f.x, d0); d1 = fmaf(g.y, f.y, d1); d2 = fmaf(g.z, f.z, d2); d3 = fmaf(g.w, f.w, d3);
                        }
                        const float d = (d0 + d1) + (d2 + d3);
                        red_add_u64(acc_t + RES_ACC(c), ((unsigned long long)__float2ll_rn(d * fx_inv) << KB) + 1ull);
                    }
                    if (PROF && tid == 0) {
                        const long long n = clock64();
                        for (int j = 0; j < RES_KCH; ++j) *reinterpret_cast<volatile long long*>(&tstore_sm[j]) = n;
                    }
                }
                if (PROF && tid == 0) { long long n = clock64(); t_acc[3] += n - tk0; t_acc[6] += n - tstep0; }
            }
            // ------------ drain the last step's all-reduce ------------
            if (p.T > 0 && NA > 0) {
#pragma unroll
                for (int j = 0; j < RES_KCH; ++j) wait_applied(j, gstep - 1u);
            }
            if (TM && NA == 0) { W0[tid] = w0r; W1[tid] = w1r; }
            __syncthreads();                                           // S5
            if (k == 0) {
                const bool bad = *reinterpret_cast<volatile unsigned*>(p.abort_flag) != 0u || !ep_finite;   // NaN / Inf features give NaN weights
                for (int c = tid; c < C; c += CT) {
                    p.w[((size_t)e * 2) * C + c] = bad ? __int_as_float(0x7fc00000) : W0[c];
                    p.w[((size_t)e * 2 + 1) * C + c] = bad ? __int_as_float(0x7fc00000) : W1[c];
                }
            }
        }
        if (PROF && tid == 0 && p.prof) {
            for (int i = 0; i < 6; ++i) p.prof[(size_t)blockIdx.x * RES_NPROF + i] = t_acc[i];
            p.prof[(size_t)blockIdx.x * RES_NPROF + 10] = t_acc[6];
            p.prof[(size_t)blockIdx.x * RES_NPROF + 11] = t_acc[7];       // tensor-memory kernel: poll rounds of thread 0
        }
        if (tid == 0 && !ok) atomicExch(p.abort_flag, 1u);
        if (TM && warp == 0) { tmem_fence_after_sync(); tmem_dealloc_512(tmem_slot); }    // every warp's last tensor-memory load precedes S5
    } else if (NA > 0 && warp < CT / 32 + NA) {
        // =====================================================================================================
        // applier warp: accumulator words -> SGD update in shared memory -> release the compute warps
        // =====================================================================================================
        long long a_acc[4] = {0, 0, 0, 0};               // PROF: poll rounds, cycles polling, sum(done - own store), chunks
        const unsigned long long cnt_mask = (1ull << KB) - 1ull;
        constexpr int NAPPL = NA > 0 ? NA : 1;
        const int aw = warp - CT / 32;                   // applier number: owns chunks aw, aw + NAPPL, ...
        for (int e = group; e < p.E; e += p.G) {
            __syncthreads();                                           // S1
            for (int c = aw * 32 + lane; c < 2 * C; c += 32 * NAPPL) cum[c] = 0ll;
            __syncthreads();                                           // S2
            __syncthreads();                                           // S3
            __syncthreads();                                           // S4
            float fx_unit;
            { float inv; fixed_point_unit(smax[1], p.T, KB, fx_unit, inv); }
            const unsigned long long* acc_ep = p.sums + RES_ACC(acc_offset(e, group, p.G, p.SPL, C));
            for (int t = 0; t < p.T; ++t) {
                const unsigned long long expect = (unsigned long long)CPG * (unsigned)(t / 2 + 1);   // arrivals so far on every word of this parity
                long long* cum_t = cum + (t & 1) * C;
                for (int j = aw; j < RES_KCH; j += NAPPL) {
                    const unsigned long long* sw = acc_ep + RES_ACC((t & 1) * C + j * CCH);
                    for (int c0 = 0; c0 < CCH; c0 += 32 * RES_AW) {
                        // every load of this lane is issued before any count is looked at (independent L2 round
                        // trips); words that are not complete yet are re-requested together, again as one batch
                        unsigned long long wv[RES_AW];
                        unsigned pending = 0u, it = 0u;
                        long long tp0 = 0;
                        if (PROF) tp0 = clock64();
#pragma unroll
                        for (int m = 0; m < RES_AW; ++m) { wv[m] = 0ull; if (c0 + m * 32 + lane < CCH) pending |= 1u << m; }
                        if (RES_VARIANT & 0x100) pending = 0u;
                        while (pending != 0u) {
#pragma unroll
                            for (int m = 0; m < RES_AW; ++m)
                                if (pending & (1u << m)) wv[m] = ld_relaxed_u64(&sw[RES_ACC(c0 + m * 32 + lane)]);
#pragma unroll
                            for (int m = 0; m < RES_AW; ++m)
                                if ((pending & (1u << m)) && ((RES_VARIANT & 0x1f) ? (wv[m] & cnt_mask) >= expect : (wv[m] & cnt_mask) == expect))
                                    pending &= ~(1u << m);
                            if (pending && (++it & 0xffu) == 0u) {
                                if (*reinterpret_cast<volatile unsigned*>(p.abort_flag) != 0u) break;
                                if (it > (RES_SPIN_LIMIT >> 2)) { atomicExch(p.abort_flag, 1u); break; }
                            }
                            if (PROF) ++a_acc[0];
                        }
                        if (PROF) {
                            __syncwarp();
                            const long long n = clock64();
                            a_acc[1] += n - tp0;
                            a_acc[2] += n - *reinterpret_cast<volatile long long*>(&tstore_sm[j]);
                            a_acc[3] += 1;
                        }
#pragma unroll
                        for (int m = 0; m < RES_AW; ++m) {
                            const int cl = c0 + m * 32 + lane;
                            if (cl < CCH) {
                                const int c = j * CCH + cl;
                                const long long cur = (long long)wv[m] >> KB;                  // cumulative sum over steps 0..t (exact)
                                const float dw = __ll2float_rn(cur - cum_t[c]) * fx_unit;        // this step's all-reduced dW
                                cum_t[c] = cur;
                                const float n0 = fmaf(p.lr, dw, W0[c]), n1 = fmaf(-p.lr, dw, W1[c]);
                                W0[c] = n0; W1[c] = n1; Wd[c] = n1 - n0;
                            }
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&applied[j]);          // release.cta: the Wd stores above are visible to the waiters
                }
            }
            __syncthreads();                                           // S5
        }
        if (PROF && lane == 0 && aw == 0 && p.prof) {
            for (int i = 0; i < 4; ++i) p.prof[(size_t)blockIdx.x * RES_NPROF + 6 + i] = a_acc[i];
        }
    } else {
        // =====================================================================================================
        // halo warp: fetch the ring of z around the tile, step after step
        // =====================================================================================================
        // the ring positions of this lane (static for the whole kernel): index into zt, global pixel or -1
        int ring_zi[RES_HWORDS], ring_q[RES_HWORDS];
#pragma unroll
        for (int m = 0; m < RES_HWORDS; ++m) {
            const int i = lane + 32 * m;
            int zy = 0, zx = 0;
            ring_zi[m] = -1; ring_q[m] = -1;
            if (i < NRING) {
                if (i < ZW) { zy = 0; zx = i; }
                else if (i < 2 * ZW) { zy = ZH - 1; zx = i - ZW; }
                else if (i < 2 * ZW + (ZH - 2)) { zy = 1 + (i - 2 * ZW); zx = 0; }
                else { zy = 1 + (i - 2 * ZW - (ZH - 2)); zx = ZW - 1; }
                const int a = y0 - 1 + zy, b = x0 - 1 + zx;
                ring_zi[m] = zy * ZW + zx;
                if (a >= 0 && a < h && b >= 0 && b < wl) ring_q[m] = a * wl + b;
            }
        }
        unsigned gstep = 0;
        for (int e = group; e < p.E; e += p.G) {
            __syncthreads();                                           // S1
            __syncthreads();                                           // S2
            __syncthreads();                                           // S3
            __syncthreads();                                           // S4
            for (int t = 0; t < p.T; ++t, ++gstep) {
                const unsigned long long* zsrc = zll + (gstep & 1u) * (unsigned)HW;
                unsigned bits[RES_HWORDS], tag[RES_HWORDS];
                unsigned pending = 0u, it = 0u;
                if (TM && RES_HALO_GATE) asm volatile("bar.sync 5, %0;" ::"n"(CT + 32) : "memory");     // this CTA's z of the step is out
#pragma unroll
                for (int m = 0; m < RES_HWORDS; ++m) { bits[m] = 0u; tag[m] = 0u; if (ring_q[m] >= 0) pending |= 1u << m; }
                if (RES_VARIANT & 0x200) pending = 0u;
                while (pending) {
#pragma unroll
                    for (int m = 0; m < RES_HWORDS; ++m)
                        if (pending & (1u << m)) ld_tagged(&zsrc[ring_q[m]], bits[m], tag[m]);
#pragma unroll
                    for (int m = 0; m < RES_HWORDS; ++m)
                        if ((pending & (1u << m)) && ((RES_VARIANT & 0x1f) ? tag[m] >= gstep + 1u : tag[m] == gstep + 1u)) pending &= ~(1u << m);
                    if (pending && (++it & 0xffu) == 0u) {
                        if (*reinterpret_cast<volatile unsigned*>(p.abort_flag) != 0u) break;
                        if (it > (RES_SPIN_LIMIT >> 2)) { atomicExch(p.abort_flag, 1u); break; }
                    }
                }
#pragma unroll
                for (int m = 0; m < RES_HWORDS; ++m)
                    if (ring_zi[m] >= 0) zt[ring_zi[m]] = (ring_q[m] >= 0) ? __uint_as_float(bits[m]) : 0.f;   // outside the image: 0
                // a tile without neighbours has nothing to poll: pace this warp on the tile's own z instead, so that it can
                // never run a whole mbarrier phase ahead of the compute warps
                if (CPG == 1 && lane == 0) (void)poll_word(&zsrc[y0 * wl + x0], gstep + 1u, p.abort_flag);
                __syncwarp();
                if (lane == 0) mbar_arrive(halo_ready);
            }
            __syncthreads();                                           // S5
        }
    }
}

// ---- host side ---------------------------------------------------------------------------------
struct ResidentPlan { int TW, TH, NP, CPG, G, CT, BPS; size_t smem; bool ok; };     // CT compute threads, BPS CTAs per SM

// best tile for CT compute threads and `bps` CTAs per SM
static ResidentPlan plan_resident_cfg(int E, int C, int h, int w, int n_sm, size_t smem_cap, size_t smem_sm, int CT, int bps,
                                      int ftw, int fth) {
    ResidentPlan best{0, 0, 0, 0, 0, CT, bps, 0, false};
    const int HW = h * w;
    double best_score = -1.0;
    if (C % RES_KCH) return best;
    for (int TW = 4; TW <= w; TW += 4) {
        if (w % TW) continue;
        for (int TH = 1; TH <= h && TH <= 14; ++TH) {
            if (h % TH) continue;
            if (ftw && (TW != ftw || TH != fth)) continue;
            const int NP = TW * TH;
            if (NP > CT / 4) continue;                                    // P1 pixel-quad mapping (at least 16 channel groups)
            if (8 * (TW + 1) * (TH + 1) > RES_MAXTASK * CT) continue;     // HR task descriptors
            if (2 * (TW + 2) + 2 * TH > 32 * RES_HWORDS) continue;        // halo ring words
            const size_t sm = res_smem_layout(C, TW, TH, CT).total;
            if (sm > smem_cap || (sm + 1024) * bps > smem_sm) continue;
            const int CPG = HW / NP;
            if (CPG > n_sm * bps) continue;
            int G = n_sm * bps / CPG;
            // most concurrent episodes per on-chip byte first, then the largest tile, then the fewest cells. The score uses
            // the groups the chip could hold, not min(that, E): a small batch must NOT fall to a smaller tile with more CTAs
            // per episode — measured on B200 at E = 1 (60x60x512, 200 steps): 20x5 (36 CTAs) 1.09 ms, 12x3 (100 CTAs) 1.68 ms,
            // 20x3 (60 CTAs) 1.60 ms (tools/resident_small_e.py): the exchange latencies grow with the CTAs per group.
            const double score = (double)G / NP + 1e-9 * NP - 1e-12 * (TW + 1) * (TH + 1);
            if (G > E) G = E;
            if (score > best_score) { best_score = score; best = ResidentPlan{TW, TH, NP, CPG, G, CT, bps, sm, true}; }
        }
    }
    return best;
}

// CWT_RESIDENT_TILE=TWxTH forces a tile, CWT_RESIDENT_BPS=1|2 the number of CTAs per SM
static ResidentPlan plan_resident(int E, int C, int h, int w, int n_sm, size_t smem_cap, size_t smem_sm) {
    int ftw = 0, fth = 0, fbps = 0;
    if (const char* s = getenv("CWT_RESIDENT_TILE")) sscanf(s, "%dx%d", &ftw, &fth);
    if (const char* s = getenv("CWT_RESIDENT_BPS")) fbps = atoi(s);
    const ResidentPlan one = plan_resident_cfg(E, C, h, w, n_sm, smem_cap, smem_sm, 512, 1, ftw, fth);
    const ResidentPlan two = plan_resident_cfg(E, C, h, w, n_sm, smem_cap, smem_sm, 256, 2, ftw, fth);
    if (fbps == 1) return one;
    if (fbps == 2) return two;
    // measured on B200 (60x60x512, E = 64): one CTA per SM 18.8 ms, two CTAs per SM 25.4 ms — the step is bound by the
    // exchange latencies, which grow with the number of CTAs per group, not by the per-SM pipes
    return one.ok ? one : two;
}

static void resident_device_limits(int& n_sm, int& smem_cap, int& smem_sm) {
    int dev = 0, a = 0, b = 0, c = 0;
    n_sm = 148; smem_cap = 232448; smem_sm = 233472;   // B200 (used when no device is visible: sizing only)
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&a, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&b, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&c, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev) == cudaSuccess && a > 0 && b > 0 && c > 0) {
        n_sm = a > 148 ? 148 : a; smem_cap = b; smem_sm = c;
    } else {
        (void)cudaGetLastError();
    }
}

static size_t resident_sum_words(const ResidentPlan& pl, int E, int C) {
    const int nslots = (E + pl.G - 1) / pl.G;
    return RES_ACC((size_t)pl.G * nslots * 2 * C);
}

constexpr int RES_MAXGRID = 320;       // >= 2 * 148

size_t fit_resident_workspace_bytes(int E, int C, int h, int w) {
    const int HW = h * w;
    int n_sm, smem_cap, smem_sm;
    resident_device_limits(n_sm, smem_cap, smem_sm);
    const ResidentPlan pl = plan_resident(E, C, h, w, n_sm, (size_t)smem_cap, (size_t)smem_sm);
    if (!pl.ok) return 0;
    // zll: 2*HW words per group; sums: 2 * C words per episode; fmaxw: one word per CTA;
    // abort flag; profile counters
    return align_up(sizeof(unsigned long long) * 2 * HW * pl.G) + align_up(sizeof(unsigned long long) * resident_sum_words(pl, E, C)) +
           align_up(sizeof(unsigned long long) * RES_MAXGRID) + 512 + align_up(sizeof(long long) * RES_NPROF * RES_MAXGRID);
}

template <int CT, int MINB, int TC, int TTW, int TTH, int TWL, int THL, bool TM = false, int NA = 1>
static int launch_resident(const ResidentParams& p, const ResidentPlan& pl, bool prof, cudaStream_t st) {
    // tensor map over f_s as [E C][h][w] fp32, box TW x TH x 128 channels (tensor-memory kernel only; zeroed otherwise)
    CUtensorMap map;
    memset(&map, 0, sizeof(map));
    if (TM && RES_TMA_STAGE) {
        LsEncodeFn enc = ls_encode_fn();
        CWT_REQUIRE(enc, CWT_ERR_CUDA, "fit_resident: cuTensorMapEncodeTiled is not available from this driver");
        cuuint64_t dims[3] = {(cuuint64_t)p.w_lo, (cuuint64_t)p.h, (cuuint64_t)p.E * p.C};
        cuuint64_t strides[2] = {(cuuint64_t)p.w_lo * 4, (cuuint64_t)p.HW * 4};
        cuuint32_t box[3] = {(cuuint32_t)p.TW, (cuuint32_t)p.TH, 128};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(p.f_s), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        CWT_REQUIRE(r == CUDA_SUCCESS, CWT_ERR_CUDA, "fit_resident: cuTensorMapEncodeTiled failed (%d)", (int)r);
    }
    void* args[] = {&map, const_cast<ResidentParams*>(&p)};
    dim3 grid(pl.G * pl.CPG), block(CT + res_helper_threads(NA));
    const void* fn = prof ? (const void*)k_fit_resident<CT, MINB, TC, TTW, TTH, TWL, THL, true, TM, NA>
                          : (const void*)k_fit_resident<CT, MINB, TC, TTW, TTH, TWL, THL, false, TM, NA>;
    CWT_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    int resident_ctas = 0;
    CWT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident_ctas, fn, CT + res_helper_threads(NA), pl.smem));
    CWT_REQUIRE(resident_ctas >= MINB, CWT_ERR_UNSUPPORTED, "fit_resident: only %d of %d CTAs per SM fit", resident_ctas, MINB);
    CWT_CUDA(cudaLaunchCooperativeKernel(fn, grid, block, args, pl.smem, st));
    count_launch();
    return CWT_OK;
}

// returns CWT_ERR_UNSUPPORTED when the shape does not fit on chip
int fit_resident(const float* f_s, const uint4* cells, const float2* cw, float* w_io, int E, int C, int h, int w,
                 int n_iter, float lr, void* ws, size_t ws_bytes, long long* prof_out, cudaStream_t st) {
    int dev = 0, n_sm = 0, smem_cap = 0, smem_sm = 0, coop = 0;
    CWT_CUDA(cudaGetDevice(&dev));
    resident_device_limits(n_sm, smem_cap, smem_sm);
    CWT_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    CWT_REQUIRE(coop, CWT_ERR_UNSUPPORTED, "fit_resident: device lacks cooperative launch");
    const ResidentPlan pl = plan_resident(E, C, h, w, n_sm, (size_t)smem_cap, (size_t)smem_sm);
    CWT_REQUIRE(pl.ok, CWT_ERR_UNSUPPORTED, "fit_resident: no tile fits C=%d, %dx%d in %d B of shared memory", C, h, w, smem_cap);
    CWT_REQUIRE(pl.G * pl.CPG <= RES_MAXGRID, CWT_ERR_UNSUPPORTED, "fit_resident: grid of %d CTAs", pl.G * pl.CPG);
    const int HW = h * w;
    Carver cv(ws, ws_bytes);
    ResidentParams p{};
    p.zll = cv.take<unsigned long long>((size_t)2 * HW * pl.G);
    p.sums = cv.take<unsigned long long>(resident_sum_words(pl, E, C));
    p.SPL = (E + pl.G - 1) / pl.G;
    p.fmaxw = cv.take<unsigned long long>(RES_MAXGRID);
    p.abort_flag = cv.take<unsigned>(64);
    p.prof = prof_out ? cv.take<long long>((size_t)RES_NPROF * pl.G * pl.CPG) : nullptr;
    CWT_REQUIRE(ws && cv.ok(), CWT_ERR_WORKSPACE, "fit_resident: workspace too small");
    // every tagged word, every accumulator and the abort flag start at zero (step / episode tags start at 1)
    const size_t sync_bytes = (size_t)(reinterpret_cast<char*>(p.abort_flag + 64) - reinterpret_cast<char*>(p.zll));
    CWT_CUDA(cudaMemsetAsync(p.zll, 0, sync_bytes, st));
    p.f_s = f_s; p.cells = cells; p.cw = cw; p.w = w_io;
    p.E = E; p.C = C; p.HW = HW; p.h = h; p.w_lo = w; p.TW = pl.TW; p.TH = pl.TH; p.CPG = pl.CPG; p.G = pl.G; p.T = n_iter; p.lr = lr;
    // arrival-count field: CPG * n_iter arrivals per accumulator word over an episode
    p.KBITS = 1;
    while ((1ll << p.KBITS) <= (long long)pl.CPG * n_iter) ++p.KBITS;
    CWT_REQUIRE(p.KBITS <= 24, CWT_ERR_UNSUPPORTED, "fit_resident: n_iter=%d too large for the on-chip all-reduce", n_iter);
    const bool prof = prof_out != nullptr;
    const bool head = (C == 512 && h == 60 && w == 60);
    int rc;
    if (pl.CT == 512) {
        // PSPNet head geometry: the tile lives in tensor memory for the steps and the compute threads apply the all-reduce
        // themselves (no applier warp). CWT_RESIDENT_TMEM=0: the shared-memory kernel of round 1 (kept for comparison).
        int tm = 1;
        if (const char* s = getenv("CWT_RESIDENT_TMEM")) tm = atoi(s);
        if (head && pl.TW == 20 && pl.TH == 5) {
            if (tm) rc = launch_resident<512, 1, 512, 20, 5, 60, 60, true, 0>(p, pl, prof, st);
            else rc = launch_resident<512, 1, 512, 20, 5, 60, 60, false, 1>(p, pl, prof, st);
        } else rc = launch_resident<512, 1, 0, 0, 0, 0, 0>(p, pl, prof, st);
    } else {
        if (head && pl.TW == 4 && pl.TH == 10) rc = launch_resident<256, 2, 512, 4, 10, 60, 60>(p, pl, prof, st);
        else rc = launch_resident<256, 2, 0, 0, 0, 0, 0>(p, pl, prof, st);
    }
    if (rc != CWT_OK) return rc;
    if (prof_out) CWT_CUDA(cudaMemcpyAsync(prof_out, p.prof, sizeof(long long) * RES_NPROF * pl.G * pl.CPG, cudaMemcpyDeviceToDevice, st));
    return CWT_OK;
}

}  // namespace cwt
