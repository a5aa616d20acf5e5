// (c) query logits -> bilinear up to H x W -> argmax -> intersection / union (+ CE): STREAMING version.
// Replaces src/test.py:192,200-204,214-223 and src/util.py:237-308 in ONE pass over f_q.
//
// One persistent CTA per SM owns a contiguous range of (episode, cell row) pairs and is split into roles:
//
//   producer warp      one TMA tile copy (cp.async.bulk.tensor.2d over the [E*C][h*w] feature matrix, mbarrier complete_tx)
//                      per stage: 32 channels x (<= CB low-res rows x w pixels) = 30 KB, through an NS-deep shared-memory
//                      ring — the HBM stream never stops while the other roles work. (Per-channel 1-D bulk copies were
//                      measured first: one UBLKCP costs ~100 issue cycles per warp, 2.3 TB/s; the tile copy reaches
//                      5.5 TB/s with both compute roles switched off);
//   8 contraction      consume the ring: warp j owns the channels c = j (mod 8); a lane keeps two pixel quads x 2V
//     warps            logit rows (+ |f_p|^2) in packed-fp32 accumulators (FFMA2: one instruction per two pixels; the
//                      weights sit in shared memory as {m, m} pairs). After the C channels of a chunk a warp only dumps
//                      its partial sums into one of two shared-memory buffers and goes on to the next chunk's stages:
//                      the ring never waits for an epilogue;
//   8 up-sample        combine the eight partials of the PREVIOUS chunk in a fixed order, apply F.normalize, write its 2V
//     warps            logit rows to one of two small logit buffers, then: column-threaded bilinear up-sample (ATen's CPU
//                      rounding order, see iou.cu), argmax, confusion counts and CE — all while the next chunk is streaming.
//
// Only new low-res rows are streamed: the last logit row of a chunk is carried over as the first row of the next one,
// so f_q is read exactly once (plus one halo row per CTA range).
#pragma once
#include "common.cuh"
#include "hires.cuh"
#include "tma_pipe.cuh"
#include <type_traits>

namespace cwt {

constexpr int LS_CW = 8;                       // contraction warps
constexpr int LS_UW = 8;                       // up-sample warps (measured: 8 + 8 beats 8 + 11 and 4 + 15)
constexpr int LS_THREADS = 32 * (1 + LS_CW + LS_UW);
#ifndef LS_CHT_N
#define LS_CHT_N 32
#endif
#ifndef LS_NS_MAX
#define LS_NS_MAX 4
#endif
#ifndef LS_TAPER
#define LS_TAPER 0                             // chunks of a CTA's range: 1 = short first chunk(s) and halving last chunks
#endif
constexpr int LS_CHT = LS_CHT_N;               // channels per ring stage = per TMA tile copy
constexpr int LS_MAXQ = 64;                    // pixel quads per chunk (two per lane)
struct LsMaps { CUtensorMap m[4]; };       // m[ns - 1]: box of ns low-res rows

struct LsParams {
    const float* wts;          // [E][2V][C]
    const float* f_q;          // [E][C][h*w]
    const void* lab;           // [E][H][W]
    unsigned long long* counts;// [E][V][2][3]
    float* logits_out;         // [E][2V][h*w] or null
    double* ce;                // [E][V][2] or null
    int E, C, h, w, H, W, CB, NS, CHT, ignore_index, normalize_mask;
    unsigned ce_mask;
};

struct LsSmem { size_t ring, red, ms, lbuf, bars, total; unsigned stage_bytes, row_bytes; };
static __host__ __device__ inline LsSmem ls_smem_layout(int V, int C, int w, int CB, int NS, int CHT) {
    LsSmem s;
    const int R2 = 2 * V;
    s.row_bytes = (unsigned)(CB * w * 4);
    s.stage_bytes = (unsigned)(CHT * s.row_bytes);
    size_t o = 0;
    s.ring = o; o += (size_t)NS * s.stage_bytes; o = (o + 127) / 128 * 128;
    s.red = o;  o += (size_t)2 * LS_CW * (R2 + 1) * LS_MAXQ * 16; o = (o + 127) / 128 * 128;      // two buffers: chunk parity
    s.ms = o;   o += (size_t)C * R2 * 8; o = (o + 127) / 128 * 128;
    s.lbuf = o; o += (size_t)2 * R2 * (CB + 1) * w * 4; o = (o + 127) / 128 * 128;
    s.bars = o; o += 8 * (2 * 8 + 4);
    s.total = o;
    return s;
}

// the chunk sequence of a CTA range [g0, g1) of global cell rows g = e * h + a; identical in every role
struct LsChunk {
    int e, a0, nc, sb, ns, carry;      // episode, first cell row, cell rows, first streamed low-res row, streamed rows, row a0 carried over
};
__device__ __forceinline__ LsChunk ls_chunk(int g, int g0, int g1, int h, int CB) {
    LsChunk c;
    c.e = g / h; c.a0 = g - c.e * h;
    c.carry = (g != g0 && c.a0 != 0) ? 1 : 0;
    int nc = c.carry ? CB : CB - 1;
#if LS_TAPER
    // a short first chunk gives the up-sample warps work after two low-res rows instead of five; halving chunks at the end of
    // the range leave only one cell row to be up-sampled after the last byte has arrived
    if (g == g0) nc = 1;
    else if (g == g0 + 1) nc = min(nc, 2);
    nc = min(nc, max(1, (g1 - g + 1) / 2));
#endif
    if (c.a0 + nc >= h - 1) nc = h - c.a0;               // the last cell row (one hi-res row) needs no further low-res row
    nc = min(nc, g1 - g);
    c.nc = nc;
    c.sb = c.a0 + c.carry;
    c.ns = min(c.a0 + nc, h - 1) - c.sb + 1;
    return c;
}

template <bool I64, int V, bool CE>
__global__ void __launch_bounds__(LS_THREADS, 1) k_logits_iou_stream(const __grid_constant__ LsMaps maps, LsParams p) {
    extern __shared__ __align__(128) unsigned char ls_raw[];
    constexpr int R2 = 2 * V;
    const int C = p.C, h = p.h, w = p.w, HW = h * w, CB = p.CB, NS = p.NS;
    constexpr int CHT = LS_CHT;
    const LsSmem L = ls_smem_layout(V, C, w, CB, NS, CHT);
    unsigned char* ring = ls_raw + L.ring;
    float* red = reinterpret_cast<float*>(ls_raw + L.red);         // [LS_CW][R2+1][LS_MAXQ*4]
    f32x2* Ms = reinterpret_cast<f32x2*>(ls_raw + L.ms);           // [C][R2] {m, m}
    float* lbuf = reinterpret_cast<float*>(ls_raw + L.lbuf);       // [2][R2][CB+1][w]
    uint64_t* full = reinterpret_cast<uint64_t*>(ls_raw + L.bars); // [8]
    uint64_t* empty = full + 8;                                    // [8]
    uint64_t* lfull = empty + 8;                                   // [2]
    uint64_t* lempty = lfull + 2;                                  // [2]
    const int LROWS = CB + 1, LBUF = R2 * LROWS * w;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long N = (long long)p.E * h;
    const int g0 = (int)(N * blockIdx.x / gridDim.x), g1 = (int)(N * (blockIdx.x + 1) / gridDim.x);

    if (tid == 0) {
        for (int i = 0; i < NS; ++i) { ls_mbar_init(&full[i], 1); ls_mbar_init(&empty[i], LS_CW); }
        for (int i = 0; i < 2; ++i) { ls_mbar_init(&lfull[i], 32 * LS_CW); ls_mbar_init(&lempty[i], 32 * LS_UW); }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const int nstage = C / CHT;

    if (warp == 0) {
        // ============================ producer ============================
        unsigned slot = 0, ph = 0;
        for (int g = g0; g < g1;) {
            const LsChunk ck = ls_chunk(g, g0, g1, h, CB);
            g += ck.nc;
            if (ck.ns <= 0) continue;
            const unsigned bytes = (unsigned)(ck.ns * w * 4);
            for (int s = 0; s < nstage; ++s) {
                ls_wait(&empty[slot], ph ^ 1u);
                if (lane == 0) {
                    ls_expect_tx(&full[slot], bytes * CHT);
                    ls_tma_2d(ring + (size_t)slot * L.stage_bytes, &maps.m[ck.ns - 1], ck.sb * w, ck.e * C + s * CHT, &full[slot]);
                }
                __syncwarp();
                if (++slot == (unsigned)NS) { slot = 0; ph ^= 1u; }
            }
        }
    } else if (warp <= LS_CW) {
        // ============================ contraction ============================
        const int cw = warp - 1, ctid = tid - 32;
        unsigned slot = 0, ph = 0;
        int cur_e = -1, chunk = 0;
        for (int g = g0; g < g1; ++chunk) {
            const LsChunk ck = ls_chunk(g, g0, g1, h, CB);
            g += ck.nc;
            if (ck.e != cur_e) {                               // new episode: its 2V weight rows as {m, m} pairs
                asm volatile("bar.sync 1, %0;" ::"n"(32 * LS_CW) : "memory");      // every warp is done with the previous weights
                const float* Mg = p.wts + (size_t)ck.e * R2 * C;
                for (int i = ctid; i < C * R2; i += 32 * LS_CW) {
                    const int r = i / C, c = i - r * C;
                    const float m = Mg[i];
                    Ms[c * R2 + r] = pk2(m, m);
                }
                asm volatile("bar.sync 1, %0;" ::"n"(32 * LS_CW) : "memory");
                cur_e = ck.e;
            }
            const int npx = ck.ns * w, nq = npx >> 2;
            const bool q0 = lane < nq, q1 = lane + 32 < nq;
            f32x2 acc[R2][4], n2[4];
#pragma unroll
            for (int r = 0; r < R2; ++r) { acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0ull; }
            n2[0] = n2[1] = n2[2] = n2[3] = 0ull;
            if (ck.ns > 0) {
                // lane's two quads in channel row cl of a stage: byte offset (cl * npx + 4 * quad) * 4; weights of channel c at
                // Ms[c * R2]. 32-bit shared-memory addresses + explicit ld.shared keep the address arithmetic out of the loop.
                const uint32_t rowb = (uint32_t)npx * 4u;
                const uint32_t lane_base = ls_u32(ring) + (uint32_t)cw * rowb + (uint32_t)lane * 16u;
                uint32_t maddr = ls_u32(Ms) + (uint32_t)cw * (R2 * 8u);
                const ulonglong2 zero2 = make_ulonglong2(0ull, 0ull);
                for (int s = 0; s < nstage; ++s) {
                    ls_wait(&full[slot], ph);
                    const uint32_t st = lane_base + slot * L.stage_bytes;
#pragma unroll
                    for (int k = 0; k < CHT / LS_CW; ++k) {
                        const uint32_t row = st + (uint32_t)(k * LS_CW) * rowb;
                        const ulonglong2 d0 = q0 ? ls_lds128(row) : zero2;
                        const ulonglong2 d1 = q1 ? ls_lds128(row + 512u) : zero2;
#pragma unroll
                        for (int r2 = 0; r2 < V; ++r2) {
                            const ulonglong2 m = ls_lds128(maddr + (uint32_t)(k * LS_CW * R2 * 8 + r2 * 16));   // {m_2r2, m_2r2, m_2r2+1, m_2r2+1}
                            acc[2 * r2][0] = fma2(m.x, d0.x, acc[2 * r2][0]); acc[2 * r2][1] = fma2(m.x, d0.y, acc[2 * r2][1]);
                            acc[2 * r2][2] = fma2(m.x, d1.x, acc[2 * r2][2]); acc[2 * r2][3] = fma2(m.x, d1.y, acc[2 * r2][3]);
                            acc[2 * r2 + 1][0] = fma2(m.y, d0.x, acc[2 * r2 + 1][0]); acc[2 * r2 + 1][1] = fma2(m.y, d0.y, acc[2 * r2 + 1][1]);
                            acc[2 * r2 + 1][2] = fma2(m.y, d1.x, acc[2 * r2 + 1][2]); acc[2 * r2 + 1][3] = fma2(m.y, d1.y, acc[2 * r2 + 1][3]);
                        }
                        n2[0] = fma2(d0.x, d0.x, n2[0]); n2[1] = fma2(d0.y, d0.y, n2[1]);
                        n2[2] = fma2(d1.x, d1.x, n2[2]); n2[3] = fma2(d1.y, d1.y, n2[3]);
                    }
                    maddr += (uint32_t)(CHT * R2 * 8);
                    __syncwarp();
                    if (lane == 0) ls_arrive(&empty[slot]);
                    if (++slot == (unsigned)NS) { slot = 0; ph ^= 1u; }
                }
            }
            // partial sums of this warp -> red[chunk parity][cw][r][px]; the up-sample warps combine the eight partials, so the
            // contraction warps go straight on to the next chunk's stages (the ring never waits for an epilogue)
            {
                const int rb = chunk & 1;
                ls_wait(&lempty[rb], ((chunk >> 1) & 1u) ^ 1u);      // the up-sample warps have read this buffer (two chunks ago)
                ulonglong2* rw = reinterpret_cast<ulonglong2*>(red + ((size_t)rb * LS_CW + cw) * (R2 + 1) * LS_MAXQ * 4);
#pragma unroll
                for (int r = 0; r < R2; ++r) {
                    rw[r * LS_MAXQ + lane] = make_ulonglong2(acc[r][0], acc[r][1]);
                    rw[r * LS_MAXQ + lane + 32] = make_ulonglong2(acc[r][2], acc[r][3]);
                }
                rw[R2 * LS_MAXQ + lane] = make_ulonglong2(n2[0], n2[1]);
                rw[R2 * LS_MAXQ + lane + 32] = make_ulonglong2(n2[2], n2[3]);
                ls_arrive(&lfull[rb]);
            }
        }
    } else {
        // ============================ up-sample / argmax / count ============================
        const int utid = tid - 32 * (1 + LS_CW);
        const int H = p.H, W = p.W;
        using LabT = typename std::conditional<I64, long long, uint8_t>::type;
        int cnt[V][4];
        float loss[V];
        int nvalid = 0, cur_e = -1, chunk = 0;
#pragma unroll
        for (int v = 0; v < V; ++v) { cnt[v][0] = cnt[v][1] = cnt[v][2] = cnt[v][3] = 0; loss[v] = 0.f; }
        bool waited = false;
        auto flush = [&](int e) {
            // programmatic dependent launch: this kernel starts streaming while the kernel that zeroes the outputs is still
            // running; the first global atomic waits for it
            if (!waited) { asm volatile("griddepcontrol.wait;" ::: "memory"); waited = true; }
#pragma unroll
            for (int v = 0; v < V; ++v) {
                int n[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) { n[k] = __reduce_add_sync(0xffffffffu, cnt[v][k]); cnt[v][k] = 0; }
                const float l = warp_sum(loss[v]);
                loss[v] = 0.f;
                if (lane == 0) {
                    const int n00 = n[0], n01 = n[1], n10 = n[2], n11 = n[3];      // [pred][code]
                    unsigned long long* c = p.counts + ((size_t)e * V + v) * 6;
                    const int I0 = n00, I1 = n11, T0 = n00 + n10, T1 = n01 + n11, U0 = n00 + n01 + n10, U1 = n11 + n10 + n01;
                    if (I0) atomicAdd(&c[0], (unsigned long long)I0);
                    if (U0) atomicAdd(&c[1], (unsigned long long)U0);
                    if (T0) atomicAdd(&c[2], (unsigned long long)T0);
                    if (I1) atomicAdd(&c[3], (unsigned long long)I1);
                    if (U1) atomicAdd(&c[4], (unsigned long long)U1);
                    if (T1) atomicAdd(&c[5], (unsigned long long)T1);
                    if (p.ce && l != 0.f) atomicAdd(&p.ce[((size_t)e * V + v) * 2], (double)l);
                }
            }
            const int nv = __reduce_add_sync(0xffffffffu, nvalid);
            nvalid = 0;
            if (p.ce && lane == 0 && nv) {
#pragma unroll
                for (int v = 0; v < V; ++v) atomicAdd(&p.ce[((size_t)e * V + v) * 2 + 1], (double)nv);
            }
        };
        int prev_a0 = 0;
        for (int g = g0; g < g1; ++chunk) {
            const LsChunk ck = ls_chunk(g, g0, g1, h, CB);
            g += ck.nc;
            if (ck.e != cur_e) { if (cur_e >= 0) flush(cur_e); cur_e = ck.e; }
            const int buf = chunk & 1;
            ls_wait(&lfull[buf], (chunk >> 1) & 1u);               // the eight partial sums of this chunk are in red[buf]
            {
                // combine the partials in a fixed order, apply F.normalize, write the chunk's logit rows
                const int npx = ck.ns * w;
                float* lbw = lbuf + (size_t)buf * LBUF;
                const float* rbase = red + (size_t)buf * LS_CW * (R2 + 1) * LS_MAXQ * 4;
                const int row_off = ck.sb - ck.a0;                   // 0, or 1 when row a0 is carried over
                for (int px = utid; px < npx; px += 32 * LS_UW) {
                    float nn = 0.f, sv[R2];
#pragma unroll
                    for (int r = 0; r < R2; ++r) sv[r] = 0.f;
#pragma unroll
                    for (int j = 0; j < LS_CW; ++j) {
                        const float* rj = rbase + (size_t)j * (R2 + 1) * LS_MAXQ * 4;
#pragma unroll
                        for (int r = 0; r < R2; ++r) sv[r] += rj[r * LS_MAXQ * 4 + px];
                        nn += rj[R2 * LS_MAXQ * 4 + px];
                    }
                    const float den = fmaxf(sqrtf(nn), 1e-12f);
                    const int py = px / w, x = px - py * w;
                    const int grow = ck.e * h + ck.sb + py;            // global row: own rows only go to logits_out
#pragma unroll
                    for (int r = 0; r < R2; ++r) {
                        float v = sv[r];
                        if ((p.normalize_mask >> (r >> 1)) & 1) v = v / den;          // F.normalize(f_q, dim=1)
                        lbw[(r * LROWS + row_off + py) * w + x] = v;
                        if (p.logits_out && grow < g1)
                            p.logits_out[((size_t)ck.e * R2 + r) * HW + (size_t)(ck.sb + py) * w + x] = v;
                    }
                }
                ls_arrive(&lempty[buf]);                             // red[buf] may be overwritten (two chunks from now)
                if (ck.carry) {                                       // row a0 = a logit row of the previous chunk
                    const float* lp = lbuf + (size_t)(buf ^ 1) * LBUF;
                    const int pr = ck.a0 - prev_a0;
                    for (int i = utid; i < R2 * w; i += 32 * LS_UW) {
                        const int r = i / w, x = i - r * w;
                        lbw[(r * LROWS) * w + x] = lp[(r * LROWS + pr) * w + x];
                    }
                }
                prev_a0 = ck.a0;
                // the logit rows are complete; also: everyone has finished up-sampling the previous chunk, whose buffer the
                // next chunk's combine overwrites
                asm volatile("bar.sync 2, %0;" ::"n"(32 * LS_UW) : "memory");
            }
            const float* lb = lbuf + (size_t)buf * LBUF;
            const LabT* labp = reinterpret_cast<const LabT*>(p.lab) + (size_t)ck.e * H * W;
            for (int X = utid; X < W; X += 32 * LS_UW) {
                const int b0 = X >> 3, b1 = min(b0 + 1, w - 1);
                const float w1 = (X & 7) * 0.125f, w0 = 1.f - w1;
                // horizontal lerp of the 2V logit rows at this column, once per low-res row: t = fma(l[b0], w0, l[b1] * w1)
                auto hlerp = [&](int row, float (&t)[R2]) {
#pragma unroll
                    for (int r = 0; r < R2; ++r) {
                        const float* lr = lb + (r * LROWS + row) * w;
                        t[r] = __fmaf_rn(lr[b0], w0, __fmul_rn(lr[b1], w1));
                    }
                };
                // the 8 labels of this column in cell row a: unconditional loads (rows past the image re-read its last row)
                auto load_labels = [&](int a, LabT (&raw)[8]) {
                    const int nr1 = min(8, H - 8 * a) - 1;
                    const LabT* q = labp + (size_t)(8 * a) * W + X;
#pragma unroll
                    for (int r = 0; r < 8; ++r) raw[r] = q[(size_t)min(r, nr1) * W];
                };
                float t0[R2], t1[R2];
                hlerp(0, t1);
                LabT nxt[8];
                load_labels(ck.a0, nxt);
                for (int ar = 0; ar < ck.nc; ++ar) {
                    const int a = ck.a0 + ar;
#pragma unroll
                    for (int r = 0; r < R2; ++r) t0[r] = t1[r];
                    if (a + 1 < h) hlerp(ar + 1, t1);
                    const int nr = min(8, H - 8 * a);
                    // label codes, 2 bits per row: 0 / 1 / 2 (ignored, invalid or past the image)
                    unsigned codes = 0u;
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        unsigned c = I64 ? (unsigned)min((unsigned long long)nxt[r], 2ull) : min((unsigned)nxt[r], 2u);
                        c = (r < nr) ? c : 2u;
                        codes |= c << (2 * r);
                    }
                    if (ar + 1 < ck.nc) load_labels(a + 1, nxt);           // in flight while this cell row is evaluated
                    const unsigned one = codes & 0x5555u, inval = (codes >> 1) & 0x5555u, zero = ~(one | inval) & 0x5555u;
                    const int n_one = __popc(one), n_zero = __popc(zero);
                    nvalid += n_one + n_zero;
                    unsigned pm[V];
                    float prod[V], sabs[V], ssgn[V];
#pragma unroll
                    for (int v = 0; v < V; ++v) { pm[v] = 0u; prod[v] = 1.f; sabs[v] = 0.f; ssgn[v] = 0.f; }
                    f32x2 ta0, ta1, tb0, tb1;          // V == 2: {variant 0, variant 1} of channel 0 (a) / channel 1 (b); V == 1: {ch 0, ch 1}
                    if (V == 2) { ta0 = pk2(t0[0], t0[2]); tb0 = pk2(t0[1], t0[3]); ta1 = pk2(t1[0], t1[2]); tb1 = pk2(t1[1], t1[3]); }
                    else { ta0 = pk2(t0[0], t0[1]); ta1 = pk2(t1[0], t1[1]); tb0 = tb1 = 0ull; }
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        const float h1 = r * 0.125f, h0 = 1.f - h1;
                        const f32x2 h0p = pk2(h0, h0), h1p = pk2(h1, h1);
                        float d[V];                                      // up(l1) - up(l0): > 0 <=> argmax = 1 (first index wins ties)
                        if (V == 2) {
                            const f32x2 u0 = fma2(ta0, h0p, mul2(ta1, h1p)), u1 = fma2(tb0, h0p, mul2(tb1, h1p));   // u = fma(t0, h0, t1 * h1)
                            upk2(fma2(u0, pk2(-1.f, -1.f), u1), d[0], d[V - 1]);
                        } else {
                            float u0, u1;
                            upk2(fma2(ta0, h0p, mul2(ta1, h1p)), u0, u1);
                            d[0] = u1 - u0;
                        }
                        const bool valid = !((inval >> (2 * r)) & 1u);
                        const unsigned ysign = (one << (31 - 2 * r)) & 0x80000000u;
#pragma unroll
                        for (int v = 0; v < V; ++v) {
                            if (d[v] > 0.f) pm[v] |= 1u << (2 * r);
                            if (CE) {
                                // -log softmax(u)[y] = softplus(t), t = -+d:  max(t, 0) = (|d| + t) / 2 and the log1p(exp(-|d|)) terms of the
                                // 8 rows are one log2 of a product (an invalid pixel has d := 0 and contributes exactly log2(2) = 1, removed below)
                                const float dz = valid ? d[v] : 0.f;
                                const float e = fast_ex2(-fabsf(dz) * 1.4426950408889634f);
                                prod[v] = fmaf(prod[v], e, prod[v]);
                                sabs[v] += fabsf(dz);
                                ssgn[v] += __uint_as_float(__float_as_uint(dz) ^ ysign);              // t = y ? -d : d
                            }
                        }
                    }
#pragma unroll
                    for (int v = 0; v < V; ++v) {
                        const int n11 = __popc(pm[v] & one), n10 = __popc(pm[v] & zero);          // [pred][label]
                        cnt[v][0] += n_zero - n10; cnt[v][1] += n_one - n11; cnt[v][2] += n10; cnt[v][3] += n11;
                        if (CE)
                            loss[v] += fmaf(0.6931471805599453f, __log2f(prod[v]) - (float)(8 - n_one - n_zero), 0.5f * (sabs[v] + ssgn[v]));
                    }
                }
            }
        }
        if (cur_e >= 0) flush(cur_e);
    }
}

// returns CWT_ERR_UNSUPPORTED (without setting an error) when the shape does not suit the streaming kernel
template <bool I64, int V>
static int launch_logits_iou_stream(const float* wts, const float* f_q, const void* lab, int normalize_mask, unsigned ce_mask,
                                    unsigned long long* counts, float* logits_out, double* ce, int E, int C, int h, int w,
                                    int H, int W, int ignore_index, cudaStream_t st) {
    LsParams p{};
    p.wts = wts; p.f_q = f_q; p.lab = lab; p.counts = counts; p.logits_out = logits_out; p.ce = ce;
    p.E = E; p.C = C; p.h = h; p.w = w; p.H = H; p.W = W; p.ignore_index = ignore_index; p.normalize_mask = normalize_mask;
    p.ce_mask = ce_mask;
    p.CB = 4 * LS_MAXQ / w < 4 ? 4 * LS_MAXQ / w : 4;
    int dev = 0, n_sm = 148, smem_cap = 232448;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&smem_cap, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    LsSmem L{};
    p.CHT = LS_CHT;
    for (p.NS = LS_NS_MAX; p.NS >= 2; --p.NS) {
        L = ls_smem_layout(V, C, w, p.CB, p.NS, p.CHT);
        if (L.total <= (size_t)smem_cap) break;
    }
    if (p.NS < 2) return CWT_ERR_UNSUPPORTED;
    auto kern = ce ? k_logits_iou_stream<I64, V, true> : k_logits_iou_stream<I64, V, false>;
    CWT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    LsMaps maps;
    memset(&maps, 0, sizeof(maps));
    LsEncodeFn enc = ls_encode_fn();
    CWT_REQUIRE(enc, CWT_ERR_CUDA, "logits_iou: cuTensorMapEncodeTiled is not available from this driver");
    for (int ns = 1; ns <= p.CB; ++ns) {
        cuuint64_t dims[2] = {(cuuint64_t)h * w, (cuuint64_t)E * C};
        cuuint64_t strides[1] = {(cuuint64_t)h * w * 4};
        cuuint32_t box[2] = {(cuuint32_t)(ns * w), (cuuint32_t)p.CHT};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&maps.m[ns - 1], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(f_q), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, ls_l2_promotion(),
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        CWT_REQUIRE(r == CUDA_SUCCESS, CWT_ERR_CUDA, "logits_iou: cuTensorMapEncodeTiled failed (%d)", (int)r);
    }
    const long long N = (long long)E * h;
    long long grid = (N + 3) / 4;                      // at least four cell rows per CTA
    if (grid > n_sm) grid = n_sm;
    if (grid < 1) grid = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(LS_THREADS); cfg.dynamicSmemBytes = L.total; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    CWT_CUDA(cudaLaunchKernelEx(&cfg, kern, maps, p));
    CWT_LAUNCHED("logits_iou_stream");
    return CWT_OK;
}

// shapes the streaming kernel takes: 16-byte aligned rows for the bulk copies, whole ring stages, <= 64 quads per chunk
static inline bool logits_iou_stream_ok(const float* f_q, int V, int C, int h, int w, int H, int W) {
    return V <= 2 && w % 4 == 0 && w >= 4 && w <= 128 && C % LS_CHT == 0 && H == 8 * (h - 1) + 1 && W == 8 * (w - 1) + 1 &&
           (reinterpret_cast<uintptr_t>(f_q) & 15u) == 0 && (size_t)C * 2 * V * 8 <= 64 * 1024;
}

}  // namespace cwt
