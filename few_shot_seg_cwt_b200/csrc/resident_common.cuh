// Device helpers shared by the two persistent fit kernels (fit_resident.cu: features resident in shared memory, 1-shot;
// fit_l2.cu: features streamed from L2 by TMA every sweep, any number of shots): mbarrier / bulk-copy wrappers, the
// sequence-tagged 8-byte words of the halo exchange, the 64-bit L2 accumulator words of the all-reduce, the cell-row form of
// the full-resolution stage and the fixed-point unit of the all-reduce.
#pragma once
#include "common.cuh"
#include "hires.cuh"

namespace cwt {

constexpr unsigned RES_SPIN_LIMIT = 1u << 24;

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
template <int CT>
__device__ __forceinline__ void compute_sync() { asm volatile("bar.sync 1, %0;" ::"n"(CT) : "memory"); }

// tagged words: {payload bits, step number} in one 64-bit access (single-copy atomic)
__device__ __forceinline__ void st_tagged(unsigned long long* p, float v, unsigned seq) {
    asm volatile("{\n .reg .b64 t;\n mov.b64 t, {%1, %2};\n st.relaxed.gpu.global.u64 [%0], t;\n}"
                 ::"l"(p), "r"(__float_as_uint(v)), "r"(seq) : "memory");
}
__device__ __forceinline__ void ld_tagged(const unsigned long long* p, unsigned& bits, unsigned& seq) {
    asm volatile("{\n .reg .b64 t;\n ld.relaxed.gpu.global.u64 t, [%2];\n mov.b64 {%0, %1}, t;\n}"
                 : "=r"(bits), "=r"(seq) : "l"(p) : "memory");
}
// fire-and-forget 64-bit add at the L2 atomic unit (SASS: REDG.E.ADD.64) / relaxed 64-bit poll load
__device__ __forceinline__ void red_add_u64(unsigned long long* p, unsigned long long v) {
    // no "memory" clobber: the add depends on registers only
    asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v));
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// 128-bit shared-memory load that the compiler may neither reorder against other volatile asm nor merge
__device__ __forceinline__ float4 lds128_v(uint32_t saddr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
    return v;
}
// ---- tensor memory (TMEM) as a plain 128-lane x 512-column x 32-bit scratchpad -------------------------------------------
// A warp reaches only the lane quarter 32 (warp % 4) .. +31; in the 32x32b shape thread i of the warp owns lane 32 (warp % 4) + i
// and an instruction moves N consecutive columns of that lane from / to N registers (SASS LDTM / STTM). No MMA is involved: the
// fit keeps its feature tile there because tcgen05.ld streams ~380 B/clk/SM (tools/micro/tmem_bench.cu) against 128 B/clk of
// shared memory, without occupying the LSU pipe.
__device__ __forceinline__ void tmem_alloc_512(uint32_t* slot_smem) {       // one full warp; the base address lands in *slot_smem
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot_smem)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_512(uint32_t base) {           // the warp that allocated
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base) : "memory");
}
__device__ __forceinline__ void tmem_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld1(uint32_t a, float* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=f"(r[0]) : "r"(a) : "memory");
}
__device__ __forceinline__ void tmem_ld2(uint32_t a, float* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=f"(r[0]), "=f"(r[1]) : "r"(a) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t a, float* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]) : "r"(a) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t a, float* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]) : "r"(a) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t a, float* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]), "=f"(r[8]),
                   "=f"(r[9]), "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15]) : "r"(a) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t a, float* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]), "=f"(r[8]), "=f"(r[9]),
          "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15]), "=f"(r[16]), "=f"(r[17]), "=f"(r[18]),
          "=f"(r[19]), "=f"(r[20]), "=f"(r[21]), "=f"(r[22]), "=f"(r[23]), "=f"(r[24]), "=f"(r[25]), "=f"(r[26]), "=f"(r[27]),
          "=f"(r[28]), "=f"(r[29]), "=f"(r[30]), "=f"(r[31])
        : "r"(a) : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t a, const float* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%4], {%0,%1,%2,%3};" ::"f"(r[0]), "f"(r[1]), "f"(r[2]), "f"(r[3]), "r"(a) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t a, const float* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31};"
        :: "f"(r[0]), "f"(r[1]), "f"(r[2]), "f"(r[3]), "f"(r[4]), "f"(r[5]), "f"(r[6]), "f"(r[7]), "f"(r[8]), "f"(r[9]),
           "f"(r[10]), "f"(r[11]), "f"(r[12]), "f"(r[13]), "f"(r[14]), "f"(r[15]), "f"(r[16]), "f"(r[17]), "f"(r[18]),
           "f"(r[19]), "f"(r[20]), "f"(r[21]), "f"(r[22]), "f"(r[23]), "f"(r[24]), "f"(r[25]), "f"(r[26]), "f"(r[27]),
           "f"(r[28]), "f"(r[29]), "f"(r[30]), "f"(r[31]), "r"(a) : "memory");
}

// one block of the P3 sweep (dW = g . F^T) of a tile held in tensor memory: N columns of this thread's lane + N / 4 quads of g (broadcast from
// shared memory) are requested together, then N / 2 FFMA2 into {d0, d1} / {d2, d3}
template <int N>
__device__ __forceinline__ void tm_p3_block(uint32_t taddr, uint32_t gaddr, f32x2& da, f32x2& db) {
    float f[N];
    float4 g[N / 4];
    if constexpr (N == 32) tmem_ld32(taddr, f);
    else if constexpr (N == 16) tmem_ld16(taddr, f);
    else if constexpr (N == 8) tmem_ld8(taddr, f);
    else tmem_ld4(taddr, f);
#pragma unroll
    for (int i = 0; i < N / 4; ++i) g[i] = lds128_v(gaddr + 16u * (uint32_t)i);
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < N; ++i) asm volatile("" : "+f"(f[i]));          // keeps the uses below the tcgen05.wait::ld
#pragma unroll
    for (int i = 0; i < N / 4; ++i) {
        da = fma2(pk2(g[i].x, g[i].y), pk2(f[4 * i], f[4 * i + 1]), da);
        db = fma2(pk2(g[i].z, g[i].w), pk2(f[4 * i + 2], f[4 * i + 3]), db);
    }
}
// poll a tagged word until it carries step number `seq`; returns the payload (0 on abort)
static __device__ __noinline__ float poll_word(const unsigned long long* src, unsigned seq, unsigned* abort_flag) {
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

// p_s = sigmoid(d_s) for the 8 pixels of a cell row; d_s is linear in s (from `left` at s = 0 towards `right` at s = 8).
// While |d| log2(e) <= 60 over the row, e_s = exp(-d_s) comes from the recurrence e_s = e_{s-1} * r (2 MUFU.EX2 per row
// instead of 8); rows with larger logits evaluate every exponent separately, clamped to +-60 (the sigmoid is within 1e-18
// of 0 / 1 beyond that). 1/(1+e) for two pixels from one MUFU.RCP: iq = 1/((1+e_a)(1+e_b)), p_a = iq (1+e_b),
// p_b = iq (1+e_a); the product stays below 2^122.
__device__ __forceinline__ void sigmoid_row(float left, float right, float (&p)[8]) {
    const float nleft = NEG_LOG2E * left, nslope = NEG_LOG2E * (right - left) * 0.125f;
    float e[8];
    if (fmaxf(fabsf(nleft), fabsf(fmaf(7.f, nslope, nleft))) <= 60.f) {
        const float r = fast_ex2(nslope), r2 = r * r;
        e[0] = fast_ex2(nleft);
        e[1] = e[0] * r;
#pragma unroll
        for (int s = 2; s < 8; ++s) e[s] = e[s - 2] * r2;          // two independent chains
    } else {
#pragma unroll
        for (int s = 0; s < 8; ++s) e[s] = fast_ex2(fminf(fmaxf(fmaf((float)s, nslope, nleft), -60.f), 60.f));
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float a = 1.f + e[2 * k], b = 1.f + e[2 * k + 1];
        const float iq = fast_rcp(a * b);
        p[2 * k] = iq * b;
        p[2 * k + 1] = iq * a;
    }
}
// one row (8 pixels) of a cell: gradient mass sent to the left / right low-res column of that row.
// General form: any mix of labels (code 0: g = c0 p; code 1: g = c1 (p - 1); ignored: 0).
__device__ __forceinline__ void hires_row(float left, float right, uint32_t rb, float c0, float c1,
                                          float& gl_out, float& gr_out) {
    float p[8];
    sigmoid_row(left, right, p);
    float gs = 0.f, gr = 0.f;
#pragma unroll
    for (int s = 0; s < 8; ++s) {
        const uint32_t sel = rb & (3u << (2 * s)), one = 1u << (2 * s);
        const uint32_t bB = (sel == one) ? __float_as_uint(c1) : 0u;
        const uint32_t bA = (sel == 0u) ? __float_as_uint(c0) : bB;
        const float g = fmaf(__uint_as_float(bA), p[s], -__uint_as_float(bB));
        gs += g;
        gr = fmaf((float)s, g, gr);
    }
    gr *= 0.125f;
    gl_out = gs - gr;
    gr_out = gr;
}
// the same for a row whose 8 pixels carry one label y (the common case away from object boundaries):
// sum_s g = c_y (sum p - 8 y), sum_s s g = c_y (sum s p - 28 y)
__device__ __forceinline__ void hires_row_uniform(float left, float right, bool y, float c0, float c1,
                                                  float& gl_out, float& gr_out) {
    float p[8];
    sigmoid_row(left, right, p);
    const float sp = ((p[0] + p[1]) + (p[2] + p[3])) + ((p[4] + p[5]) + (p[6] + p[7]));
    const float ssp = (fmaf(2.f, p[2], p[1]) + fmaf(4.f, p[4], 3.f * p[3])) + (fmaf(6.f, p[6], 5.f * p[5]) + 7.f * p[7]);
    const float c = y ? c1 : c0;
    const float gs = c * (sp - (y ? 8.f : 0.f));
    const float gr = c * (ssp - (y ? 28.f : 0.f)) * 0.125f;
    gl_out = gs - gr;
    gr_out = gr;
}

// fixed-point unit of an episode: |dW_c| <= max|F| per step (l1 norm of the logit gradient <= 1), cumulative over T
// steps, x4 head-room, rounded up to a power of two; vb = magnitude bits of the value field
__device__ __forceinline__ void fixed_point_unit(unsigned fmax_bits, int T, int KB, float& unit, float& inv) {
    int ex = 0;
    (void)frexpf(4.f * (float)T * (fmax_bits < 0x7f800000u ? __uint_as_float(fmax_bits) : 1.f), &ex);
    ex = min(max(ex, -40), 100);
    const int vb = 63 - KB;
    unit = ldexpf(1.f, ex - vb);
    inv = ldexpf(1.f, vb - ex);
}
// accumulator words of episode e (= group + slot * G): channel c of a step with parity q at sums[acc_offset + q * C + c]
__device__ __forceinline__ size_t acc_offset(int e, int group, int G, int SPL, int C) {
    const int slot = e / G;
    return ((size_t)group * SPL + slot) * 2 * C;
}

}  // namespace cwt
