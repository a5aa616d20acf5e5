// CWT_ATTN_TCGEN05 — the K/V projection of MultiHeadAttentionOne as a tcgen05 / TMEM GEMM.
//
// Reference: k = w_qkvs(f_q^T) with f_q [HW, C] and w_qkvs.weight A [nH*C, C]  (src/model/transformer.py:68-69),
// then attn = q k^T / sqrt(C) (:24-25). K and V are the same tensor (k is v, one shared weight), so the
// HW x C x C GEMM is done ONCE per head (the reference does it twice), on the 5th-generation tensor cores:
//
//   pre-pass     X[e][p][c] = f_q[e][c][p] transposed to K-major and split  x = hi + lo  (2 x bf16), |x_p|^2;
//                A split the same way. 3 products  hi*hi + hi*lo + lo*hi  with fp32 accumulation in TMEM give
//                ~2^-17 relative error per term — the 1e-4 parity budget cannot be met by one bf16 pass
//                (5e-4) nor safely by one TF32 pass (6-8e-5).
//   main kernel  CTA = (128-pixel tile, head, episode). TMA (cp.async.bulk.tensor, SWIZZLE_128B) feeds a
//                2-stage shared-memory ring; one elected thread issues tcgen05.mma (M=128, N<=256, K=16,
//                kind::f16, cta_group::1); the 128 x C fp32 tile of K lives only in TMEM (C <= 512 columns).
//   epilogue     4 warps read their TMEM lane quadrant with tcgen05.ld and contract each pixel row with the
//                Lq projected query rows Q_h: scores[e][l*nH+h][p] = K[p,:] . Q_h[l,:]  — K never goes to HBM.
//
// Softmax, the V-side (re-associated: (P X) A_h^T), fc, residual and LayerNorm follow in transformer.cu.
// Arithmetically this path does 3 * 2*HW*C*C flops per head where the re-associated path (CWT_ATTN_REASSOC)
// needs 2*Lq*HW*C — it exists because the north-star contract names the GEMM; bench.py reports both.
#include "common.cuh"
#include <cuda.h>
#include <cuda_bf16.h>

namespace cwt {

constexpr int KP_BM = 128;        // pixels per tile (UMMA M)
constexpr int KP_BK = 64;         // channels per k-block = one 128-byte swizzle row of bf16
constexpr int KP_STAGES = 2;
constexpr int KP_THREADS = 128;

__host__ __device__ static inline int kp_hwpad(int HW) { return (HW + KP_BM - 1) / KP_BM * KP_BM; }
__host__ __device__ static inline int kp_npass(int C) { return (C + 255) / 256; }

// ---- pre-pass: transpose + bf16 hi/lo split + squared norms -------------------------------------------------
// grid (HWpad/32, E), block 256: 32 pixels x all channels, 64-channel chunks through shared memory.
// read: 128-byte pixel rows per channel (coalesced); write: thread = (pixel, 8 channels) -> one 16-byte store of hi
// and one of lo, 8 adjacent lanes cover 128 contiguous bytes of the K-major row.
__global__ void __launch_bounds__(256)
k_split_transpose(const float* __restrict__ k, __nv_bfloat16* __restrict__ Xh, __nv_bfloat16* __restrict__ Xl,
                  float* __restrict__ n2, int C, int HW, int HWpad) {
    __shared__ float tile[64][33];
    const int e = blockIdx.y, p0 = blockIdx.x * 32, tid = threadIdx.x;
    const int rc = tid >> 5, rp = tid & 31;                // read phase: 8 channel rows x 32 pixels per pass
    const int wp = tid >> 3, wg = tid & 7;                 // write phase: pixel, group of 8 channels
    float nsum = 0.f;
    for (int c0 = 0; c0 < C; c0 += 64) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = rc + 8 * i;
            float v = 0.f;
            if (c0 + c < C && p0 + rp < HW) v = k[((size_t)e * C + c0 + c) * HW + p0 + rp];
            tile[c][rp] = v;
        }
        __syncthreads();
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float a = tile[wg * 8 + 2 * j][wp], b = tile[wg * 8 + 2 * j + 1][wp];
            const __nv_bfloat16 ah = __float2bfloat16_rn(a), bh = __float2bfloat16_rn(b);
            const __nv_bfloat16 al = __float2bfloat16_rn(a - __bfloat162float(ah));
            const __nv_bfloat16 bl = __float2bfloat16_rn(b - __bfloat162float(bh));
            hi[j] = (uint32_t)__bfloat16_as_ushort(ah) | ((uint32_t)__bfloat16_as_ushort(bh) << 16);
            lo[j] = (uint32_t)__bfloat16_as_ushort(al) | ((uint32_t)__bfloat16_as_ushort(bl) << 16);
            nsum = fmaf(a, a, fmaf(b, b, nsum));
        }
        if (c0 + wg * 8 < C) {
            const size_t row = ((size_t)e * HWpad + p0 + wp) * C + c0 + wg * 8;
            *reinterpret_cast<uint4*>(Xh + row) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(Xl + row) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
        __syncthreads();
    }
    nsum += __shfl_xor_sync(0xffffffffu, nsum, 1);
    nsum += __shfl_xor_sync(0xffffffffu, nsum, 2);
    nsum += __shfl_xor_sync(0xffffffffu, nsum, 4);
    if (n2 && wg == 0 && p0 + wp < HW) n2[(size_t)e * HW + p0 + wp] = nsum;
}

__global__ void __launch_bounds__(256)
k_split_weights(const float* __restrict__ A, __nv_bfloat16* __restrict__ Ah, __nv_bfloat16* __restrict__ Al, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float a = A[i];
    const __nv_bfloat16 h = __float2bfloat16_rn(a);
    Ah[i] = h;
    Al[i] = __float2bfloat16_rn(a - __bfloat162float(h));
}

// ---- PTX helpers ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t kp_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void kp_mbar_init(uint64_t* b, unsigned n) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(kp_smem(b)), "r"(n) : "memory");
}
__device__ __forceinline__ void kp_mbar_expect_tx(uint64_t* b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(kp_smem(b)), "r"(bytes) : "memory");
}
// bounded wait: a wrong descriptor / tensor map must fail loudly (trap) instead of hanging the GPU
__device__ __forceinline__ void kp_mbar_wait(uint64_t* b, unsigned parity) {
    const long long t0 = clock64();
    for (;;) {
        unsigned ok;
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(kp_smem(b)), "r"(parity) : "memory");
        if (ok) return;
        if (clock64() - t0 > 4000000000ll) __trap();      // ~2 s
    }
}
__device__ __forceinline__ void kp_tma_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(kp_smem(dst)), "l"(map), "r"(kp_smem(bar)), "r"(c0), "r"(c1) : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): rows of 128 B, 8-row atoms
// 1024 B apart (SBO), LBO unused (1), version 1 (Blackwell), layout_type 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t kp_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fffu);             // start address      bits [0,14)
    d |= (uint64_t)1u << 16;                             // leading byte offset bits [16,30)  (ignored for swizzled K-major)
    d |= (uint64_t)(1024u >> 4) << 32;                   // stride byte offset  bits [32,46)
    d |= (uint64_t)1u << 46;                             // version             bits [46,48)
    d |= (uint64_t)2u << 61;                             // layout type         bits [61,64)
    return d;
}
// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): fp32 accumulate, bf16 x bf16, both K-major
__device__ __forceinline__ uint32_t kp_instr_desc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void kp_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}
__device__ __forceinline__ void kp_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(kp_smem(bar)) : "memory");
}
__device__ __forceinline__ void kp_tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct KprojParams {
    const float* Qp;      // [E*Lq][nH*C] projected queries
    float* sraw;          // [E][Lq*nH][HW]
    int E, Lq, nH, C, HW, HWpad;
};

// dynamic smem: [stage][ Xh 16K | Xl 16K | Ah Nper*128 | Al Nper*128 ] (1024-aligned) | Qs [Lq][C] | barriers | tmem slot
__global__ void __launch_bounds__(KP_THREADS, 1)
k_kproj_scores(const __grid_constant__ CUtensorMap mXh, const __grid_constant__ CUtensorMap mXl,
               const __grid_constant__ CUtensorMap mAh, const __grid_constant__ CUtensorMap mAl, KprojParams p) {
    extern __shared__ __align__(1024) unsigned char kp_smem_raw[];
    const int C = p.C, npass = kp_npass(C), Nper = C / npass, nkb = C / KP_BK;
    const unsigned x_bytes = KP_BM * KP_BK * 2, a_bytes = (unsigned)Nper * KP_BK * 2;
    const unsigned stage_bytes = 2 * x_bytes + 2 * a_bytes;
    // SWIZZLE_128B atoms need 1024-byte aligned tiles (the launch reserves 1 KB of slack for this)
    unsigned char* stages = kp_smem_raw + ((1024u - (kp_smem(kp_smem_raw) & 1023u)) & 1023u);
    float* Qs = reinterpret_cast<float*>(stages + KP_STAGES * stage_bytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(Qs + (size_t)p.Lq * C);
    uint64_t* full = bars;                  // [KP_STAGES] TMA -> MMA
    uint64_t* empty = bars + KP_STAGES;     // [KP_STAGES] MMA -> TMA
    uint64_t* done = bars + 2 * KP_STAGES;  // accumulator complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * KP_STAGES + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile = blockIdx.x, hh = blockIdx.y, e = blockIdx.z;

    if (tid == 0) {
        for (int s = 0; s < KP_STAGES; ++s) { kp_mbar_init(&full[s], 1); kp_mbar_init(&empty[s], 1); }
        kp_mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {      // one warp allocates all 512 TMEM columns (the tile needs C <= 512 fp32 columns)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(kp_smem(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // the projected query rows of this (episode, head)
    for (int i = tid; i < p.Lq * C; i += KP_THREADS) {
        const int l = i / C, n = i - l * C;
        Qs[i] = p.Qp[((size_t)e * p.Lq + l) * p.nH * C + (size_t)hh * C + n];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    const int n_iter = npass * nkb;
    if (warp == 0 && lane == 0) {
        // ===== TMA producer =====
        for (int it = 0; it < n_iter; ++it) {
            const int s = it % KP_STAGES, round = it / KP_STAGES;
            const int pass = it / nkb, kb = it - pass * nkb;
            if (round > 0) kp_mbar_wait(&empty[s], (round - 1) & 1);
            unsigned char* st = stages + (size_t)s * stage_bytes;
            kp_mbar_expect_tx(&full[s], stage_bytes);
            const int xrow = e * p.HWpad + tile * KP_BM, arow = hh * C + pass * Nper;
            kp_tma_2d(st, &mXh, kb * KP_BK, xrow, &full[s]);
            kp_tma_2d(st + x_bytes, &mXl, kb * KP_BK, xrow, &full[s]);
            kp_tma_2d(st + 2 * x_bytes, &mAh, kb * KP_BK, arow, &full[s]);
            kp_tma_2d(st + 2 * x_bytes + a_bytes, &mAl, kb * KP_BK, arow, &full[s]);
        }
    } else if (warp == 1 && lane == 0) {
        // ===== MMA issuer (one thread issues for the whole CTA) =====
        const uint32_t idesc = kp_instr_desc(KP_BM, Nper);
        for (int it = 0; it < n_iter; ++it) {
            const int s = it % KP_STAGES, round = it / KP_STAGES;
            const int pass = it / nkb, kb = it - pass * nkb;
            kp_mbar_wait(&full[s], round & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t st = kp_smem(stages + (size_t)s * stage_bytes);
            const uint32_t d = tmem_base + (uint32_t)(pass * Nper);
#pragma unroll
            for (int ks = 0; ks < KP_BK / 16; ++ks) {
                // advancing K by 16 bf16 = 32 bytes inside the 128-byte swizzle row: +2 in the (>>4) address field
                const uint64_t xh = kp_smem_desc(st + ks * 32), xl = kp_smem_desc(st + x_bytes + ks * 32);
                const uint64_t ah = kp_smem_desc(st + 2 * x_bytes + ks * 32), al = kp_smem_desc(st + 2 * x_bytes + a_bytes + ks * 32);
                kp_mma(d, xh, ah, idesc, (kb | ks) != 0);     // hi * hi  (first MMA of a pass overwrites)
                kp_mma(d, xh, al, idesc, true);               // hi * lo
                kp_mma(d, xl, ah, idesc, true);               // lo * hi
            }
            kp_commit(&empty[s]);                             // frees the smem stage when these MMAs have read it
        }
        kp_commit(done);                                      // accumulator tile complete
    }
    __syncwarp();

    // ===== epilogue: all 4 warps, warp w owns TMEM lanes [32w, 32w+32) = pixels of the tile =====
    kp_mbar_wait(done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int cb = 0; cb < C; cb += 32) {
        uint32_t v[32];
        kp_tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)cb, v);
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            if (l < p.Lq) {
                const float4* q4 = reinterpret_cast<const float4*>(Qs + (size_t)l * C + cb);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 q = q4[j];
                    acc[l] = fmaf(__uint_as_float(v[4 * j]), q.x, acc[l]);
                    acc[l] = fmaf(__uint_as_float(v[4 * j + 1]), q.y, acc[l]);
                    acc[l] = fmaf(__uint_as_float(v[4 * j + 2]), q.z, acc[l]);
                    acc[l] = fmaf(__uint_as_float(v[4 * j + 3]), q.w, acc[l]);
                }
            }
        }
    }
    const int pix = tile * KP_BM + warp * 32 + lane;
    if (pix < p.HW) {
#pragma unroll
        for (int l = 0; l < 4; ++l)
            if (l < p.Lq) p.sraw[((size_t)e * (p.Lq * p.nH) + l * p.nH + hh) * p.HW + pix] = acc[l];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

// ---- host ----------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// bf16 matrix [rows][cols] row-major, box [box_rows][64 cols], 128-byte swizzle
static int make_map(CUtensorMap* m, void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    CWT_REQUIRE(fn, CWT_ERR_CUDA, "kproj_tcgen05: cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * 2};
    cuuint32_t box[2] = {(cuuint32_t)KP_BK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CWT_REQUIRE(r == CUDA_SUCCESS, CWT_ERR_CUDA, "kproj_tcgen05: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return CWT_OK;
}

size_t kproj_tcgen05_workspace_bytes(int E, int Lq, int nH, int C, int HW) {
    (void)Lq;
    const size_t x = align_up((size_t)E * kp_hwpad(HW) * C * 2);
    const size_t a = align_up((size_t)nH * C * C * 2);
    return 2 * x + 2 * a + 1024;
}

// sraw[e][l*nH+h][p] = Q_h[e,l,:] . (A_h x_p)   (un-normalised, un-scaled scores);  n2[e][p] = |x_p|^2 (nullable)
int kproj_scores_tcgen05(const float* k, const float* w_qkvs, const float* Qp, float* sraw, float* n2,
                         int E, int Lq, int nH, int C, int HW, void* ws, size_t ws_bytes, cudaStream_t st) {
    CWT_REQUIRE(C % 64 == 0 && C <= 512, CWT_ERR_UNSUPPORTED, "CWT_ATTN_TCGEN05 needs C %% 64 == 0 and C <= 512 (C=%d)", C);
    CWT_REQUIRE(Lq <= 4, CWT_ERR_UNSUPPORTED, "CWT_ATTN_TCGEN05 supports Lq <= 4 (Lq=%d)", Lq);
    const int npass = kp_npass(C), Nper = C / npass;
    CWT_REQUIRE(Nper % 16 == 0, CWT_ERR_UNSUPPORTED, "CWT_ATTN_TCGEN05: C/%d must be a multiple of 16", npass);
    const int HWpad = kp_hwpad(HW);
    Carver cv(ws, ws_bytes);
    __nv_bfloat16* Xh = cv.take<__nv_bfloat16>((size_t)E * HWpad * C);
    __nv_bfloat16* Xl = cv.take<__nv_bfloat16>((size_t)E * HWpad * C);
    __nv_bfloat16* Ah = cv.take<__nv_bfloat16>((size_t)nH * C * C);
    __nv_bfloat16* Al = cv.take<__nv_bfloat16>((size_t)nH * C * C);
    CWT_REQUIRE(ws && cv.ok(), CWT_ERR_WORKSPACE, "kproj_tcgen05: workspace too small");

    k_split_transpose<<<dim3(HWpad / 32, E), 256, 0, st>>>(k, Xh, Xl, n2, C, HW, HWpad);
    CWT_LAUNCHED("split_transpose");
    const size_t na = (size_t)nH * C * C;
    k_split_weights<<<(unsigned)((na + 255) / 256), 256, 0, st>>>(w_qkvs, Ah, Al, na);
    CWT_LAUNCHED("split_weights");

    CUtensorMap mXh, mXl, mAh, mAl;
    int rc;
    if ((rc = make_map(&mXh, Xh, (uint64_t)E * HWpad, C, KP_BM))) return rc;
    if ((rc = make_map(&mXl, Xl, (uint64_t)E * HWpad, C, KP_BM))) return rc;
    if ((rc = make_map(&mAh, Ah, (uint64_t)nH * C, C, Nper))) return rc;
    if ((rc = make_map(&mAl, Al, (uint64_t)nH * C, C, Nper))) return rc;

    KprojParams p{Qp, sraw, E, Lq, nH, C, HW, HWpad};
    const size_t stage_bytes = 2 * (size_t)KP_BM * KP_BK * 2 + 2 * (size_t)Nper * KP_BK * 2;
    const size_t smem = KP_STAGES * stage_bytes + (size_t)Lq * C * 4 + (2 * KP_STAGES + 1) * 8 + 16 + 1024;
    CWT_CUDA(cudaFuncSetAttribute(k_kproj_scores, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(HWpad / KP_BM, nH, E);
    k_kproj_scores<<<grid, KP_THREADS, smem, st>>>(mXh, mXl, mAh, mAl, p);
    CWT_LAUNCHED("kproj_scores_tcgen05");
    return CWT_OK;
}

}  // namespace cwt
