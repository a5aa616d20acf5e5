// Zero-compressed feature transport: device-side expansion.
//
// Post-ReLU backbone features are ~half zeros, and an evaluation sweep whose episodes live in host memory
// is bound by the host -> device copy (15.2 MB per 1-shot episode; 8 GPUs share ~180 GB/s of host bandwidth).
// The host format therefore ships, per feature tensor, a bit mask (1 bit per fp32 element: set = the element's BIT PATTERN is
// non-zero, so -0.0 / NaN survive), the count of set bits before every block of 32 mask words, and the packed non-zero
// values; this kernel rebuilds the dense NCHW fp32 tensor at HBM speed — bit-identical, so parity is untouched.
//
//   out[32 w + l] = bit l of mask[w] ? vals[prefix(w) + popc(mask[w] & ((1 << l) - 1))] : 0
//
// The tensor is a matrix of n_rows rows (episodes) x W mask words; the prefix count is stored once per BLOCK of 32 words
// (1024 elements: 0.1 % of the dense bytes) and completed inside the warp with a shuffle scan of the words' popcounts.
// A warp expands one block per round (coalesced 128-byte stores; the value loads of a word are consecutive).
#include "common.cuh"

namespace cwt {

__global__ void __launch_bounds__(256)
k_expand_zero_compressed(const uint32_t* __restrict__ mask, const uint32_t* __restrict__ boff, const float* __restrict__ vals,
                         float* __restrict__ out, int n_rows, int W, unsigned base) {
    const int lane = threadIdx.x & 31;
    const int bpr = (W + 31) >> 5;                                     // blocks per row
    const long long n_blocks = (long long)n_rows * bpr;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long b = warp; b < n_blocks; b += n_warps) {
        const int row = (int)(b / bpr), blk = (int)(b - (long long)row * bpr);
        const int w0 = blk * 32, nw = min(32, W - w0);
        const size_t wbase = (size_t)row * W + w0;
        const uint32_t m = lane < nw ? mask[wbase + lane] : 0u;
        // exclusive prefix of the popcounts inside the block
        const unsigned cnt = __popc(m);
        unsigned inc = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += t;
        }
        const unsigned o = boff[b] - base + (inc - cnt);
        // a lane covers 4 consecutive elements (one 128-bit store), 8 lanes cover a word, 4 words per round
#pragma unroll 4
        for (int k = 0; k < nw; k += 4) {
            const int wsel = k + (lane >> 3);                         // this lane's word within the block
            const uint32_t mk = __shfl_sync(0xffffffffu, m, wsel);
            const uint32_t ok = __shfl_sync(0xffffffffu, o, wsel);
            const int b0 = (lane & 7) * 4;
            unsigned r = ok + __popc(mk & ((1u << b0) - 1u));
            float4 v;
            v.x = (mk >> b0) & 1u ? vals[r++] : 0.f;
            v.y = (mk >> (b0 + 1)) & 1u ? vals[r++] : 0.f;
            v.z = (mk >> (b0 + 2)) & 1u ? vals[r++] : 0.f;
            v.w = (mk >> (b0 + 3)) & 1u ? vals[r++] : 0.f;
            if (wsel < nw) *reinterpret_cast<float4*>(out + (wbase + wsel) * 32 + b0) = v;
        }
    }
}

}  // namespace cwt

extern "C" int cwt_expand_zero_compressed_f32(const uint32_t* mask, const uint32_t* block_offsets, const float* vals,
                                              float* out, int n_rows, int words_per_row, unsigned base_offset, void* stream) {
    using namespace cwt;
    CWT_REQUIRE(n_rows >= 0 && words_per_row >= 0, CWT_ERR_INVALID_ARG, "expand: negative size");
    if (n_rows == 0 || words_per_row == 0) return CWT_OK;
    CWT_REQUIRE(mask && block_offsets && out, CWT_ERR_INVALID_ARG, "expand: null pointer");
    CWT_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15u) == 0, CWT_ERR_INVALID_ARG, "expand: out must be 16-byte aligned");
    const long long n_blocks = (long long)n_rows * ((words_per_row + 31) / 32);
    long long ctas = (n_blocks + 7) / 8;                               // one warp per block of 32 words, 8 warps per CTA
    if (ctas > 148 * 16) ctas = 148 * 16;
    k_expand_zero_compressed<<<(unsigned)ctas, 256, 0, static_cast<cudaStream_t>(stream)>>>(mask, block_offsets, vals, out, n_rows,
                                                                                             words_per_row, base_offset);
    CWT_LAUNCHED("expand_zero_compressed");
    return CWT_OK;
}
