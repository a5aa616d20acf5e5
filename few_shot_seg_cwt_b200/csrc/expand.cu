// Zero-compressed feature transport: device-side expansion.
//
// Post-ReLU backbone features are ~half zeros, and an evaluation sweep whose episodes live in host memory
// is bound by the host -> device copy (15.2 MB per 1-shot episode; 8 GPUs share ~180 GB/s of host bandwidth).
// The host format therefore ships, per feature tensor, a bit mask (1 bit per fp32 element: set = the element's BIT PATTERN is
// non-zero, so -0.0 / NaN survive), the exclusive prefix count of set bits per 32-element word, and the packed non-zero
// values; this kernel rebuilds the dense NCHW fp32 tensor at HBM speed — bit-identical, so parity is untouched.
//
//   out[32 w + l] = bit l of mask[w] ? vals[woff[w] - base + popc(mask[w] & ((1 << l) - 1))] : 0
//
// A warp expands 32 consecutive words per round (coalesced 128-byte stores; the value loads of a word are consecutive).
#include "common.cuh"

namespace cwt {

__global__ void __launch_bounds__(256)
k_expand_zero_compressed(const uint32_t* __restrict__ mask, const uint32_t* __restrict__ woff, const float* __restrict__ vals,
                         float* __restrict__ out, long long n_words, unsigned base) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long w0 = warp * 32; w0 < n_words; w0 += n_warps * 32) {
        const long long wi = w0 + lane;
        const uint32_t m = wi < n_words ? mask[wi] : 0u;
        const uint32_t o = wi < n_words ? woff[wi] - base : 0u;
        const int nw = (int)min(32ll, n_words - w0);
        // two rounds of 16 words: a lane covers 4 consecutive elements (one 128-bit store), 8 lanes cover a word
#pragma unroll 4
        for (int k = 0; k < nw; k += 4) {
            const int wsel = k + (lane >> 3);                         // this lane's word within the 32
            const uint32_t mk = __shfl_sync(0xffffffffu, m, wsel);
            const uint32_t ok = __shfl_sync(0xffffffffu, o, wsel);
            const int b0 = (lane & 7) * 4;
            unsigned r = ok + __popc(mk & ((1u << b0) - 1u));
            float4 v;
            v.x = (mk >> b0) & 1u ? vals[r++] : 0.f;
            v.y = (mk >> (b0 + 1)) & 1u ? vals[r++] : 0.f;
            v.z = (mk >> (b0 + 2)) & 1u ? vals[r++] : 0.f;
            v.w = (mk >> (b0 + 3)) & 1u ? vals[r++] : 0.f;
            if (wsel < nw) *reinterpret_cast<float4*>(out + (w0 + wsel) * 32 + b0) = v;
        }
    }
}

}  // namespace cwt

extern "C" int cwt_expand_zero_compressed_f32(const uint32_t* mask, const uint32_t* word_offsets, const float* vals,
                                              float* out, long long n_words, unsigned base_offset, void* stream) {
    using namespace cwt;
    CWT_REQUIRE(n_words >= 0, CWT_ERR_INVALID_ARG, "expand: negative size");
    if (n_words == 0) return CWT_OK;
    CWT_REQUIRE(mask && word_offsets && out, CWT_ERR_INVALID_ARG, "expand: null pointer");
    CWT_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15u) == 0, CWT_ERR_INVALID_ARG, "expand: out must be 16-byte aligned");
    long long blocks = (n_words + 255) / 256;                          // one warp per 32 words, 8 warps per CTA
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_expand_zero_compressed<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(mask, word_offsets, vals, out,
                                                                                               n_words, base_offset);
    CWT_LAUNCHED("expand_zero_compressed");
    return CWT_OK;
}
