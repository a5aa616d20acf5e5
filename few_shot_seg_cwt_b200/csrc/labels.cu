// (a-2) label statistics on the device: replaces the per-episode D2H copy + numpy `where`
// of the reference (src/test.py:169-171, src/train.py:211-213,237-239, model_util.py:27-31).
#include "common.cuh"

namespace cwt {

template <bool I64>
__global__ void __launch_bounds__(256)
k_prep_labels(const void* __restrict__ labels, long long npix, int ignore_index,
              uint8_t* __restrict__ packed, int32_t* __restrict__ counts) {
    const int img = blockIdx.y;
    const size_t base = (size_t)img * (size_t)npix;
    int c[4] = {0, 0, 0, 0};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npix;
         i += (long long)gridDim.x * blockDim.x) {
        int code = load_label_code<I64>(labels, base + (size_t)i, ignore_index);
        if (packed) packed[base + (size_t)i] = (uint8_t)code;
        c[0] += (code == 0); c[1] += (code == 1); c[2] += (code == 2); c[3] += (code == 3);
    }
    __shared__ int sm[4];
    if (threadIdx.x < 4) sm[threadIdx.x] = 0;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int v = __reduce_add_sync(0xffffffffu, c[k]);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&sm[k], v);
    }
    __syncthreads();
    if (threadIdx.x < 4 && sm[threadIdx.x]) atomicAdd(&counts[img * 4 + threadIdx.x], sm[threadIdx.x]);
}

int prep_labels(const void* labels, int label_kind, int n_img, long long npix, int ignore_index,
                uint8_t* packed, int32_t* counts, cudaStream_t st) {
    CWT_REQUIRE(labels && counts, CWT_ERR_INVALID_ARG, "prep_labels: null pointer");
    CWT_REQUIRE(label_kind == CWT_LABEL_U8 || label_kind == CWT_LABEL_I64, CWT_ERR_INVALID_ARG,
                "prep_labels: label_kind %d", label_kind);
    CWT_REQUIRE(n_img >= 0 && npix >= 0, CWT_ERR_INVALID_ARG, "prep_labels: negative size");
    if (n_img == 0) return CWT_OK;
    CWT_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * 4 * (size_t)n_img, st));
    if (npix == 0) return CWT_OK;
    int bx = (int)((npix + 256 * 8 - 1) / (256 * 8));
    if (bx < 1) bx = 1;
    if (bx > 64) bx = 64;
    dim3 grid(bx, n_img);
    if (label_kind == CWT_LABEL_I64)
        k_prep_labels<true><<<grid, 256, 0, st>>>(labels, npix, ignore_index, packed, counts);
    else
        k_prep_labels<false><<<grid, 256, 0, st>>>(labels, npix, ignore_index, packed, counts);
    CWT_LAUNCHED("prep_labels");
    return CWT_OK;
}


// Cell packing for the full-resolution stage: the H x W label image is cut into the h x w cells
// of the align_corners grid (cell (a,b) = pixels [8a,8a+8) x [8b,8b+8)); each cell stores its
// 8 rows as 8 x uint16 (2 bits per pixel: 0 / 1 / 2 ignored / 3 invalid), out-of-image pixels
// are "ignored". 16 bytes per cell -> one 128-bit load per cell per SGD step instead of 64 bytes.
template <bool I64>
__global__ void __launch_bounds__(256)
k_pack_label_cells(const void* __restrict__ labels, int h, int w, int H, int W, int ignore_index,
                   uint4* __restrict__ cells, int32_t* __restrict__ counts) {
    const int img = blockIdx.y;
    const size_t base = (size_t)img * H * W;
    int c[4] = {0, 0, 0, 0};
    for (int cell = blockIdx.x * blockDim.x + threadIdx.x; cell < h * w; cell += gridDim.x * blockDim.x) {
        const int a = cell / w, b = cell - a * w;
        // branch-free: all 64 loads are unconditional (coordinates clamped into the image) and issued per row before use;
        // pixels outside the image are coded 2 and masked out of the counts
        const int nx = min(max(W - 8 * b, 0), 8);                       // in-image pixels of this cell's rows
        const uint32_t xmask = ((1u << (2 * nx)) - 1u) & 0x5555u;
        uint32_t rows[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int Y = 8 * a + r;
            const size_t rb = base + (size_t)min(Y, H - 1) * W;
            uint32_t raw[8];
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                const size_t i = rb + min(8 * b + s, W - 1);
                long long v = I64 ? reinterpret_cast<const long long*>(labels)[i] : (long long)reinterpret_cast<const uint8_t*>(labels)[i];
                raw[s] = (v == 0) ? 0u : ((v == 1) ? 1u : ((v == (long long)ignore_index) ? 2u : 3u));
            }
            uint32_t bits = 0;
#pragma unroll
            for (int s = 0; s < 8; ++s) bits |= ((s < nx && Y < H) ? raw[s] : 2u) << (2 * s);
            rows[r] = bits;
            const uint32_t m = (Y < H) ? xmask : 0u, lo = bits & 0x5555u, hi = (bits >> 1) & 0x5555u;
            c[0] += __popc(~lo & ~hi & m); c[1] += __popc(lo & ~hi & m); c[2] += __popc(~lo & hi & m); c[3] += __popc(lo & hi & m);
        }
        cells[(size_t)img * h * w + cell] = make_uint4(rows[0] | (rows[1] << 16), rows[2] | (rows[3] << 16),
                                                       rows[4] | (rows[5] << 16), rows[6] | (rows[7] << 16));
    }
    __shared__ int sm[4];
    if (threadIdx.x < 4) sm[threadIdx.x] = 0;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int v = __reduce_add_sync(0xffffffffu, c[k]);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&sm[k], v);
    }
    __syncthreads();
    if (threadIdx.x < 4 && sm[threadIdx.x]) atomicAdd(&counts[img * 4 + threadIdx.x], sm[threadIdx.x]);
}

int pack_label_cells(const void* labels, int label_kind, int n_img, int h, int w, int H, int W, int ignore_index,
                     uint4* cells, int32_t* counts, cudaStream_t st) {
    CWT_REQUIRE(labels && cells && counts, CWT_ERR_INVALID_ARG, "pack_label_cells: null pointer");
    if (n_img == 0) return CWT_OK;
    CWT_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * 4 * (size_t)n_img, st));
    dim3 grid((h * w + 255) / 256, n_img);
    if (label_kind == CWT_LABEL_I64)
        k_pack_label_cells<true><<<grid, 256, 0, st>>>(labels, h, w, H, W, ignore_index, cells, counts);
    else
        k_pack_label_cells<false><<<grid, 256, 0, st>>>(labels, h, w, H, W, ignore_index, cells, counts);
    CWT_LAUNCHED("pack_label_cells");
    return CWT_OK;
}

}  // namespace cwt

extern "C" int cwt_prep_labels(const void* labels, int label_kind, int n_img, long long npix, int ignore_index,
                               uint8_t* packed_or_null, int32_t* counts, void* stream) {
    return cwt::prep_labels(labels, label_kind, n_img, npix, ignore_index, packed_or_null, counts,
                            static_cast<cudaStream_t>(stream));
}
