// (f-4) The validation transform of the reference on the device: Resize (aspect-preserving cv2.resize to a multiple-of-8
// side, zero / mean padding to size x size; label: nearest, padded with 255) -> ToTensor (/255) -> Normalize
// (src/dataset/transform.py:109-163, :58-107; composed in src/dataset/dataset.py:78-84). One fused kernel per image:
// HWC float32 RGB [0,255] in, normalised CHW float32 + label out — the three host passes (cv2.resize, the float64
// pad buffer, the tensor arithmetic) become one pass over the output.
//
// cv2.resize semantics restated (OpenCV imgproc/resize.cpp; third-party, not vendored by the reference):
//   INTER_LINEAR, float32 source:  fx = (float)((dx + 0.5) * scale - 0.5), scale = 1 / ((double)dst / src);
//       sx = floor(fx); fx -= sx; sx < 0 -> (0, fx = 0); sx >= src-1 -> (src-1, fx = 0); horizontal pass first
//       (S[sx] (1-fx) + S[sx+1] fx on both rows), then vertical — float arithmetic, no fixed point for float sources
//   INTER_NEAREST:  sx = min(floor(dx * scale), src - 1)
#include "common.cuh"

namespace cwt {

struct TransformP {
    const float* img; const uint8_t* lab; float* out; void* lab_out;
    int oh, ow, nh, nw, size, lab_i64, pad_label;
    double sy, sx;                      // source / destination scale factors (image), as OpenCV computes them
    float mean[3], stdv[3], pad[3];
};

__device__ __forceinline__ void lin_coord(int d, double scale, int src, int& s, float& f) {
    f = (float)(((double)d + 0.5) * scale - 0.5);
    s = (int)floorf(f);
    f -= (float)s;
    if (s < 0) { s = 0; f = 0.f; }
    if (s >= src - 1) { s = src - 1; f = 0.f; }
}

__global__ void __launch_bounds__(256)
k_resize_pad_normalize(TransformP p) {
    const int X = blockIdx.x * blockDim.x + threadIdx.x, Y = blockIdx.y;
    if (X >= p.size) return;
    const bool inside = (Y < p.nh && X < p.nw);
    float v[3] = {p.pad[0], p.pad[1], p.pad[2]};
    long long lv = p.pad_label;
    if (inside) {
        int sy, sx; float fy, fx;
        lin_coord(Y, p.sy, p.oh, sy, fy);
        lin_coord(X, p.sx, p.ow, sx, fx);
        const int sy1 = min(sy + 1, p.oh - 1), sx1 = min(sx + 1, p.ow - 1);
        const float a0 = 1.f - fx, a1 = fx, b0 = 1.f - fy, b1 = fy;
        const float* r0 = p.img + ((size_t)sy * p.ow) * 3;
        const float* r1 = p.img + ((size_t)sy1 * p.ow) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float h0 = __fadd_rn(__fmul_rn(r0[sx * 3 + c], a0), __fmul_rn(r0[sx1 * 3 + c], a1));
            const float h1 = __fadd_rn(__fmul_rn(r1[sx * 3 + c], a0), __fmul_rn(r1[sx1 * 3 + c], a1));
            v[c] = __fadd_rn(__fmul_rn(h0, b0), __fmul_rn(h1, b1));
        }
        if (p.lab) {
            const int ly = min((int)floor((double)Y * p.sy), p.oh - 1), lx = min((int)floor((double)X * p.sx), p.ow - 1);
            lv = p.lab[(size_t)ly * p.ow + lx];
        }
    }
    const size_t plane = (size_t)p.size * p.size, o = (size_t)Y * p.size + X;
#pragma unroll
    for (int c = 0; c < 3; ++c)                                        // ToTensor: /255 ; Normalize: (x - mean) / std
        p.out[c * plane + o] = __fdiv_rn(__fsub_rn(__fdiv_rn(v[c], 255.f), p.mean[c]), p.stdv[c]);
    if (p.lab_out) {
        if (p.lab_i64) reinterpret_cast<long long*>(p.lab_out)[o] = lv;
        else reinterpret_cast<uint8_t*>(p.lab_out)[o] = (uint8_t)lv;
    }
}

}  // namespace cwt

using namespace cwt;

extern "C" int cwt_resize_pad_normalize_f32(const float* image_hwc, const uint8_t* label_or_null, int ori_h, int ori_w,
                                            int new_h, int new_w, int size, const float* mean3, const float* std3,
                                            const float* pad3_or_null, int pad_label, float* out_chw, void* label_out_or_null,
                                            int label_out_kind, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CWT_REQUIRE(image_hwc && out_chw && mean3 && std3, CWT_ERR_INVALID_ARG, "resize_pad_normalize: null pointer");
    CWT_REQUIRE(ori_h >= 1 && ori_w >= 1 && new_h >= 1 && new_w >= 1 && new_h <= size && new_w <= size, CWT_ERR_INVALID_ARG,
                "resize_pad_normalize: bad sizes %dx%d -> %dx%d in %d", ori_h, ori_w, new_h, new_w, size);
    CWT_REQUIRE(label_out_kind == CWT_LABEL_U8 || label_out_kind == CWT_LABEL_I64, CWT_ERR_INVALID_ARG,
                "resize_pad_normalize: label_out_kind %d", label_out_kind);
    TransformP p;
    p.img = image_hwc; p.lab = label_or_null; p.out = out_chw; p.lab_out = label_or_null ? label_out_or_null : nullptr;
    p.oh = ori_h; p.ow = ori_w; p.nh = new_h; p.nw = new_w; p.size = size; p.lab_i64 = (label_out_kind == CWT_LABEL_I64);
    p.pad_label = pad_label;
    p.sy = 1.0 / ((double)new_h / (double)ori_h);                    // OpenCV: scale = 1 / inv_scale, inv_scale = dst / src
    p.sx = 1.0 / ((double)new_w / (double)ori_w);
    for (int c = 0; c < 3; ++c) { p.mean[c] = mean3[c]; p.stdv[c] = std3[c]; p.pad[c] = pad3_or_null ? pad3_or_null[c] : 0.f; }
    dim3 grid((size + 255) / 256, size);
    k_resize_pad_normalize<<<grid, 256, 0, st>>>(p);
    CWT_LAUNCHED("resize_pad_normalize");
    return CWT_OK;
}
