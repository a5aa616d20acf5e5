/*
 * cwt_b200.h — C ABI of the B200-native CWT few-shot segmentation head.
 *
 * One shared object (libcwt_b200.so), plain pointers and sizes, no C++/torch types.
 * Every entry point replaces a piece of the reference's Python/ATen hot path
 * (TeamOfProfGuo/Few_Shot_Seg_CWT; paths below are relative to the reference checkout).
 * The reference has no FFI of its own (it is 100 % Python), so these are the calls
 * a maintainer would bind with ctypes — see INTEGRATION.md for the stubs.
 *
 * Conventions
 *   - return 0 (CWT_OK) or a negative error code; nothing throws across the ABI;
 *     cwt_last_error() returns a per-thread message for the last failure;
 *   - the caller owns every buffer (inputs, outputs, workspace); the library never
 *     allocates device memory, never frees and never keeps a pointer past the call;
 *   - all device pointers, all work is enqueued asynchronously on `stream`
 *     (a cudaStream_t passed as void*; 0 = legacy default stream); no hidden syncs;
 *   - re-entrant / thread-safe for distinct streams and distinct workspaces;
 *   - features are contiguous NCHW fp32: f[img][c][y][x]; labels are uint8 or int64
 *     with values {0, 1, ignore_index}; any other value is counted as "invalid"
 *     (torch's CrossEntropyLoss would raise on it) and treated as ignored;
 *   - geometry: the bilinear maps are align_corners=True with H = 8*(h-1)+1 and
 *     W = 8*(w-1)+1 (473 <-> 60, 417 <-> 53: the PSPNet zoom-8 head), which makes the
 *     scale exactly 1/8; other ratios return CWT_ERR_UNSUPPORTED.
 */
#ifndef CWT_B200_H_
#define CWT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CWT_OK                 0
#define CWT_ERR_INVALID_ARG   -1
#define CWT_ERR_UNSUPPORTED   -2
#define CWT_ERR_WORKSPACE     -3
#define CWT_ERR_CUDA          -4

#define CWT_LABEL_U8   0
#define CWT_LABEL_I64  1

/* fit algorithms */
#define CWT_FIT_AUTO       0
#define CWT_FIT_STREAM     1   /* features streamed from HBM/L2 every step (any S)            */
#define CWT_FIT_RESIDENT   2   /* features staged once into shared memory, persistent kernel  */
#define CWT_FIT_L2         3   /* persistent kernel, features streamed from L2 every sweep (any S; C=512, 60x60): the
                                  default for S > 1 — G episodes at a time stay L2-resident for all n_iter steps */

/* transformer score algorithms */
#define CWT_ATTN_REASSOC   0   /* scores = (Q_h A_h) X^T : skinny contraction on CUDA cores   */
#define CWT_ATTN_TCGEN05   1   /* K = X A_h^T as a tcgen05/TMEM GEMM (3xbf16 split), scores in the epilogue */

int         cwt_version(void);
const char* cwt_last_error(void);
/* number of kernels this library has launched from the CALLING THREAD (per-thread like the error string; bench / smoke
 * bookkeeping: "gpu_launches") */
long long   cwt_launch_count(void);
/* developer-only entry points (phase profile of the resident fit, L2 bandwidth probe) live in cwt_b200_debug.h */

/* ---------------------------------------------------------------------------------------
 * (a-2) label statistics.  Replaces the per-episode D2H copy + numpy `where` of
 * src/test.py:169-171 / src/train.py:211-213,237-239 and model_util.py:27-31.
 *   labels  [n_img, npix] uint8 or int64
 *   packed  [n_img, npix] uint8 out (0, 1, 2 = ignored, 3 = invalid) or NULL
 *   counts  [n_img, 4] int32 out: n0, n1, n_ignored, n_invalid  (zeroed by the call)
 * ------------------------------------------------------------------------------------- */
int cwt_prep_labels(const void* labels, int label_kind, int n_img, long long npix, int ignore_index,
                    uint8_t* packed_or_null, int32_t* counts, void* stream);

/* ---------------------------------------------------------------------------------------
 * (a-1..a-3) support-classifier fit.  Replaces the inline loop of src/test.py:164-187
 * (= src/train.py:206-231) and PSPNet.inner_loop, src/model/pspnet.py:189-205:
 * n_iter x { 1x1 conv 2xC, bilinear up to HxW, class-weighted CE (ignore), backward, SGD }.
 *   f_s     [E,S,C,h,w] fp32        s_label [E,S,H,W]
 *   w0      [E,2,C] initial weights (the reference draws them with kaiming-uniform, a-1)
 *   class_weight_or_null [E,2]; NULL => [1, n0/n1] counted on the device over all S shots
 *   w_out   [E,2,C]
 *   loss_trace_or_null [n_iter,E]: the CE value of every step (slower kernels when given)
 *   label_counts_or_null [E,4] int32 out: n0,n1,n_ignored,n_invalid per episode; n1 == 0 is
 *       the reference's ZeroDivisionError — the caller checks it (the Python layer raises)
 * ------------------------------------------------------------------------------------- */
size_t cwt_fit_workspace_bytes(int E, int S, int C, int h, int w, int H, int W);
int cwt_fit_classifier_f32(const float* f_s, const void* s_label, int label_kind, const float* w0,
                           const float* class_weight_or_null, float* w_out,
                           float* loss_trace_or_null, int32_t* label_counts_or_null,
                           int E, int S, int C, int h, int w, int H, int W,
                           int n_iter, float lr, int ignore_index, int algo,
                           void* workspace, size_t ws_bytes, void* stream);

/* Deferred error check of a fit (no host sync): status[e] = OR of
 *   CWT_FIT_BAD_LABEL   a support label outside {0, 1, ignore_index}      (torch's CrossEntropyLoss raises)
 *   CWT_FIT_NO_FG       no foreground pixel and no explicit class weight   (the reference's ZeroDivisionError,
 *                       src/test.py:174: len(back_pix[0]) / len(target_pix[0]))
 *   CWT_FIT_NONFINITE   a fitted weight is NaN / Inf (non-finite features, or the resident kernel's watchdog aborted)
 * from the label counts the fit wrote and the fitted weights. The caller reads `status` back together with the
 * episode's results (e.g. one batch late) and raises then.
 *   label_counts [E,4] int32 (as written by the fit)   w_fit [E,2,C]   status [E] int32 out */
#define CWT_FIT_BAD_LABEL  1
#define CWT_FIT_NO_FG      2
#define CWT_FIT_NONFINITE  4
int cwt_fit_status(const int32_t* label_counts, const float* w_fit, int has_class_weight, int32_t* status,
                   int E, int C, void* stream);

/* ---------------------------------------------------------------------------------------
 * (a-3 / f-3) the same fit for a classifier WITH a bias: nn.Conv2d(C, 2, 1, bias=True), which the
 * reference builds inside CosCls when cls_type[2] == 'b' (src/model/pspnet.py:294,319; fitted by
 * PSPNet.inner_loop, pspnet.py:189-205).  logits = W . F + bias_scale * b  (bias_scale = 1 for the
 * dot classifier; CosCls multiplies conv(x_norm) including its bias by scale_factor — the caller
 * folds scale_factor into the features and passes bias_scale = scale_factor).  Plain SGD on W and b.
 *   b0 [E,2] initial bias, b_out [E,2]; everything else as cwt_fit_classifier_f32 (streaming algorithm)
 * ------------------------------------------------------------------------------------- */
size_t cwt_fit_bias_workspace_bytes(int E, int S, int C, int h, int w, int H, int W);
int cwt_fit_classifier_bias_f32(const float* f_s, const void* s_label, int label_kind, const float* w0,
                                const float* b0, const float* class_weight_or_null, float* w_out, float* b_out,
                                float* loss_trace_or_null, int32_t* label_counts_or_null,
                                int E, int S, int C, int h, int w, int H, int W,
                                int n_iter, float lr, float bias_scale, int ignore_index,
                                void* workspace, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * (f-3) PSPNet.inner_loop (src/model/pspnet.py:189-205) on the reference's cosine classifier CosCls
 * with any cls_type (src/model/pspnet.py:290-323): scores = scale_factor * (weight . x^ + bias) on
 * x^ = F.normalize(x, dim=1, eps=1e-5) (the caller passes x^, cwt_normalize_features_f32), SGD over
 * classifier.parameters().  flags: CWT_COSCLS_R 'r' weight-norm reparametrisation (weight = weight_v,
 * weight_g_or_null = weight_g [E,2]), CWT_COSCLS_N 'n' weight rows re-normalised in place at every
 * forward, CWT_COSCLS_T 't' scale_factor is learned; a bias exists iff bias_or_null != NULL ('b').
 * weight [E,2,C], weight_g [E,2], bias [E,2], scale [E] are IN/OUT (initial values in, fitted out).
 * Everything else as cwt_fit_classifier_f32 (streaming algorithm).
 * ------------------------------------------------------------------------------------- */
#define CWT_COSCLS_R 1
#define CWT_COSCLS_N 2
#define CWT_COSCLS_T 4
size_t cwt_fit_coscls_workspace_bytes(int E, int S, int C, int h, int w, int H, int W);
int cwt_fit_coscls_f32(const float* x_norm, const void* s_label, int label_kind,
                       float* weight, float* weight_g_or_null, float* bias_or_null, float* scale,
                       const float* class_weight_or_null, float* loss_trace_or_null, int32_t* label_counts_or_null,
                       int flags, int E, int S, int C, int h, int w, int H, int W,
                       int n_iter, float lr, int ignore_index,
                       void* workspace, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * (f-3) PSPNet.increment_inner_loop with more than two classes (src/model/pspnet.py:207-221; the
 * multi-way incremental setting of src/train_cca.py:146,322): n_iter plain-SGD steps on ONE K-class
 * bias-free 1x1 classifier with CrossEntropyLoss(weight[K], ignore_index) of the logits up-sampled
 * to HxW (Adapt_SegLoss -> weighted_adpt_ce_loss, src/model/model_util.py:76-98).
 *   f_s [S,C,h,w]   s_label [S,H,W] values in [0,K) or ignore_index   weight [K,C] IN/OUT
 *   class_weight [K] (ones, weight[fg_idx] = (bg_cnt/fg_cnt)^tp)   inv_sum_weight [1] = 1 / sum_i w[y_i]
 *   (both device pointers: the caller counts the labels on the device, no host sync)
 * ------------------------------------------------------------------------------------- */
size_t cwt_fit_multiclass_workspace_bytes(int K, int S, int C, int h, int w, int H, int W);
int cwt_fit_multiclass_f32(const float* f_s, const void* s_label, int label_kind, float* weight,
                           const float* class_weight, const float* inv_sum_weight,
                           int K, int S, int C, int h, int w, int H, int W,
                           int n_iter, float lr, int ignore_index,
                           void* workspace, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * (f-3) the same inner loop with SegLoss('wt_dc' | 'dc'), the per-channel sigmoid dice loss:
 * PSPNet.inner_loop, src/model/pspnet.py:189-205, with criterion = SegLoss(args.inner_loss_type),
 * src/model/model_util.py:18-19 -> weighted_dice_loss, model_util.py:40-73 (weighted_val 1,
 * reduction 'sum', input_type 'lg'):  n_iter x { 1x1 conv 2xC, bilinear up to HxW, sigmoid per
 * channel, sum over (image, channel) rows of 1 - 2 sum(t p) / clamp(sum p^2 + sum t^2, 1e-8),
 * divided by S, backward, SGD }.  Arguments as cwt_fit_classifier_f32 (no class weight: the dice
 * loss has none; labels equal to ignore_index are in neither target but their p^2 counts).
 *   loss_trace_or_null [n_iter,E]: the dice loss of every step
 * ------------------------------------------------------------------------------------- */
size_t cwt_fit_dice_workspace_bytes(int E, int S, int C, int h, int w, int H, int W);
int cwt_fit_classifier_dice_f32(const float* f_s, const void* s_label, int label_kind, const float* w0,
                                float* w_out, float* loss_trace_or_null,
                                int E, int S, int C, int h, int w, int H, int W,
                                int n_iter, float lr, int ignore_index,
                                void* workspace, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * (a-5, a-6) MultiHeadAttentionOne forward.  Replaces src/model/transformer.py:54-83
 * (+ ScaledDotProductAttention :23-30) for the call transformer(W, f_q, f_q) of
 * src/test.py:197 / src/train.py:257 (k is v, one shared projection w_qkvs).
 *   q        [E,Lq,C]  (the classifier weights; Lq = 2 on the path, <= 4 supported)
 *   k        [E,C,HW]  query features; normalize_k != 0 fuses F.normalize(dim=1) (src/test.py:194)
 *   w_qkvs   [nH*C, C]   fc_w [C, nH*C]   fc_b, ln_g, ln_b [C]     (d_model = d_k = d_v = C)
 *   keep_attn_or_null [nH,E,Lq,HW] uint8 keep-mask of the attention dropout (scaled 1/(1-p_attn))
 *   keep_out_or_null  [E,Lq,C]     uint8 keep-mask of the output dropout   (scaled 1/(1-p_out))
 *   out      [E,Lq,C]
 *   saved_or_null: activations kept for cwt_transformer_bwd_f32 (size: cwt_transformer_saved_bytes)
 * ------------------------------------------------------------------------------------- */
size_t cwt_transformer_workspace_bytes(int E, int Lq, int nH, int C, int HW, int algo);
size_t cwt_transformer_saved_bytes(int E, int Lq, int nH, int C, int HW);
int cwt_transformer_fwd_f32(const float* q, const float* k, int normalize_k,
                            const float* w_qkvs, const float* fc_w, const float* fc_b,
                            const float* ln_g, const float* ln_b,
                            const uint8_t* keep_attn_or_null, const uint8_t* keep_out_or_null,
                            float p_attn, float p_out, float* out, void* saved_or_null,
                            int E, int Lq, int nH, int C, int HW, int algo,
                            void* workspace, size_t ws_bytes, void* stream);

/* (a-13) backward of the block w.r.t. its five parameters, given dL/d(out) [E,Lq,C].
 * Gradients are SUMMED over the E episodes of the batch into d_* (which the call overwrites).
 * No gradient w.r.t. q or k (src/train.py:246-253: .data / no_grad). */
int cwt_transformer_bwd_f32(const float* d_out, const float* q, const float* k, int normalize_k,
                            const float* w_qkvs, const float* fc_w, const float* ln_g,
                            const uint8_t* keep_attn_or_null, const uint8_t* keep_out_or_null,
                            float p_attn, float p_out, const void* saved,
                            float* d_w_qkvs, float* d_fc_w, float* d_fc_b, float* d_ln_g, float* d_ln_b,
                            int E, int Lq, int nH, int C, int HW,
                            void* workspace, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * (a-4, a-7..a-12) query logits -> bilinear up to HxW -> argmax -> intersection/union (+CE).
 * Replaces src/test.py:192,200-204,214-223 and batch_intersectionAndUnionGPU /
 * intersectionAndUnionGPU, src/util.py:237-308.
 *   wts      [E,V,2,C]  V weight sets per episode (e.g. adapted and baseline classifier)
 *   f_q      [E,C,h,w]
 *   normalize_mask: bit v set => variant v scores the L2-normalised features (src/test.py:194,204)
 *   iu_counts  [E,V,2,3] int64 out: per class I, U, T   (U = O + T - I as util.py:306)
 *   logits60_or_null [E,V,2,h,w] fp32 out
 *   ce_or_null [E,V,2] double out: sum of -log p[y] over valid pixels, number of valid pixels
 *   weighted CE for training (a-13): see cwt_query_loss_grad below.
 * ------------------------------------------------------------------------------------- */
size_t cwt_logits_iou_workspace_bytes(int E, int V, int C, int h, int w, int H, int W);
int cwt_logits_iou(const float* wts, const float* f_q, const void* q_label, int label_kind,
                   int normalize_mask, long long* iu_counts, float* logits60_or_null,
                   double* ce_or_null, int E, int V, int C, int h, int w, int H, int W,
                   int ignore_index, void* workspace, size_t ws_bytes, void* stream);

/* Same tail on ready-made low-resolution logits: the drop-in for
 * batch_intersectionAndUnionGPU(logits[n,2,h,w] -> counts[n,2,3]).  (src/util.py:237-277) */
int cwt_upsample_argmax_iou(const float* logits60, const void* label, int label_kind,
                            long long* iu_counts, double* ce_or_null,
                            int n, int h, int w, int H, int W, int ignore_index, void* stream);

/* intersectionAndUnionGPU on ready-made integer predictions (src/util.py:280-308), 2 classes+.
 *   preds, target [n, npix] (same label_kind), counts [n, num_classes, 3] int64 */
int cwt_intersection_union(const void* preds, const void* target, int label_kind, long long* counts,
                           int n, long long npix, int num_classes, int ignore_index, void* stream);

/* (a-13) training loss on the query: weighted CE at HxW of up(logits60) and its gradient
 * w.r.t. logits60 (the adjoint bilinear map), class weight [1, n0/(n1+1e-12)] counted on
 * the device (src/train.py:237-243,262-264).
 *   logits60 [E,2,h,w]; label [E,H,W]; loss [E] fp32 out; d_logits60 [E,2,h,w] out */
size_t cwt_query_loss_workspace_bytes(int E, int h, int w, int H, int W);
int cwt_query_loss_grad(const float* logits60, const void* label, int label_kind,
                        float* loss, float* d_logits60, int E, int h, int w, int H, int W,
                        int ignore_index, void* workspace, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * The two skinny contractions behind everything above, exposed for the training step (a-13):
 *   cwt_rows_times_feat : out[e][r][p] = sum_c M[e][r][c] * fn[e][c][p]   (logits = W' X^T, src/train.py:259-261)
 *   cwt_feat_times_rows : out[e][r][c] = sum_p P[e][r][p] * fn[e][c][p]   (its adjoint: dW' = dlogits X)
 * fn = f, or F.normalize(f, dim=1) when normalize != 0.  R <= 16.
 * ------------------------------------------------------------------------------------- */
size_t cwt_skinny_workspace_bytes(int E, int R, int C, int HW);
int cwt_rows_times_feat(const float* M, const float* f, int normalize, float* out,
                        int E, int R, int C, int HW, void* workspace, size_t ws_bytes, void* stream);
int cwt_feat_times_rows(const float* P, const float* f, int normalize, float* out,
                        int E, int R, int C, int HW, void* workspace, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * Feature side of the cosine classifier: out[i][c][p] = scale * f[i][c][p] / max(|f[i,:,p]|_2, eps).
 * Replaces `x_norm = F.normalize(x, p=2, dim=1, eps=0.00001)` and the constant `scale_factor = 2.0` of
 * CosCls.forward (src/model/pspnet.py:302-310): with cls_type 'oooo'/'0000' (no weight-norm, bias or learnable
 * temperature) the cosine classifier is the bias-free 1x1 classifier of cwt_fit_classifier_f32 on these features.
 * f, out: [n_img][C][HW] fp32 (out may not alias f).
 * ------------------------------------------------------------------------------------- */
int cwt_normalize_features_f32(const float* f, float* out, int n_img, int C, int HW, float eps, float scale, void* stream);

/* ---------------------------------------------------------------------------------------
 * (f-4) the reference's validation transform on the device, one fused kernel per image:
 * Resize (src/dataset/transform.py:109-163: aspect-preserving cv2.resize INTER_LINEAR to new_h x new_w,
 * padded to size x size; label cv2.INTER_NEAREST, padded with pad_label) -> ToTensor (/255, transform.py:58-82)
 * -> Normalize ((x - mean) / std, transform.py:85-107), composed in src/dataset/dataset.py:78-84.
 *   image_hwc [ori_h,ori_w,3] fp32 RGB in [0,255] (device)   label_or_null [ori_h,ori_w] uint8 (device)
 *   mean3 / std3 / pad3_or_null: HOST pointers (3 floats; pad in [0,255] units, NULL = zero padding)
 *   out_chw [3,size,size] fp32   label_out [size,size] uint8 or int64 (label_out_kind)
 * ------------------------------------------------------------------------------------- */
int cwt_resize_pad_normalize_f32(const float* image_hwc, const uint8_t* label_or_null, int ori_h, int ori_w,
                                 int new_h, int new_w, int size, const float* mean3, const float* std3,
                                 const float* pad3_or_null, int pad_label, float* out_chw, void* label_out_or_null,
                                 int label_out_kind, void* stream);

/* ---------------------------------------------------------------------------------------
 * Zero-compressed transport of post-ReLU features (the host -> device leg in front of the path: the reference moves
 * dense tensors with .cuda(), src/test.py:153-157; ~half of the elements are zeros). Lossless. The tensor is a matrix of
 * n_rows rows (episodes) of 32 * words_per_row fp32 elements:
 *   mask          [n_rows][words_per_row] uint32   bit l of word w set <=> element 32 w + l has a non-zero BIT PATTERN
 *   block_offsets [n_rows][ceil(words_per_row / 32)] uint32   number of set bits before each block of 32 words, counted
 *                 over the whole compressed batch; base_offset = the count before row 0 of THIS call (so that a slice of
 *                 a larger compressed batch expands without re-basing)
 *   vals          packed fp32 values of the set bits, in element order
 *   out           [n_rows][32 * words_per_row] fp32, 16-byte aligned: the dense tensor, bit-identical to the original
 * ------------------------------------------------------------------------------------- */
int cwt_expand_zero_compressed_f32(const uint32_t* mask, const uint32_t* block_offsets, const float* vals,
                                   float* out, int n_rows, int words_per_row, unsigned base_offset, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CWT_B200_H_ */
