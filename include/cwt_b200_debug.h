/*
 * cwt_b200_debug.h — developer-only entry points of libcwt_b200.so (profiling / microbenchmarks).
 * Not part of the drop-in boundary (include/cwt_b200.h): nothing in the product path calls them and they keep no
 * process-global state — the profile buffer is an argument of the call.
 */
#ifndef CWT_B200_DEBUG_H_
#define CWT_B200_DEBUG_H_

#include "cwt_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* cwt_fit_classifier_f32 (same arguments) on the INSTRUMENTED build of the resident kernel:
 * prof_out [grid][12] int64 (device) receives per-CTA cycle counters (P1, halo wait, HR, P3, waits, ...; see
 * tools/prof_resident.py). 1-shot, CWT_FIT_RESIDENT only. */
int cwt_debug_fit_classifier_prof_f32(const float* f_s, const void* s_label, int label_kind, const float* w0,
                                      const float* class_weight_or_null, float* w_out,
                                      int E, int C, int h, int w, int H, int W,
                                      int n_iter, float lr, int ignore_index,
                                      void* workspace, size_t ws_bytes, long long* prof_out, void* stream);

/* L2 read-bandwidth microbenchmark — `ctas` x 512 threads sweep `bytes` of `buf` `iters` times with
 * L1-bypassing 128-bit loads (time it with events; bytes * iters / time = bandwidth) */
int cwt_debug_l2_read(const void* buf, size_t bytes, int iters, int ctas, void* sink, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CWT_B200_DEBUG_H_ */
