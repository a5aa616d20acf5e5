// placeholder until the tcgen05 K-projection lands
#include "common.cuh"
namespace cwt {
size_t kproj_tcgen05_workspace_bytes(int, int, int, int, int) { return 0; }
int kproj_scores_tcgen05(const float*, const float*, const float*, float*, float*, int, int, int, int, int,
                         void*, size_t, cudaStream_t) {
    set_error("CWT_ATTN_TCGEN05 is not available in this build");
    return CWT_ERR_UNSUPPORTED;
}
}  // namespace cwt
