// MR_BF16 (tcgen05) implementation of the CNN news encoder.  (under construction)
#include "news_cnn_tc.cuh"

namespace mr {
int64_t news_cnn_tc_workspace_bytes(const mr_cnn_shape*, int) { return 256; }
int news_cnn_tc_fwd(const mr_cnn_shape*, const void*, int, const float*, const void*, int, const void*, const float*,
                    const float*, const float*, const float*, const float*, void*, void*, float*, float*, void*, int64_t,
                    cudaStream_t) {
  return set_err(MR_ERR_UNSUPPORTED, "MR_BF16 news encoder not built yet");
}
int news_cnn_tc_bwd(const mr_cnn_shape*, const void*, int, const float*, const void*, const float*, const float*,
                    const float*, const void*, const void*, const float*, const float*, const float*, float*, float*,
                    float*, float*, float*, void*, void*, int64_t, cudaStream_t) {
  return set_err(MR_ERR_UNSUPPORTED, "MR_BF16 news encoder not built yet");
}
}  // namespace mr
