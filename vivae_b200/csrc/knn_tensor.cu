// Tensor-core screening path (tcgen05 / TMEM / TMA) -- placeholder until the kernel lands.
#include "common.cuh"

namespace vvb {
bool knn_tensor_supported(int64_t, int, int) { return false; }
size_t knn_tensor_workspace_bytes(int64_t, int, int, int64_t) { return 256; }
int launch_knn_tensor(const float *, int64_t, int, int, int64_t, int64_t, int64_t *, float *, void *, size_t,
                      cudaStream_t, vvb_knn_stats *) {
  set_error("tensor-core kNN path not built");
  return VVB_ENOTSUP;
}
}  // namespace vvb
