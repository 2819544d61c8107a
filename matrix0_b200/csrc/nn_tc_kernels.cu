// BF16 tensor-core path of the PolicyValueNet forward (tcgen05 / TMEM / TMA implicit GEMM).
// Placeholder translation unit: the entry points exist so that the library links; the kernels land
// in the next commit.  Calling the bf16 precision before that fails loudly.
#include "nn.cuh"

struct m0_net;
namespace m0 {
int tc_net_prepare(::m0_net*, cudaStream_t) {
  m0_set_error("bf16 tensor-core path not available in this build");
  return M0_ERR_STATE;
}
int tc_net_forward(::m0_net*, const float*, int, float*, float*, cudaStream_t) {
  m0_set_error("bf16 tensor-core path not available in this build");
  return M0_ERR_STATE;
}
void tc_net_release(::m0_net*) {}
}  // namespace m0
