// tcgen05 / TMEM bf16 GEMM (placeholder until the tensor-core kernel lands in this file).
#include "common.cuh"
namespace nfdpm {
int gemm_nt_tc(const void*, int64_t, const void*, int64_t, void*, int64_t, int, int, int, int, int, const float*,
               const float*, cudaStream_t) {
  return fail("nfdpm_gemm_nt: bf16 tensor-core path not built in this version");
}
}  // namespace nfdpm
