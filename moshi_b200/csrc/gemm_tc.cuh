// tcgen05 / TMA weight-streaming GEMM for the LM linears (placeholder interface; see gemm_tc.cu).
#pragma once
#include "common.cuh"

namespace b200 {
namespace tc {

struct GemmPlanCache {
  void clear() {}
};
inline bool supported(int, int, int, int) { return false; }
inline int prepare_plans(GemmPlanCache&) { return B200_OK; }
inline int linear(GemmPlanCache&, const __nv_bfloat16*, long long, const __nv_bfloat16*, __nv_bfloat16*, long long,
                  const __nv_bfloat16*, long long, int, int, int, int, int, cudaStream_t) {
  B200_FAIL(B200_ERR_INVALID, "tcgen05 GEMM path not built");
}

}  // namespace tc
}  // namespace b200
