// tcgen05 / TMA weight-streaming GEMM for the LM linears:  y[M][N] = epi(x[M][K] . w[N][K]^T)
//
// Skinny-M decode GEMMs are HBM-bound on the weights, so the kernel is organised around streaming
// each weight row exactly once ("swap-AB"): the 128-row weight tile is the UMMA A operand (M_umma =
// 128), the B <= 256 sessions are the UMMA N dimension, and the fp32 accumulator [128 x Mpad] lives
// in TMEM.  Both operands are K-major in HBM, which is the native TMA / UMMA SWIZZLE_128B layout, so
// no repacking of the reference's [out, in] weight layout is needed.
#pragma once

#include <cuda.h>

#include "common.cuh"

namespace b200 {
namespace tc {

struct PlanKey {
  const void* ptr; long long ld; int rows, cols, box_rows;
  bool operator<(const PlanKey& o) const {
    if (ptr != o.ptr) return ptr < o.ptr;
    if (ld != o.ld) return ld < o.ld;
    if (rows != o.rows) return rows < o.rows;
    if (cols != o.cols) return cols < o.cols;
    return box_rows < o.box_rows;
  }
};

struct GemmPlanCache {
  std::map<PlanKey, CUtensorMap> maps;
  void clear() { maps.clear(); }
};

bool supported(int M, int N, int K, int epi);
// implementation the LM uses for this shape: 1 = SIMT weight-streaming kernel, 2 = tcgen05 kernel
int auto_pick(int M, int N, int K, int epi);
int prepare_plans(GemmPlanCache& cache);
int linear(GemmPlanCache& cache, const __nv_bfloat16* x, long long ldx, const __nv_bfloat16* w, __nv_bfloat16* y,
           long long ldy, const __nv_bfloat16* res, long long ldr, int M, int N, int K, int epi, int gate_rows,
           cudaStream_t stream);

}  // namespace tc
}  // namespace b200
