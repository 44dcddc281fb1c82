// K2 tensor-core contraction (tcgen05).  Placeholder until the fp32 path is
// parity-green on hardware: reports "unsupported" so saga_cqt_exec uses cqt.cu.
#include "cqt_plan.cuh"
#include "saga_common.cuh"

namespace saga {
void cqt_umma_plan_init(saga_cqt_plan*) {}
void cqt_umma_plan_free(saga_cqt_plan*) {}
int cqt_umma_exec(const saga_cqt_plan*, const CqtLevels&, int, int64_t, int64_t, float*, float2*, int64_t,
                  int64_t, cudaStream_t) {
  return SAGA_ERR_UNSUPPORTED;
}
}  // namespace saga
