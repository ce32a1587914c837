// coarse_tc.cu -- tcgen05/TMEM/TMA coarse matching (placeholder until the tensor-core kernels land).
#include "common.cuh"

namespace pope {
bool coarse_tc_supported(const CoarseProblem&) { return false; }
cudaError_t coarse_tc_run(const CoarseProblem&, const CoarseScratch&, cudaStream_t) { return cudaErrorNotSupported; }
}  // namespace pope
