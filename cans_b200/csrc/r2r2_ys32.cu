// explicit instantiation: y transforms with peer-mapped (split) rows for the distributed solve, float
#include "r2r2_inst.cuh"
namespace cb {
template int r2r2_run<float, true, true>(const R2Args<float>&, int, int, bool, cudaStream_t);
}  // namespace cb
