// explicit instantiation: y transforms with peer-mapped (split) rows for the distributed solve, double
#include "r2r2_inst.cuh"
namespace cb {
template int r2r2_run<double, true, true>(const R2Args<double>&, int, int, bool, cudaStream_t);
}  // namespace cb
