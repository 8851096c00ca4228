// explicit instantiation: two-for-one transforms, double, y mode (see r2r2_inst.cuh)
#include "r2r2_inst.cuh"
namespace cb {
template int r2r2_run<double, true, false>(const R2Args<double>&, int, int, bool, cudaStream_t);
template int r2r2_query<true, false>(int, int, int[4]);
}  // namespace cb
