// explicit instantiation: two-for-one transforms, float, y mode (see r2r2_inst.cuh)
#include "r2r2_inst.cuh"
namespace cb {
template int r2r2_run<float, true, false>(const R2Args<float>&, int, int, bool, cudaStream_t);
template int r2r2_query<true, true>(int, int, int[4]);
}  // namespace cb
