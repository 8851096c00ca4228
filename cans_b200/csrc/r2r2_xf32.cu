// explicit instantiation: forward x transforms with the fused fillps source, float (see r2r2_inst.cuh)
#include "r2r2_inst.cuh"
namespace cb {
template int r2r2_run_fillps<float>(const R2Args<float>&, const R2Fill<float>&, int, cudaStream_t);
}  // namespace cb
