// explicit instantiation: forward x transforms with the fused fillps source, double (see r2r2_inst.cuh)
#include "r2r2_inst.cuh"
namespace cb {
template int r2r2_run_fillps<double>(const R2Args<double>&, const R2Fill<double>&, int, cudaStream_t);
}  // namespace cb
