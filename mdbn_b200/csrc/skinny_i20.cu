// cd_skinny_kernel<20, *> instantiations (a file of their own so that the build compiles them in parallel)
#include "skinny_kernel.cuh"
namespace mdbn {
namespace sk {
template int launch<20, false>(mdbn_ctx*, const CUtensorMap*, const Params&, const Geometry&, cudaStream_t);
template int launch<20, true>(mdbn_ctx*, const CUtensorMap*, const Params&, const Geometry&, cudaStream_t);
}  // namespace sk
}  // namespace mdbn
