// D3Q27 LBGK: collide-and-stream kernels for all wall / iolet link policies.
#include "instantiate.cuh"
namespace hlb {
template void launch_collide_stream<27, K_LBGK>(int, int, int, const StepArgs&, const void*, int64_t, int64_t, const uint32_t*,
                                             int64_t, int64_t, void*);
template void launch_site_tma<27, K_LBGK>(int, int, int, const StepArgs&, const void*, const CUtensorMap*, const CUtensorMap*, int64_t, int,
                                       const uint32_t*, int64_t, void*);
}
