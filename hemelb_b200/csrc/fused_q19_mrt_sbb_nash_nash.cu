// Fused mid-domain kernel, D3Q19 MRT, SBB walls, NASH inlet / NASH outlet.
#include "fused_impl.cuh"
namespace hlb {
template void launch_fused_bundle<19, K_MRT, W_SBB, I_NASH, I_NASH>(const StepArgs&, const MrtArgs<19>&, const IoletDev*, const double*,
                                                       const MidItem*, int64_t, void*);
}
