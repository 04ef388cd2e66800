// Fused mid-domain kernel, D3Q15 MRT, BFL walls, NASH inlet / NASH outlet.
#include "fused_impl.cuh"
namespace hlb {
template void launch_fused_bundle<15, K_MRT, W_BFL, I_NASH, I_NASH>(const StepArgs&, const MrtArgs<15>&, const IoletDev*, const double*,
                                                       const MidItem*, int64_t, void*);
}
