// Fused mid-domain kernel, D3Q27 TRT, BFL walls, LADD inlet / NASH outlet.
#include "fused_impl.cuh"
namespace hlb {
template void launch_fused_bundle<27, K_TRT, W_BFL, I_LADD, I_NASH>(const StepArgs&, const MrtArgs<27>&, const IoletDev*, const double*,
                                                       const MidItem*, int64_t, void*);
}
