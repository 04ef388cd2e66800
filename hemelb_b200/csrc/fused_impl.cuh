// Definition of one fused mid-domain bundle launcher (see fused_mid_kernel in kernels.cuh).  Included
// only by the fused_q*_*.cu translation units, each of which instantiates exactly one bundle.
#pragma once
#include <cuda_runtime.h>
#include "kernels.cuh"

namespace hlb {

template <int Q, int KERNEL, int WALL, int INLET, int OUTLET>
void launch_fused_bundle(const StepArgs& A, const MrtArgs<Q>& M, const IoletDev* inletIolets, const double* inletDensity,
                         const MidItem* items, int64_t nItems, void* stream) {
  fused_mid_kernel<Q, KERNEL, WALL, INLET, OUTLET><<<(unsigned)nItems, site_threads<Q>(), 0, (cudaStream_t)stream>>>(
      A, M, inletIolets, inletDensity, items);
}

}  // namespace hlb
