// Included by the per-(lattice, collision kernel) translation units: instantiates the twelve
// (wall link, iolet link) streamer variants and the run-time dispatch over them.
#pragma once
#include <cuda_runtime.h>
#include "kernels.cuh"

namespace hlb {

template <int Q, int KERNEL>
void launch_collide_stream(int wall, int iolet, const StepArgs& A, const void* mrt, int64_t first, int64_t count,
                           void* stream) {
  if (count <= 0) return;
  constexpr int T = site_threads<Q>();
  const dim3 block(T);
  const dim3 grid((unsigned)((count + T - 1) / T));
  cudaStream_t s = (cudaStream_t)stream;
  const MrtArgs<Q>& M = *(const MrtArgs<Q>*)mrt;
#define HLB_CASE(W, I)                                                              \
  if (wall == W && iolet == I) {                                                    \
    collide_stream_kernel<Q, KERNEL, W, I><<<grid, block, 0, s>>>(A, M, first, count); \
    if constexpr (W == W_GZS)                                                       \
      gzs_links_kernel<Q, KERNEL, I><<<(unsigned)((count + kGzsTile - 1) / kGzsTile), kGzsThreads, 0, s>>>(A, M, first, count); \
    return;                                                                         \
  }
  HLB_CASE(W_NONE, I_NONE)
  HLB_CASE(W_SBB, I_NONE)
  HLB_CASE(W_BFL, I_NONE)
  HLB_CASE(W_GZS, I_NONE)
  HLB_CASE(W_NONE, I_NASH)
  HLB_CASE(W_NONE, I_LADD)
  HLB_CASE(W_SBB, I_NASH)
  HLB_CASE(W_SBB, I_LADD)
  HLB_CASE(W_BFL, I_NASH)
  HLB_CASE(W_BFL, I_LADD)
  HLB_CASE(W_GZS, I_NASH)
  HLB_CASE(W_GZS, I_LADD)
#undef HLB_CASE
}

template <int Q, int KERNEL>
bool launch_fused_mid(int wall, int inlet, int outlet, const StepArgs& A, const void* mrt, const IoletDev* inletIolets,
                      const double* inletDensity, const MidItem* items, int64_t nItems, void* stream) {
  // only where it pays (see fused_mid_kernel): MRT and D3Q27; the bundles live in fused_q*_*.cu
  if constexpr (KERNEL == K_MRT || Q > 19) {
    const MrtArgs<Q>& M = *(const MrtArgs<Q>*)mrt;
#define HLB_FUSED(W, I, O)                                                                                   \
  if (wall == W && inlet == I && outlet == O) {                                                              \
    if (nItems > 0) launch_fused_bundle<Q, KERNEL, W, I, O>(A, M, inletIolets, inletDensity, items, nItems, stream); \
    return true;                                                                                             \
  }
    // the policy bundles of BASELINE.json's configs (GZS keeps its two-kernel form)
    HLB_FUSED(W_SBB, I_NASH, I_NASH)
    HLB_FUSED(W_BFL, I_NASH, I_NASH)
    HLB_FUSED(W_BFL, I_LADD, I_NASH)
#undef HLB_FUSED
  }
  return false;
}

}  // namespace hlb
