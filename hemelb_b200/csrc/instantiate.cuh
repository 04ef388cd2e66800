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

}  // namespace hlb
