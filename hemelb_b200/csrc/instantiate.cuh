// Included by the per-(lattice, collision kernel) translation units: instantiates the twelve
// (wall link, inlet link, outlet link) bundles of the site kernel and the run-time dispatch over
// them.
#pragma once
#include <algorithm>
#include <cuda_runtime.h>
#include "kernels.cuh"

namespace hlb {

template <int Q, int KERNEL>
void launch_collide_stream(int wall, int inlet, int outlet, const StepArgs& A, const void* mrt, int64_t first,
                           int64_t count, const uint32_t* gzsList, int64_t gzsFirst, int64_t gzsCount, void* stream) {
  // (count == 0 with gzsCount > 0: the per-link kernel alone -- the caller runs it on a stream of its own)
  if (count <= 0 && gzsCount <= 0) return;
  constexpr int T = site_threads<Q>();
  const dim3 block(T);
  const dim3 grid((unsigned)((std::max<int64_t>(count, 1) + T - 1) / T));
  cudaStream_t s = (cudaStream_t)stream;
  const MrtArgs<Q>& M = *(const MrtArgs<Q>*)mrt;
#define HLB_CASE(W, I, O)                                                                       \
  if (wall == W && inlet == I && outlet == O) {                                                 \
    if (count > 0) collide_stream_kernel<Q, KERNEL, W, I, O><<<grid, block, 0, s>>>(A, M, first, count); \
    if constexpr (W == W_GZS) {                                                                 \
      if (A.wallOn && gzsCount > 0) {                                                           \
        StepArgs G = A;                                                                         \
        if (gzsList) G.siteList = gzsList;                                                      \
        unsigned gzsGrid = (unsigned)((gzsCount + kGzsTile - 1) / kGzsTile);                    \
        if (A.gzsGridLimit > 0 && gzsGrid > (unsigned)A.gzsGridLimit) gzsGrid = (unsigned)A.gzsGridLimit; \
        gzs_links_kernel<Q, KERNEL><<<gzsGrid, kGzsThreads, 0, s>>>(G, M, gzsFirst, gzsCount);  \
      }                                                                                         \
    }                                                                                           \
    return;                                                                                     \
  }
  HLB_CASE(W_SBB, I_NASH, I_NASH)
  HLB_CASE(W_SBB, I_NASH, I_LADD)
  HLB_CASE(W_SBB, I_LADD, I_NASH)
  HLB_CASE(W_SBB, I_LADD, I_LADD)
  HLB_CASE(W_BFL, I_NASH, I_NASH)
  HLB_CASE(W_BFL, I_NASH, I_LADD)
  HLB_CASE(W_BFL, I_LADD, I_NASH)
  HLB_CASE(W_BFL, I_LADD, I_LADD)
  HLB_CASE(W_GZS, I_NASH, I_NASH)
  HLB_CASE(W_GZS, I_NASH, I_LADD)
  HLB_CASE(W_GZS, I_LADD, I_NASH)
  HLB_CASE(W_GZS, I_LADD, I_LADD)
#undef HLB_CASE
}

// the TMA-staged persistent form over the device sites [0, count) (kernels.cuh, site_tma_kernel)
template <int Q, int KERNEL>
void launch_site_tma(int wall, int inlet, int outlet, const StepArgs& A, const void* mrt, const CUtensorMap* mapF,
                     const CUtensorMap* mapN, int64_t count, int nSm, const uint32_t* gzsList, int64_t gzsCount, void* stream) {
  if (count <= 0) return;
  using C = TmaCfg<Q>;
  const int64_t nTiles = (count + kTile - 1) / kTile;
  const unsigned grid = (unsigned)(nTiles < nSm ? nTiles : nSm);
  cudaStream_t s = (cudaStream_t)stream;
  const MrtArgs<Q>& M = *(const MrtArgs<Q>*)mrt;
#define HLB_CASE(W, I, O)                                                                                     \
  if (wall == W && inlet == I && outlet == O) {                                                               \
    static bool configured = false;                                                                           \
    if (!configured) {                                                                                        \
      cudaFuncSetAttribute(site_tma_kernel<Q, KERNEL, W, I, O>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                           C::smemBytes);                                                                     \
      configured = true;                                                                                      \
    }                                                                                                         \
    site_tma_kernel<Q, KERNEL, W, I, O><<<grid, C::threads, C::smemBytes, s>>>(A, M, *mapF, *mapN, count, nTiles); \
    if constexpr (W == W_GZS) {                                                                               \
      if (gzsCount > 0) {                                                                                     \
        StepArgs G = A;                                                                                       \
        G.siteList = gzsList;                                                                                 \
        gzs_links_kernel<Q, KERNEL><<<(unsigned)((gzsCount + kGzsTile - 1) / kGzsTile), kGzsThreads, 0, s>>>( \
            G, M, 0, gzsCount);                                                                               \
      }                                                                                                       \
    }                                                                                                         \
    return;                                                                                                   \
  }
  HLB_CASE(W_SBB, I_NASH, I_NASH)
  HLB_CASE(W_SBB, I_NASH, I_LADD)
  HLB_CASE(W_SBB, I_LADD, I_NASH)
  HLB_CASE(W_SBB, I_LADD, I_LADD)
  HLB_CASE(W_BFL, I_NASH, I_NASH)
  HLB_CASE(W_BFL, I_NASH, I_LADD)
  HLB_CASE(W_BFL, I_LADD, I_NASH)
  HLB_CASE(W_BFL, I_LADD, I_LADD)
  HLB_CASE(W_GZS, I_NASH, I_NASH)
  HLB_CASE(W_GZS, I_NASH, I_LADD)
  HLB_CASE(W_GZS, I_LADD, I_NASH)
  HLB_CASE(W_GZS, I_LADD, I_LADD)
#undef HLB_CASE
}

}  // namespace hlb
