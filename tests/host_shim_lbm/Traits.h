// Test infrastructure: Traits.h for the harness around the reference's own lb::LBM (tests/host_lbm_real.cc).
//
// lb::LBM<TRAITS> reads ten member types off its TRAITS (Lattice, Kernel, Collision, Streamer, WallBoundary,
// InletBoundary, OutletBoundary, WallInletBoundary, WallOutletBoundary, Stencil).  The reference's Traits.h supplies
// them from seven template template parameters whose *defaults* pull in every streamer it has (lb/Streamers.h ->
// JunkYang.h needs boost::ublas) and redblood/stencil.h; this one takes the same parameters in the same order,
// defaults them to D3Q19 / LBGK / Normal and the gpu:: streamers, and includes only what those need.  A HemeLB
// build keeps its own Traits.h and names the gpu:: streamers in an instantiation.
#pragma once
#include "lb/lattices/D3Q15.h"
#include "lb/lattices/D3Q19.h"
#include "lb/lattices/D3Q27.h"
#include "lb/kernels/LBGK.h"
#include "lb/kernels/MRT.h"
#include "lb/kernels/DHumieresD3Q15MRTBasis.h"
#include "lb/kernels/DHumieresD3Q19MRTBasis.h"
#include "lb/collisions/Normal.h"
#include "lb/streamers/StreamerTypeFactory.h"
#include "lb/streamers/GpuStreamers.h"

namespace hemelb {
  namespace redblood::stencil { struct FourPoint; }  // (named by the last parameter only)

  template <typename L = lb::D3Q19,
            template <lb::lattice_type> class K = lb::LBGK,
            template <class> class C = lb::Normal,
            template <class> class BULK = lb::gpu::Bulk,
            template <class> class WALL = lb::gpu::Wall<lb::gpu::BouzidiFirdaousLallemand>::template type,
            template <class> class IN = lb::gpu::Inlet<lb::gpu::NashZerothOrderPressure>::template type,
            template <class> class OUT = lb::gpu::Outlet<lb::gpu::NashZerothOrderPressure>::template type,
            typename S = redblood::stencil::FourPoint>
  struct Traits {
    typedef L Lattice;
    typedef K<L> Kernel;
    typedef C<Kernel> Collision;
    typedef S Stencil;
    // the six streamers lb::LBM constructs (lb.h:40-52)
    typedef BULK<Collision> Streamer;
    typedef WALL<Collision> WallBoundary;
    typedef IN<Collision> InletBoundary;
    typedef OUT<Collision> OutletBoundary;
    typedef typename lb::CombineWallAndIoletStreamers<WallBoundary, InletBoundary>::type WallInletBoundary;
    typedef typename lb::CombineWallAndIoletStreamers<WallBoundary, OutletBoundary>::type WallOutletBoundary;
  };
}
