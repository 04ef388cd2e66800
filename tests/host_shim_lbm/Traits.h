// Traits.h for the real-lb::LBM harness: hemelb::Traits as Code/Traits.h declares it -- same template
// parameters, same member names -- but without the defaults that drag in every streamer of the reference
// (lb/Streamers.h -> JunkYang.h needs boost::ublas; redblood/stencil.h).  A HemeLB build keeps its own
// Traits.h and names the gpu:: streamers in an instantiation.  Test infrastructure.
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

namespace hemelb
{
  namespace redblood::stencil { struct FourPoint; }
  template <
      typename LATTICE = lb::D3Q19,
      template<lb::lattice_type> class KERNEL = lb::LBGK,
      template<class> class COLLISION = lb::Normal,
      template<class> class STREAMER = lb::gpu::Bulk,
      template<class> class WALL_BOUNDARY = lb::gpu::Wall<lb::gpu::BouzidiFirdaousLallemand>::template type,
      template<class> class INLET_BOUNDARY = lb::gpu::Inlet<lb::gpu::NashZerothOrderPressure>::template type,
      template<class> class OUTLET_BOUNDARY = lb::gpu::Outlet<lb::gpu::NashZerothOrderPressure>::template type,
      typename STENCIL = redblood::stencil::FourPoint
  >
  struct Traits
  {
    using Lattice = LATTICE;
    using Kernel = KERNEL<Lattice>;
    using Collision = COLLISION<Kernel>;
    using Streamer = STREAMER<Collision>;
    using WallBoundary = WALL_BOUNDARY<Collision>;
    using InletBoundary = INLET_BOUNDARY<Collision>;
    using OutletBoundary = OUTLET_BOUNDARY<Collision>;
    using WallInletBoundary = typename lb::CombineWallAndIoletStreamers<WallBoundary, InletBoundary>::type;
    using WallOutletBoundary = typename lb::CombineWallAndIoletStreamers<WallBoundary, OutletBoundary>::type;
    using Stencil = STENCIL;
  };
}
