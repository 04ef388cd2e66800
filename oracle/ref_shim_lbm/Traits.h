// TEST INFRASTRUCTURE (oracle/_ref/libhemelb_reflbm.so only): Traits.h for running the reference's own lb::LBM with the
// reference's own CPU streamers.  The reference's Traits.h takes its defaults from lb/Streamers.h, which includes
// every streamer there is -- JunkYang.h needs boost::ublas, absent here; this header names the same ten member
// types for lb::LBM, includes the streamers the driver instantiates, and defaults to D3Q19 LBGK + BFL + Nash.
#pragma once
#include "lb/lattices/D3Q15.h"
#include "lb/lattices/D3Q19.h"
#include "lb/lattices/D3Q27.h"
#include "lb/kernels/LBGK.h"
#include "lb/kernels/MRT.h"
#include "lb/kernels/DHumieresD3Q15MRTBasis.h"
#include "lb/kernels/DHumieresD3Q19MRTBasis.h"
#include "lb/collisions/Normal.h"
#include "lb/streamers/BulkStreamer.h"
#include "lb/streamers/StreamerTypeFactory.h"
#include "lb/streamers/SimpleBounceBack.h"
#include "lb/streamers/BouzidiFirdaousLallemand.h"
#include "lb/streamers/GuoZhengShi.h"
#include "lb/streamers/NashZerothOrderPressure.h"
#include "lb/streamers/LaddIolet.h"

namespace hemelb {
  namespace lb {
    // wall-only + iolet-only link factories -> the factory with both links (what lb/Streamers.h:71-99 provides)
    template <typename WALL_ONLY, typename IOLET_ONLY> struct CombineWallAndIoletStreamers;
    template <typename C, template <typename> class W, template <typename> class I>
    struct CombineWallAndIoletStreamers<StreamerTypeFactory<W<C>, NullLink<C>>, StreamerTypeFactory<NullLink<C>, I<C>>> {
      typedef StreamerTypeFactory<W<C>, I<C>> type;
    };
  }
  namespace lb::cpu {
    // the reference's streamers as the template templates Traits wants (lb/Streamers.h:21-62 builds the same types)
    template <class C> using SbbWall = StreamerTypeFactory<BounceBackLink<C>, NullLink<C>>;
    template <class C> using BflWall = StreamerTypeFactory<BouzidiFirdaousLallemandLink<C>, NullLink<C>>;
    template <class C> using GzsWall = StreamerTypeFactory<GuoZhengShiLink<C>, NullLink<C>>;
    template <class C> using NashIolet = StreamerTypeFactory<NullLink<C>, NashZerothOrderPressureLink<C>>;
    template <class C> using LaddIolet = StreamerTypeFactory<NullLink<C>, LaddIoletLink<C>>;
  }
  namespace redblood::stencil { struct FourPoint; }

  template <typename L = lb::D3Q19, template <lb::lattice_type> class K = lb::LBGK, template <class> class C = lb::Normal,
            template <class> class BULK = lb::BulkStreamer, template <class> class WALL = lb::cpu::BflWall,
            template <class> class IN = lb::cpu::NashIolet, template <class> class OUT = lb::cpu::NashIolet,
            typename S = redblood::stencil::FourPoint>
  struct Traits {
    typedef L Lattice;
    typedef K<L> Kernel;
    typedef C<Kernel> Collision;
    typedef S Stencil;
    typedef BULK<Collision> Streamer;
    typedef WALL<Collision> WallBoundary;
    typedef IN<Collision> InletBoundary;
    typedef OUT<Collision> OutletBoundary;
    typedef typename lb::CombineWallAndIoletStreamers<WallBoundary, InletBoundary>::type WallInletBoundary;
    typedef typename lb::CombineWallAndIoletStreamers<WallBoundary, OutletBoundary>::type WallOutletBoundary;
  };
}
