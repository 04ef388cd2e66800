// Compile-only check (tests/test_host_headers.py): the Gpu*Streamer policy classes satisfy the
// reference's own lb::streamer concept (Code/lb/concepts.h:90-103) and plug into hemelb::Traits'
// wall + iolet combination (Code/lb/Streamers.h:71-99).  Containers come from oracle/ref_shim.
#include "lb/lattices/D3Q15.h"
#include "lb/lattices/D3Q19.h"
#include "lb/lattices/D3Q27.h"
#include "lb/kernels/LBGK.h"
#include "lb/kernels/MRT.h"
#include "lb/kernels/DHumieresD3Q19MRTBasis.h"
#include "lb/collisions/Normal.h"
#include "lb/streamers/StreamerTypeFactory.h"
#include "lb/streamers/GpuStreamers.h"

using namespace hemelb;
namespace g = hemelb::lb::gpu;

template <class C>
void check() {
  using Bulk = g::Bulk<C>;
  using Wall = typename g::Wall<g::BouzidiFirdaousLallemand>::template type<C>;
  using In = typename g::Inlet<g::NashZerothOrderPressure>::template type<C>;
  using Out = typename g::Outlet<g::LaddIolet>::template type<C>;
  using WallIn = typename lb::CombineWallAndIoletStreamers<Wall, In>::type;
  using WallOut = typename lb::CombineWallAndIoletStreamers<Wall, Out>::type;
  static_assert(lb::streamer<Bulk>);
  static_assert(lb::streamer<Wall>);
  static_assert(lb::streamer<In>);
  static_assert(lb::streamer<Out>);
  static_assert(lb::streamer<WallIn>);
  static_assert(lb::streamer<WallOut>);
  static_assert(std::is_same_v<WallIn, g::GpuStreamer<C, 4, g::BouzidiFirdaousLallemand, g::NashZerothOrderPressure>>);
  static_assert(std::is_same_v<WallOut, g::GpuStreamer<C, 5, g::BouzidiFirdaousLallemand, g::LaddIolet>>);
}

template void check<lb::Normal<lb::LBGK<lb::D3Q15>>>();
template void check<lb::Normal<lb::LBGK<lb::D3Q19>>>();
template void check<lb::Normal<lb::LBGK<lb::D3Q27>>>();
template void check<lb::Normal<lb::MRT<lb::DHumieresD3Q19MRTBasis>>>();
