// Shadow header (oracle/_ref build only): stand-in for the map-backed store of remote sites that
// GuoZhengShi.h reads (Code/geometry/neighbouring/NeighbouringDomain.h).  The driver resolves a
// global site id straight to the owning emulated rank's f_old row.
#pragma once
#include <functional>
#include <span>
#include "units.h"
namespace hemelb::geometry::neighbouring {
  class NeighbouringDomain {};
  struct NeighbouringSite {
    const distribn_t* f;
    template <class L> auto GetFOld() const {
      return std::span<distribn_t, L::NUMVECTORS>{const_cast<distribn_t*>(f), L::NUMVECTORS};
    }
  };
  class NeighbouringFieldData {
  public:
    std::function<const distribn_t*(site_t)> resolve;
    NeighbouringSite GetSite(site_t gid) const { return NeighbouringSite{resolve(gid)}; }
  };
}
