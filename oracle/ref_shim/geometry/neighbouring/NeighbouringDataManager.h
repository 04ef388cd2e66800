// Shadow header (oracle/_ref build only): GuoZhengShi.h's constructor registers the remote sites
// it will read; the oracle driver resolves them directly from the owning rank's f_old.
#pragma once
#include <vector>
#include "units.h"
#include "geometry/neighbouring/RequiredSiteInformation.h"
namespace hemelb::geometry::neighbouring {
  class NeighbouringDataManager {
  public:
    std::vector<site_t> needed;
    void RegisterNeededSite(site_t gid, RequiredSiteInformation = RequiredSiteInformation(true)) {
      needed.push_back(gid);
    }
  };
}
