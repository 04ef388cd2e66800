// Shadow header (oracle/_ref build only): a vector-backed stand-in for geometry::Domain exposing
// exactly the members the reference's geometry/Site.h, streamers and MacroscopicPropertyCache
// use (Code/geometry/Domain.h:160-530).  Tables are filled by the oracle driver.
#pragma once
#include <vector>
#include <functional>
#include "units.h"
#include "constants.h"
#include "util/Vector3D.h"
#include "geometry/SiteData.h"
#include "geometry/Site.h"
#include "geometry/neighbouring/NeighbouringDomain.h"
namespace hemelb::geometry {
  class Domain {
  public:
    template <class> friend class Site;
    site_t nSites = 0;
    int numVectors = 0;
    int localRank = 0;
    std::vector<site_t> neighbourIndices;
    std::vector<distribn_t> distanceToWall;
    std::vector<util::Vector3D<distribn_t>> wallNormalAtSite;
    std::vector<SiteData> siteData;
    std::vector<util::Vector3D<site_t>> globalSiteCoords;
    // coordinate lookups supplied by the driver (GZS only)
    std::function<proc_t(const util::Vector3D<site_t>&)> procOf;
    std::function<site_t(const util::Vector3D<site_t>&)> contigOf;
    std::function<site_t(const util::Vector3D<site_t>&)> globalIdOf;

    site_t const& GetLocalFluidSiteCount() const { return nSites; }
    Site<Domain> GetSite(site_t i) { return Site<Domain>(i, *this); }
    Site<const Domain> GetSite(site_t i) const { return Site<const Domain>(i, *this); }
    template <class L> distribn_t GetCutDistance(site_t i, int d) const {
      return distanceToWall[i * (L::NUMVECTORS - 1) + d - 1];
    }
    distribn_t* GetCutDistances(site_t i) { return &distanceToWall[i * (numVectors - 1)]; }
    const distribn_t* GetCutDistances(site_t i) const { return &distanceToWall[i * (numVectors - 1)]; }
    util::Vector3D<distribn_t>& GetNormalToWall(site_t i) { return wallNormalAtSite[i]; }
    const util::Vector3D<distribn_t>& GetNormalToWall(site_t i) const { return wallNormalAtSite[i]; }
    template <class L> site_t GetStreamedIndex(site_t i, unsigned d) const {
      return neighbourIndices[i * L::NUMVECTORS + d];
    }
    SiteData& GetSiteData(site_t i) { return siteData[i]; }
    const SiteData& GetSiteData(site_t i) const { return siteData[i]; }
    const util::Vector3D<site_t>& GetGlobalSiteCoords(site_t i) const { return globalSiteCoords[i]; }
    proc_t GetLocalRank() const { return localRank; }
    proc_t GetProcIdFromGlobalCoords(const util::Vector3D<site_t>& c) const { return procOf(c); }
    site_t GetContiguousSiteId(const util::Vector3D<site_t>& c) const { return contigOf(c); }
    // extraction / checkpoint sources (LbDataSourceIterator.cc:92-112, LocalDistributionInput.cc:133-146)
    bool GetContiguousSiteId(const util::Vector3D<site_t>& c, proc_t& rank, site_t& index) const {
      rank = procOf(c);
      if (rank == SITE_OR_BLOCK_SOLID) return false;
      index = contigOf(c);
      return true;
    }
    bool IsValidLatticeSite(const util::Vector3D<site_t>& c) const {
      for (int k = 0; k < 3; ++k) if (c[k] < 0 || c[k] >= latticeExtent[k]) return false;
      return true;
    }
    util::Vector3D<site_t> latticeExtent{site_t(1) << 40, site_t(1) << 40, site_t(1) << 40};
    struct LatticeInfoView { unsigned n; unsigned GetNumVectors() const { return n; } };
    LatticeInfoView GetLatticeInfo() const { return LatticeInfoView{(unsigned)numVectors}; }
    site_t GetGlobalNoncontiguousSiteIdFromGlobalCoords(const util::Vector3D<site_t>& c) const { return globalIdOf(c); }
    neighbouring::NeighbouringDomain ndom;
    neighbouring::NeighbouringDomain& GetNeighbouringData() { return ndom; }
    const neighbouring::NeighbouringDomain& GetNeighbouringData() const { return ndom; }
  };
}
