// Test-only stand-in for geometry::Domain carrying the member NAMES the drop-in FieldData reads
// through `friend class FieldData` (Code/geometry/Domain.h:55,71-82,257-282,340,495-530), so the
// host headers can be compile-checked here without MPI / Boost.
#pragma once
#include <chrono>
#include <cstdio>
#include <memory>
#include <span>
#include <stdexcept>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>
#include "units.h"
#include "constants.h"
#include "util/Vector3D.h"
#include "geometry/SiteData.h"
#include "geometry/Site.h"
#include "geometry/neighbouring/NeighbouringDomain.h"
#include "lb/lattices/LatticeInfo.h"
struct HostDomainFiller;
namespace hemelb::geometry {
  struct NeighbouringProcessor { proc_t Rank; site_t SharedDistributionCount; site_t FirstSharedDistribution; };
  // stand-in for net::MpiCommunicator: size / rank as the harness sets them; Broadcast goes
  // through a file when one is named (two harness processes, one per GPU, sharing the NCCL id)
  struct FakeComm {
    int size = 1, rank = 0;
    std::string idFile;
    int Size() const { return size; }
    int Rank() const { return rank; }
    template <class T, std::size_t N> void Broadcast(std::span<T, N> data, int root) const {
      if (idFile.empty()) return;
      if (rank == root) {
        const std::string tmp = idFile + ".tmp";
        FILE* fh = fopen(tmp.c_str(), "wb");
        if (!fh || fwrite(data.data(), sizeof(T), data.size(), fh) != data.size()) throw std::runtime_error("FakeComm: cannot write " + tmp);
        fclose(fh);
        if (rename(tmp.c_str(), idFile.c_str())) throw std::runtime_error("FakeComm: cannot publish " + idFile);
      } else {
        for (int tries = 0; tries < 6000; ++tries) {
          if (FILE* fh = fopen(idFile.c_str(), "rb")) {
            const size_t n = fread(data.data(), sizeof(T), data.size(), fh);
            fclose(fh);
            if (n == data.size()) return;
          }
          std::this_thread::sleep_for(std::chrono::milliseconds(10));
        }
        throw std::runtime_error("FakeComm: nothing arrived in " + idFile);
      }
    }
  };
  class FieldData;
  class Domain {
    friend class FieldData;
    friend struct ::HostDomainFiller;  // tests/host_lbm_run.cc fills the tables
    template <class> friend class Site;
  public:
    explicit Domain(const lb::LatticeInfo& li) : latticeInfo(li) {}
    FakeComm const& GetCommunicator() const { return comms; }
    site_t const& GetLocalFluidSiteCount() const { return nSites; }
    site_t GetMidDomainSiteCount() const { site_t n = 0; for (auto c : mid) n += c; return n; }
    site_t const& GetMidDomainCollisionCount(unsigned t) const { return mid[t]; }
    site_t const& GetDomainEdgeCollisionCount(unsigned t) const { return edge[t]; }
    int GetLocalRank() const { return comms.rank; }
    Site<Domain> GetSite(site_t i) { return Site<Domain>(i, *this); }
    Site<const Domain> GetSite(site_t i) const { return Site<const Domain>(i, *this); }
    template <class L> distribn_t GetCutDistance(site_t i, int d) const { return distanceToWall[i * (L::NUMVECTORS - 1) + d - 1]; }
    distribn_t* GetCutDistances(site_t i) { return &distanceToWall[i]; }
    const distribn_t* GetCutDistances(site_t i) const { return &distanceToWall[i]; }
    util::Vector3D<distribn_t>& GetNormalToWall(site_t i) { return wallNormalAtSite[i]; }
    const util::Vector3D<distribn_t>& GetNormalToWall(site_t i) const { return wallNormalAtSite[i]; }
    template <class L> site_t GetStreamedIndex(site_t i, unsigned d) const { return neighbourIndices[i * L::NUMVECTORS + d]; }
    SiteData& GetSiteData(site_t i) { return siteData[i]; }
    const SiteData& GetSiteData(site_t i) const { return siteData[i]; }
    const util::Vector3D<site_t>& GetGlobalSiteCoords(site_t i) const { return globalSiteCoords[i]; }
    // where any fluid site of the whole geometry lives (Domain.h:215-242; the reference asks its
    // distributed store): filled by the harness from the case file when there are several ranks
    site_t GetGlobalNoncontiguousSiteIdFromGlobalCoords(const util::Vector3D<site_t>& c) const {
      return (c.x() * sites.y() + c.y()) * sites.z() + c.z();
    }
    proc_t GetProcIdFromGlobalCoords(const util::Vector3D<site_t>& c) const {
      auto it = whereIs.find(GetGlobalNoncontiguousSiteIdFromGlobalCoords(c));
      return it == whereIs.end() ? SITE_OR_BLOCK_SOLID : it->second.first;
    }
    proc_t ProcProvidingSiteByGlobalNoncontiguousId(site_t id) const {
      auto it = whereIs.find(id);
      return it == whereIs.end() ? SITE_OR_BLOCK_SOLID : it->second.first;
    }
    site_t GetLocalContiguousIdFromGlobalNoncontiguousId(site_t id) const { return whereIs.at(id).second; }
  private:
    util::Vector3D<site_t> sites{1, 1, 1};
    std::unordered_map<site_t, std::pair<proc_t, site_t>> whereIs;  // global id -> (rank, local contiguous id)
    const lb::LatticeInfo& latticeInfo;
    site_t nSites = 0, mid[COLLISION_TYPES] = {}, edge[COLLISION_TYPES] = {};
    site_t totalSharedFs = 0;
    std::vector<NeighbouringProcessor> neighbouringProcs;
    std::vector<distribn_t> distanceToWall;
    std::vector<util::Vector3D<distribn_t>> wallNormalAtSite;
    std::vector<SiteData> siteData;
    std::vector<util::Vector3D<site_t>> globalSiteCoords;
    std::vector<site_t> neighbourIndices, streamingIndicesForReceivedDistributions;
    std::shared_ptr<neighbouring::NeighbouringDomain> neighbouringData;
    FakeComm comms;
  };
}
