// geometry/neighbouring/NeighbouringDataManager.h -- the GuoZhengShi site halo for a build whose
// distributions live on the B200.
//
// Stands in for Code/geometry/neighbouring/NeighbouringDataManager.{h,cc} (same class, same public
// members, so configuration/SimBuilder.h:153-160,235-236 and lb::InitParams compile unchanged) when
// hemelb_b200/host precedes Code/ on the include path.  What stays as in the reference: the registry of
// needed sites (RegisterNeededSite keeps a global id once) and ShareNeeds, which tells every rank,
// through the reference's own net, which of its sites the others need.  What changes: the per-step
// transfer.  The reference's RequestComms (phase 0 of every time step, NeighbouringDataManager.cc:
// 101-142) sends site.GetFOld(Q) of every served site out of the host array and receives into
// NeighbouringFieldData; through the device-backed FieldData that would pull the whole f_old array to
// the host every step.  Here it is one call: hlb_gpu_exchange_site_halo packs the served rows on the
// device and moves them GPU to GPU (ncclSend / ncclRecv).  The Gpu streamers read the lists kept here
// when they build the engine (lb/streamers/GpuStreamers.h: hlb_gpu_set_gzs_remote / _serve).
#ifndef HEMELB_GEOMETRY_NEIGHBOURING_NEIGHBOURINGDATAMANAGER_H
#define HEMELB_GEOMETRY_NEIGHBOURING_NEIGHBOURINGDATAMANAGER_H

#include <algorithm>
#include <span>
#include <vector>

#include "geometry/FieldData.h"
#include "geometry/neighbouring/NeighbouringDomain.h"
#include "geometry/neighbouring/RequiredSiteInformation.h"
#include "net/IteratedAction.h"
#include "net/net.h"  // (as the reference's header does: InterfaceDelegationNet.h alone does not stand on its own)

namespace hemelb::geometry::neighbouring
{
  class NeighbouringDataManager : public net::IteratedAction
  {
    public:
      NeighbouringDataManager(const FieldData& local, NeighbouringFieldData& neighbouring, net::InterfaceDelegationNet& net) :
          localFieldData(local), neighbouringFieldData(neighbouring), net(net), needsOfOthers(net.Size())
      {
      }

      // (the requirements are not used: whole sites travel, as in the reference)
      void RegisterNeededSite(site_t globalId, RequiredSiteInformation = RequiredSiteInformation(true))
      {
        if (std::find(neededSites.begin(), neededSites.end(), globalId) == neededSites.end())
          neededSites.push_back(globalId);
      }

      // every rank tells every other how many and which of its sites it needs (collective)
      void ShareNeeds()
      {
        const int n = net.Size();
        std::vector<std::vector<site_t>> mine(n);
        std::vector<int> howMany(n, 0), howManyOfMe(n, 0);
        for (site_t id : neededSites)
        {
          const proc_t owner = ProcForSite(id);
          mine[owner].push_back(id);
          ++howMany[owner];
        }
        net.RequestAllToAllSend(howMany);
        net.RequestAllToAllReceive(howManyOfMe);
        net.Dispatch();
        for (proc_t other = 0; other < n; ++other)
        {
          net.RequestSendV(std::span<const site_t>(mine[other]), other);
          needsOfOthers[other].resize(howManyOfMe[other]);
          net.RequestReceiveV(std::span<site_t>(needsOfOthers[other]), other);
        }
        net.Dispatch();
        shared = true;
      }

      std::vector<site_t>& GetNeedsForProc(proc_t proc) { return needsOfOthers[proc]; }
      std::vector<site_t>& GetNeededSites() { return neededSites; }
      bool NeedsShared() const { return shared; }

      // The reference ships the served sites' SiteData, wall distances and normals once
      // (NeighbouringDataManager.cc:46-93) for GuoZhengShiLink to read from NeighbouringDomain.  The
      // device streamer needs none of it (a link that extrapolates reads the neighbour's f_old only).
      void TransferNonFieldDependentInformation() {}

      void TransferFieldDependentInformation() { RequestComms(); }

      virtual proc_t ProcForSite(site_t site)
      {
        return localFieldData.GetDomain().ProcProvidingSiteByGlobalNoncontiguousId(site);
      }

      // phase 0 of every time step (SimBuilder.h:160 registers this action there)
      void RequestComms() override
      {
        if (hlb_gpu_t engine = const_cast<FieldData&>(localFieldData).EngineIfBuilt())
          FieldData::Check(hlb_gpu_exchange_site_halo(engine));
        // before the first step the engine does not exist yet: the streamers build it with these
        // lists and run the first exchange themselves
      }

    private:
      const FieldData& localFieldData;
      NeighbouringFieldData& neighbouringFieldData;
      net::InterfaceDelegationNet& net;
      std::vector<site_t> neededSites;
      std::vector<std::vector<site_t>> needsOfOthers;
      bool shared = false;
  };
}
#endif
