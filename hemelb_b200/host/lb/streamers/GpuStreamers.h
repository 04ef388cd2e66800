// lb/streamers/GpuStreamers.h -- lb::streamer policy classes that run on the B200 engine.
//
// Drop-in for the reference's streamer policies (Code/lb/streamers/BulkStreamer.h:57-99,
// StreamerTypeFactory.h:24-109): same constructor (InitParams&), same
//   StreamAndCollide(first, count, lbmParams, latDat, propertyCache) / PostStep(...)
// so `Traits<LATTICE, KERNEL, Normal, GpuBulk, GpuWall<W>::type, GpuInlet<I>::type, ...>` feeds
// lb::LBM<Traits> unchanged.  Each class forwards (slot, first, count) to the C ABI; the work
// happens in hand-written sm_100a kernels.  Policy names follow CMake/HemeLbOptions.cmake.
#ifndef HEMELB_LB_STREAMERS_GPUSTREAMERS_H
#define HEMELB_LB_STREAMERS_GPUSTREAMERS_H

#include <algorithm>
#include <cmath>
#include <limits>
#include <type_traits>
#include <vector>

#include "geometry/FieldData.h"
#include "geometry/neighbouring/NeighbouringDataManager.h"
#include "geometry/neighbouring/RequiredSiteInformation.h"
#include "lb/concepts.h"
#include "lb/LbmParameters.h"
#include "lb/MacroscopicPropertyCache.h"
#include "lb/iolets/BoundaryValues.h"
#include "lb/iolets/InOutLetCosine.h"
#include "lb/iolets/InOutLetParabolicVelocity.h"
#include "lb/iolets/InOutLetVelocity.h"
#include "lb/kernels/LBGK.h"
#include "lb/kernels/MRT.h"
#include "lb/kernels/TRT.h"
#include "lb/streamers/Common.h"

namespace hemelb::lb::gpu {

  // ---- compile-time policy -> C ABI enums ------------------------------------------------------
  template <class K> struct kernel_id;
  template <lattice_type L> struct kernel_id<LBGK<L>> { static constexpr int value = HLB_KERNEL_LBGK; };
  template <moment_basis M> struct kernel_id<MRT<M>> { static constexpr int value = HLB_KERNEL_MRT; };
  template <lattice_type L> struct kernel_id<TRT<L>> { static constexpr int value = HLB_KERNEL_TRT; };

  struct SimpleBounceBack { static constexpr int value = HLB_WALL_SIMPLEBOUNCEBACK; };
  struct BouzidiFirdaousLallemand { static constexpr int value = HLB_WALL_BFL; };
  struct GuoZhengShi { static constexpr int value = HLB_WALL_GZS; };
  struct NashZerothOrderPressure { static constexpr int value = HLB_IOLET_NASHZEROTHORDERPRESSURE; };
  struct LaddIolet { static constexpr int value = HLB_IOLET_LADD; };
  struct NoLink { static constexpr int value = -1; };

  inline uint32_t CacheMask(MacroscopicPropertyCache& c) {  // SimulationMaster.impl.h:223-241
    uint32_t m = 0;
    if (c.densityCache.RequiresRefresh()) m |= HLB_CACHE_DENSITY;
    if (c.velocityCache.RequiresRefresh()) m |= HLB_CACHE_VELOCITY;
    if (c.wallShearStressMagnitudeCache.RequiresRefresh()) m |= HLB_CACHE_WALL_SHEAR_STRESS;
    if (c.vonMisesStressCache.RequiresRefresh()) m |= HLB_CACHE_VON_MISES_STRESS;
    if (c.shearRateCache.RequiresRefresh()) m |= HLB_CACHE_SHEAR_RATE;
    if (c.stressTensorCache.RequiresRefresh()) m |= HLB_CACHE_STRESS_TENSOR;
    if (c.tractionCache.RequiresRefresh()) m |= HLB_CACHE_TRACTION;
    if (c.tangentialProjectionTractionCache.RequiresRefresh()) m |= HLB_CACHE_TANGENTIAL_TRACTION;
    return m;
  }

  // One streamer class for all six LBM slots: SLOT 0 mid-fluid, 1 wall, 2 inlet, 3 outlet,
  // 4 inlet-wall, 5 outlet-wall (Code/lb/lb.h:102-107).
  template <collision_type C, int SLOT, class WALL, class IOLET>
  class GpuStreamer {
  public:
    using CollisionType = C;
    using KernelType = typename C::KernelType;
    using LatticeType = typename C::LatticeType;
    using VarsType = typename C::VarsType;

    explicit GpuStreamer(InitParams& ip) {
      auto& pol = geometry::GpuPolicyFor(ip.latDat);
      pol.kernel = kernel_id<KernelType>::value;
      pol.tau = ip.lbmParams->GetTau();
      if constexpr (WALL::value >= 0) pol.wall = WALL::value;
      if constexpr (IOLET::value >= 0) {
        if (SLOT == 2 || SLOT == 4) { pol.inlet = IOLET::value; pol.inletValues = ip.boundaryObject; }
        else { pol.outlet = IOLET::value; pol.outletValues = ip.boundaryObject; }
      }
      if constexpr (WALL::value == HLB_WALL_GZS) RegisterRemoteNeeds(ip, pol);
    }

    void StreamAndCollide(const site_t first, const site_t count, const LbmParameters*,
                          geometry::FieldData& latDat, MacroscopicPropertyCache& cache) {
      hlb_gpu_t h = latDat.Engine();
      PushStepScalars(h, latDat, cache);
      geometry::FieldData::Check(hlb_gpu_stream_and_collide(h, SLOT, first, count));
      // the last domain-edge range of LBM::PreSend (lb.hpp:176-212) releases the halo send: the
      // outlet-wall streamer's first call of a step (its second is LBM::PreReceive's, lb.hpp:214-255)
      if (SLOT == 5 && (latDat.Policy().lastSlotStreams++ & 1u) == 0)
        geometry::FieldData::Check(hlb_gpu_edge_done(h));
    }

    void PostStep(const site_t first, const site_t count, const LbmParameters*, geometry::FieldData& latDat,
                  MacroscopicPropertyCache& cache) {
      geometry::FieldData::Check(hlb_gpu_post_step(latDat.Engine(), SLOT, first, count));
      // after the last PostStep of LBM::PostReceive (lb.hpp:257-309: domain-edge ranges, then the
      // mid-domain ones) the refreshed caches are brought to the host
      if (SLOT == 5 && (latDat.Policy().lastSlotPostSteps++ & 1u) == 1) PullCaches(latDat, cache);
    }

  private:
    // What the constructor of the reference's GuoZhengShiLink does (GuoZhengShi.h:36-104): for every
    // wall link of the sites this streamer owns whose opposite direction is neither wall nor iolet, and
    // whose neighbour in that direction lives on another rank, register that site with the
    // NeighbouringDataManager.  The links themselves are kept as well: they are the rows of
    // hlb_gpu_set_gzs_remote.
    static void RegisterRemoteNeeds(InitParams& ip, geometry::GpuPolicy& pol) {
      if (!ip.neighbouringDataManager) return;  // one rank: nothing can be remote
      pol.gzsManager = ip.neighbouringDataManager;
      for (auto [first, last] : ip.siteRanges)
        for (site_t i = first; i < last; ++i) {
          auto site = ip.latDat->GetSite(i);
          if (!site.IsWall()) continue;
          auto const& here = site.GetGlobalSiteCoords();
          for (Direction d = 1; d < LatticeType::NUMVECTORS; ++d) {
            if (!site.HasWall(d)) continue;
            const Direction opp = LatticeType::INVERSEDIRECTIONS[d];
            if (site.HasWall(opp) || site.HasIolet(opp)) continue;
            const LatticeVector there = here + LatticeVector(LatticeType::CX[opp], LatticeType::CY[opp], LatticeType::CZ[opp]);
            const proc_t owner = ip.latDat->GetProcIdFromGlobalCoords(there);
            if (owner == SITE_OR_BLOCK_SOLID || owner == ip.latDat->GetLocalRank()) continue;
            const site_t globalId = ip.latDat->GetGlobalNoncontiguousSiteIdFromGlobalCoords(there);
            geometry::neighbouring::RequiredSiteInformation wanted(false);
            wanted.Require(geometry::neighbouring::terms::Density);
            wanted.Require(geometry::neighbouring::terms::Velocity);
            ip.neighbouringDataManager->RegisterNeededSite(globalId, wanted);
            pol.gzsLinks.push_back({i, (int)opp, globalId});
          }
        }
    }
    static void PushStepScalars(hlb_gpu_t h, geometry::FieldData& latDat, MacroscopicPropertyCache& cache) {
      auto& pol = latDat.Policy();
      std::vector<double> in, out;
      // by the index the sites carry: BoundaryValues::GetBoundaryDensity(id) = iolets[id]
      if (pol.inletValues)
        for (unsigned i = 0; i < pol.inletValues->GetGlobalIoletCount(); ++i) in.push_back(pol.inletValues->GetBoundaryDensity(i));
      if (pol.outletValues)
        for (unsigned i = 0; i < pol.outletValues->GetGlobalIoletCount(); ++i) out.push_back(pol.outletValues->GetBoundaryDensity(i));
      const auto t = pol.inletValues ? pol.inletValues->GetTimeStep() : (pol.outletValues ? pol.outletValues->GetTimeStep() : 1);
      const uint32_t mask = CacheMask(cache) | (pol.monitorRequested ? (uint32_t)HLB_CACHE_MONITOR : 0u);
      if (pol.scalarsPushed && pol.pushedStep == t && pol.pushedMask == mask && pol.pushedIn == in && pol.pushedOut == out)
        return;  // same step, same values: the engine already has them
      geometry::FieldData::Check(hlb_gpu_set_step_scalars(h, t, in.data(), out.data(), mask));
      pol.scalarsPushed = true;
      pol.pushedStep = t;
      pol.pushedMask = mask;
      pol.pushedIn.swap(in);
      pol.pushedOut.swap(out);
    }
    static void PullCaches(geometry::FieldData& latDat, MacroscopicPropertyCache& cache) {
      hlb_gpu_t h = latDat.Engine();
      const site_t n = latDat.GetDomain().GetLocalFluidSiteCount();
      std::vector<double> buf;
      auto scalar = [&](auto& c, uint32_t bit) {
        if (!c.RequiresRefresh()) return;
        buf.resize(n);
        geometry::FieldData::Check(hlb_gpu_get_cache(h, bit, buf.data()));
        for (site_t i = 0; i < n; ++i) c.Put(i, buf[i]);
      };
      auto vec = [&](auto& c, uint32_t bit) {
        if (!c.RequiresRefresh()) return;
        buf.resize(3 * n);
        geometry::FieldData::Check(hlb_gpu_get_cache(h, bit, buf.data()));
        for (site_t i = 0; i < n; ++i) c.Put(i, util::Vector3D<distribn_t>(buf[3 * i], buf[3 * i + 1], buf[3 * i + 2]));
      };
      scalar(cache.densityCache, HLB_CACHE_DENSITY);
      vec(cache.velocityCache, HLB_CACHE_VELOCITY);
      scalar(cache.wallShearStressMagnitudeCache, HLB_CACHE_WALL_SHEAR_STRESS);
      scalar(cache.vonMisesStressCache, HLB_CACHE_VON_MISES_STRESS);
      scalar(cache.shearRateCache, HLB_CACHE_SHEAR_RATE);
      vec(cache.tractionCache, HLB_CACHE_TRACTION);
      vec(cache.tangentialProjectionTractionCache, HLB_CACHE_TANGENTIAL_TRACTION);
      if (cache.stressTensorCache.RequiresRefresh()) {
        buf.resize(9 * n);
        geometry::FieldData::Check(hlb_gpu_get_cache(h, HLB_CACHE_STRESS_TENSOR, buf.data()));
        for (site_t i = 0; i < n; ++i) {
          util::Matrix3D m;
          for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) m[a][b] = buf[9 * i + 3 * a + b];
          cache.stressTensorCache.Put(i, m);
        }
      }
    }
  };

  // Traits-ready aliases: STREAMER, WALL_BOUNDARY, INLET_BOUNDARY, OUTLET_BOUNDARY template
  // template parameters of hemelb::Traits (Code/Traits.h:17-39)
  template <class C> using Bulk = GpuStreamer<C, 0, NoLink, NoLink>;
  template <class W> struct Wall { template <class C> using type = GpuStreamer<C, 1, W, NoLink>; };
  template <class I> struct Inlet { template <class C> using type = GpuStreamer<C, 2, NoLink, I>; };
  template <class I> struct Outlet { template <class C> using type = GpuStreamer<C, 3, NoLink, I>; };
}

namespace hemelb::lb {
  // primary template lives in Code/lb/Streamers.h:71-74 (re-declared so this header stands alone)
  template <typename WS, typename IS> struct CombineWallAndIoletStreamers;
  // wall + iolet combination (Code/lb/Streamers.h:71-99): inlet-wall is slot 4, outlet-wall slot 5
  template <class C, class W, class I>
  struct CombineWallAndIoletStreamers<gpu::GpuStreamer<C, 1, W, gpu::NoLink>, gpu::GpuStreamer<C, 2, gpu::NoLink, I>> {
    using type = gpu::GpuStreamer<C, 4, W, I>;
  };
  template <class C, class W, class I>
  struct CombineWallAndIoletStreamers<gpu::GpuStreamer<C, 1, W, gpu::NoLink>, gpu::GpuStreamer<C, 3, gpu::NoLink, I>> {
    using type = gpu::GpuStreamer<C, 5, W, I>;
  };
}

namespace hemelb::geometry {
  // Build the device engine from the Domain's tables the first time a streamer needs it.
  inline void FieldData::EnsureEngine() {
    if (m_gpu) return;
    Domain& d = *m_domain;
    GpuPolicy& pol = Policy();
    if (pol.kernel < 0) throw Exception() << "hemelb_b200: no Gpu streamer was constructed for this Domain";
    const int Q = d.latticeInfo.GetNumVectors();
    const site_t N = d.GetLocalFluidSiteCount();
    hlb_gpu_config cfg{};
    cfg.lattice = Q;
    cfg.kernel = pol.kernel;
    cfg.wall = pol.wall < 0 ? HLB_WALL_SIMPLEBOUNCEBACK : pol.wall;
    cfg.inlet = pol.inlet < 0 ? HLB_IOLET_NASHZEROTHORDERPRESSURE : pol.inlet;
    cfg.outlet = pol.outlet < 0 ? HLB_IOLET_NASHZEROTHORDERPRESSURE : pol.outlet;
    cfg.tau = pol.tau;
    cfg.rank = d.GetLocalRank();
    cfg.nranks = d.GetCommunicator().Size();
    int ndev = 1;
    Check(hlb_gpu_device_count(&ndev));
    cfg.device = cfg.rank % ndev;  // rank r -> GPU r on the NVSwitch box
    cfg.n_sites = N;
    for (unsigned t = 0; t < COLLISION_TYPES; ++t) {
      cfg.mid_count[t] = d.GetMidDomainCollisionCount(t);
      cfg.edge_count[t] = d.GetDomainEdgeCollisionCount(t);
    }
    cfg.total_shared_fs = d.totalSharedFs;
    cfg.reorder = 1;
    cfg.n_neighbours = (int)d.neighbouringProcs.size();
    // LBM::PrepareBoundaryObjects (lb.hpp:128-152): the minimum density over every iolet
    double minDensity = std::numeric_limits<double>::max();
    for (lb::BoundaryValues* bv : {pol.inletValues, pol.outletValues})
      if (bv)
        for (unsigned i = 0; i < bv->GetLocalIoletCount(); ++i)
          minDensity = std::min(minDensity, (double)bv->GetLocalIolet(i)->GetDensityMin());
    // One record per iolet of the SIMULATION, in the order of the BoundaryValues' own list: the id a
    // site carries (SiteData::GetIoletId) indexes that list -- BoundaryValues::GetBoundaryDensity(id)
    // reads iolets[id] (BoundaryValues.cc:162-165).  (The reference's link streamers take the
    // iolet's geometry from GetLocalIolet(id) = iolets[localIoletIDs[id]], NashZerothOrderPressure.h:39,
    // LaddIolet.h:46: the same object whenever the rank holds iolets 0..id, out of range or another
    // iolet otherwise; the device table holds iolets[id].)
    const bool velocityLinks = pol.wall == HLB_WALL_GZS;
    auto records = [minDensity, velocityLinks](lb::BoundaryValues* bv, int linkPolicy) {
      std::vector<double> r;
      if (!bv) return r;
      for (unsigned i = 0; i < bv->GetGlobalIoletCount(); ++i) {
        lb::InOutLet* io = bv->GetGlobalIolet(i);
        double rec[HLB_IOLET_RECORD_DOUBLES] = {0};
        auto const& n = io->GetNormal();
        auto const& p = io->GetPosition();
        for (int k = 0; k < 3; ++k) { rec[1 + k] = n[k]; rec[4 + k] = p[k]; }
        if (auto* v = dynamic_cast<lb::InOutLetParabolicVelocity*>(io)) {
          rec[0] = 1; rec[7] = v->GetRadius(); rec[8] = v->GetMaxSpeed();
          // the warm-up length has a setter only (InOutLetParabolicVelocity.h:29-36): read it off the
          // ramp it produces, max * t / warmUpLength at the centre line for t = 1 < warmUpLength
          // (InOutLetParabolicVelocity.cc:33-39); without a ramp at t = 1 there is none at any step
          if (v->GetMaxSpeed() != 0.0) {
            const double at1 = v->GetVelocity(v->GetPosition(), 1).GetMagnitude();
            const double ratio = std::abs(v->GetMaxSpeed()) / at1;
            if (at1 > 0.0 && ratio > 1.5) rec[13] = std::floor(ratio + 0.5);
          }
        } else if (dynamic_cast<lb::InOutLetVelocity*>(io)) {
          // a velocity iolet whose profile the device does not evaluate (Womersley, file): the Ladd
          // and GuoZhengShi links would silently apply zero velocity
          if (linkPolicy == HLB_IOLET_LADD || velocityLinks)
            throw Exception() << "hemelb_b200: velocity iolet " << i << " is not an InOutLetParabolicVelocity; the "
                                 "device evaluates only the parabolic profile (LaddIolet.h:45-63, GuoZhengShi.h:163-189)";
          rec[0] = 1;
        } else if (auto* c = dynamic_cast<lb::InOutLetCosine*>(io)) {
          rec[9] = c->GetDensityMean(); rec[10] = c->GetDensityAmp(); rec[11] = c->GetPhase(); rec[12] = c->GetPeriod();
        } else {
          rec[9] = io->GetDensityMin(); rec[12] = 1.0;  // densities still arrive per step from BoundaryValues
        }
        rec[14] = minDensity;
        r.insert(r.end(), rec, rec + HLB_IOLET_RECORD_DOUBLES);
      }
      return r;
    };
    auto rin = records(pol.inletValues, pol.inlet), rout = records(pol.outletValues, pol.outlet);
    cfg.n_inlets = (int)(rin.size() / HLB_IOLET_RECORD_DOUBLES);
    cfg.n_outlets = (int)(rout.size() / HLB_IOLET_RECORD_DOUBLES);
    Check(hlb_gpu_create(&cfg, &m_gpu));
    pol.engine = m_gpu;
    Check(hlb_gpu_set_neighbour_indices(m_gpu, 0, N, d.neighbourIndices.data()));
    std::vector<uint32_t> wall(N), iol(N);
    std::vector<int32_t> ioid(N);
    std::vector<double> normals(3 * N);
    std::vector<int64_t> coords(3 * N);
    for (site_t i = 0; i < N; ++i) {
      auto const& sd = d.GetSiteData(i);
      wall[i] = sd.GetWallIntersectionData();
      iol[i] = sd.GetIoletIntersectionData();
      ioid[i] = sd.GetIoletId();
      for (int k = 0; k < 3; ++k) { normals[3 * i + k] = d.GetNormalToWall(i)[k]; coords[3 * i + k] = d.GetGlobalSiteCoords(i)[k]; }
    }
    // only the boundary-typed id ranges carry cut links: [mid0, midTotal) and [midTotal+edge0, N)
    const site_t midTotal = d.GetMidDomainSiteCount();
    const site_t ranges[2][2] = {{d.GetMidDomainCollisionCount(0), midTotal},
                                 {midTotal + d.GetDomainEdgeCollisionCount(0), N}};
    for (auto& r : ranges) {
      const site_t a = r[0], n = r[1] - r[0];
      if (n <= 0) continue;
      Check(hlb_gpu_set_site_data(m_gpu, a, n, wall.data() + a, iol.data() + a, ioid.data() + a));
      Check(hlb_gpu_set_wall_distances(m_gpu, a, n, d.distanceToWall.data() + a * (Q - 1)));
      Check(hlb_gpu_set_wall_normals(m_gpu, a, n, normals.data() + 3 * a));
    }
    Check(hlb_gpu_set_site_coords(m_gpu, 0, N, coords.data()));  // every site: internal z-run renumbering
    std::vector<int> nr;
    std::vector<int64_t> nc, nf;
    for (auto const& p : d.neighbouringProcs) { nr.push_back(p.Rank); nc.push_back(p.SharedDistributionCount); nf.push_back(p.FirstSharedDistribution); }
    if (!nr.empty()) {
      Check(hlb_gpu_set_neighbours(m_gpu, nr.data(), nc.data(), nf.data()));
      Check(hlb_gpu_set_streaming_indices(m_gpu, d.streamingIndicesForReceivedDistributions.data()));
    }
    if (cfg.n_inlets) Check(hlb_gpu_set_iolets(m_gpu, 0, cfg.n_inlets, rin.data()));
    if (cfg.n_outlets) Check(hlb_gpu_set_iolets(m_gpu, 1, cfg.n_outlets, rout.data()));
    if (pol.gzsManager && cfg.nranks > 1) {
      // GuoZhengShi across ranks: the links registered by the streamers, grouped by the rank that owns
      // the neighbour (registration order kept inside a group: it is the order of the manager's
      // neededSites, hence of the owner's GetNeedsForProc list); what this rank serves, from the lists
      // NeighbouringDataManager::ShareNeeds gathered (SimBuilder.h:235 runs it before the first step)
      if (!pol.gzsManager->NeedsShared())
        throw Exception() << "hemelb_b200: NeighbouringDataManager::ShareNeeds has not run before the first time step";
      std::vector<GpuPolicy::GzsLink> links = pol.gzsLinks;
      std::vector<int32_t> owner(links.size());
      std::vector<size_t> order(links.size());
      for (size_t k = 0; k < links.size(); ++k) {
        owner[k] = d.ProcProvidingSiteByGlobalNoncontiguousId(links[k].globalId);
        order[k] = k;
      }
      std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return owner[a] < owner[b]; });
      std::vector<int64_t> lsite, lkey;
      std::vector<int32_t> ldir, lowner;
      for (size_t k : order) {
        lsite.push_back(links[k].site);
        ldir.push_back(links[k].direction);
        lowner.push_back(owner[k]);
        lkey.push_back(links[k].globalId);
      }
      if (!lsite.empty())
        Check(hlb_gpu_set_gzs_remote(m_gpu, (int64_t)lsite.size(), lsite.data(), ldir.data(), lowner.data(), lkey.data()));
      std::vector<int32_t> srank;
      std::vector<int64_t> ssite;
      for (proc_t other = 0; other < cfg.nranks; ++other)
        if (other != cfg.rank)
          for (site_t globalId : pol.gzsManager->GetNeedsForProc(other)) {
            srank.push_back(other);
            ssite.push_back(d.GetLocalContiguousIdFromGlobalNoncontiguousId(globalId));
          }
      if (!srank.empty()) Check(hlb_gpu_set_gzs_serve(m_gpu, (int64_t)srank.size(), srank.data(), ssite.data()));
    }
    Check(hlb_gpu_finalise(m_gpu));
    if (cfg.nranks > 1) {
      // NCCL bootstrap over the reference's own MPI communicator: rank 0 creates the id
      char id[128];
      if (cfg.rank == 0) Check(hlb_gpu_comm_unique_id(id));
      d.GetCommunicator().Broadcast(std::span<char>(id, 128), 0);
      Check(hlb_gpu_comm_init(m_gpu, id));
    }
    m_needFirstSiteHalo = pol.gzsManager != nullptr && cfg.nranks > 1;
  }
}
#endif
