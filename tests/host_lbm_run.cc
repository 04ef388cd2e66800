// Runnable check of the C++ host-side drop-in (hemelb_b200/host): the Gpu*Streamer policy classes
// and the device-backed geometry::FieldData, driven the way lb::LBM<Traits> drives its streamers
// -- InitCollisions (Code/lb/lb.hpp:75-114), then per time step RequestComms / PreSend / PreReceive
// / PostReceive / EndIteration (lb.hpp:162-314) and SimulationMaster's swap (SimulationMaster.
// impl.h:218-219) -- with the reference's own LbmParameters, SimulationState, InOutLet classes,
// SiteData and MacroscopicPropertyCache.  lb::LBM itself needs net::Net / MPI / the XML
// configuration and cannot be built here; this harness makes the calls it makes, in its order.
//
//   host_lbm_run <case.bin> <out.bin>
//
// Linked against tests/host_mock_abi.cc it records the C-ABI calls (CPU test of the call sequence);
// linked against libhemelb_b200.so it runs on the GPU and the result is compared with the oracle.
// Test infrastructure: built by __graft_entry__.build() where /root/reference exists.
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "lb/lattices/D3Q15.h"
#include "lb/lattices/D3Q19.h"
#include "lb/lattices/D3Q27.h"
#include "lb/kernels/LBGK.h"
#include "lb/kernels/MRT.h"
#include "lb/kernels/DHumieresD3Q19MRTBasis.h"
#include "lb/collisions/Normal.h"
#include "lb/iolets/BoundaryValues.h"
#include "lb/iolets/InOutLetCosine.h"
#include "lb/iolets/InOutLetParabolicVelocity.h"
#include "lb/MacroscopicPropertyCache.h"
#include "lb/SimulationState.h"
#include "lb/streamers/GpuStreamers.h"
#include "lb/StabilityTester.h"  // hemelb_b200/host: the device-side stand-in
#include "lb/IncompressibilityChecker.hpp"  // hemelb_b200/host: likewise

using namespace hemelb;
namespace g = hemelb::lb::gpu;

namespace {
  struct Case {
    int64_t head[32];
    double phys[4];  // dt, dx, rho, eta -> LbmParameters
    std::vector<int64_t> neighbourIndices, globalCoords, procs3, streamingIndices;
    std::vector<uint32_t> wallMask, ioletMask;
    std::vector<int32_t> ioletId, siteType;
    std::vector<double> distanceToWall, wallNormal, inletRec, outletRec, f0;
    std::vector<int64_t> whereIs;  // several ranks: {x, y, z, rank, local id} of every fluid site of the geometry
    int64_t extent[3] = {1, 1, 1};
    int Q() const { return (int)head[1]; }
    int64_t N() const { return head[6]; }
  };

  template <class T> void get(FILE* fh, std::vector<T>& v, size_t n) {
    v.resize(n);
    if (n && fread(v.data(), sizeof(T), n, fh) != n) throw std::runtime_error("case file truncated");
  }

  Case read_case(const char* path) {
    Case c;
    FILE* fh = fopen(path, "rb");
    if (!fh) throw std::runtime_error(std::string("cannot open ") + path);
    if (fread(c.head, sizeof(int64_t), 32, fh) != 32 || c.head[0] != 0x484C4231) throw std::runtime_error("bad case header");
    if (fread(c.phys, sizeof(double), 4, fh) != 4) throw std::runtime_error("case file truncated");
    const size_t N = c.N(), Q = c.Q(), S = c.head[19];
    get(fh, c.neighbourIndices, N * Q);
    get(fh, c.wallMask, N);
    get(fh, c.ioletMask, N);
    get(fh, c.ioletId, N);
    get(fh, c.siteType, N);
    get(fh, c.distanceToWall, N * (Q - 1));
    get(fh, c.wallNormal, 3 * N);
    get(fh, c.globalCoords, 3 * N);
    get(fh, c.inletRec, HLB_IOLET_RECORD_DOUBLES * c.head[20]);
    get(fh, c.outletRec, HLB_IOLET_RECORD_DOUBLES * c.head[21]);
    get(fh, c.f0, N * Q + 1 + S);
    get(fh, c.procs3, 3 * c.head[26]);  // {rank, SharedDistributionCount, FirstSharedDistribution} per neighbour
    get(fh, c.streamingIndices, S);
    if (c.head[27] > 0) {  // the site -> (rank, local id) table, for GuoZhengShi across ranks
      if (fread(c.extent, sizeof(int64_t), 3, fh) != 3) throw std::runtime_error("case file truncated");
      get(fh, c.whereIs, 5 * c.head[27]);
    }
    fclose(fh);
    return c;
  }
}

// fills the private tables of the test Domain (tests/host_shim/geometry/Domain.h)
struct HostDomainFiller {
  static void Fill(geometry::Domain& d, const Case& c) {
    const site_t N = c.N();
    const int Q = c.Q();
    d.nSites = N;
    for (int t = 0; t < 6; ++t) {
      d.mid[t] = c.head[7 + t];
      d.edge[t] = c.head[13 + t];
    }
    d.totalSharedFs = c.head[19];
    d.neighbourIndices.assign(c.neighbourIndices.begin(), c.neighbourIndices.end());
    d.distanceToWall = c.distanceToWall;
    d.wallNormalAtSite.resize(N);
    d.globalSiteCoords.resize(N);
    d.siteData.resize(N);
    for (site_t i = 0; i < N; ++i) {
      d.wallNormalAtSite[i] = util::Vector3D<distribn_t>(c.wallNormal[3 * i], c.wallNormal[3 * i + 1], c.wallNormal[3 * i + 2]);
      d.globalSiteCoords[i] = util::Vector3D<site_t>(c.globalCoords[3 * i], c.globalCoords[3 * i + 1], c.globalCoords[3 * i + 2]);
      // the reference's own SiteData, through its GeometrySite constructor (SiteDataBare.cc:23-73)
      geometry::GeometrySite gs(true);
      gs.links.resize(Q - 1);
      for (int dir = 1; dir < Q; ++dir) {
        using CutType = io::formats::geometry::CutType;
        auto& link = gs.links[dir - 1];
        if ((c.wallMask[i] >> (dir - 1)) & 1u) link.type = CutType::WALL;
        else if ((c.ioletMask[i] >> (dir - 1)) & 1u) {
          link.type = c.siteType[i] == 2 ? CutType::INLET : CutType::OUTLET;
          link.ioletId = c.ioletId[i];
        }
      }
      d.siteData[i] = geometry::SiteData(gs);
    }
    d.neighbouringData = std::make_shared<geometry::neighbouring::NeighbouringDomain>();
    d.comms.rank = (int)c.head[24];
    d.comms.size = (int)c.head[25];
    if (const char* f = getenv("HLB_HOST_ID_FILE")) d.comms.idFile = f;
    for (size_t p = 0; p * 3 < c.procs3.size(); ++p)
      d.neighbouringProcs.push_back({(proc_t)c.procs3[3 * p], c.procs3[3 * p + 1], c.procs3[3 * p + 2]});
    d.streamingIndicesForReceivedDistributions.assign(c.streamingIndices.begin(), c.streamingIndices.end());
    d.sites = util::Vector3D<site_t>(c.extent[0], c.extent[1], c.extent[2]);
    for (size_t k = 0; k * 5 < c.whereIs.size(); ++k) {
      const int64_t* w = &c.whereIs[5 * k];
      d.whereIs[(w[0] * c.extent[1] + w[1]) * c.extent[2] + w[2]] = {(proc_t)w[3], (site_t)w[4]};
    }
  }
};

namespace {
  void make_iolets(const std::vector<double>& rec, std::vector<std::unique_ptr<lb::InOutLet>>& store, lb::BoundaryValues& bv,
                   lb::SimulationState* state) {
    bv.state = state;
    for (size_t i = 0; i * HLB_IOLET_RECORD_DOUBLES < rec.size(); ++i) {
      const double* q = &rec[i * HLB_IOLET_RECORD_DOUBLES];
      lb::InOutLet* io;
      if ((int)q[0] == 0) {
        auto* c = new lb::InOutLetCosine();
        c->SetDensityMean(q[9]);
        c->SetDensityAmp(q[10]);
        c->SetPhase(q[11]);
        c->SetPeriod(q[12]);
        c->SetWarmup((unsigned)q[13]);
        io = c;
      } else {
        auto* v = new lb::InOutLetParabolicVelocity();
        v->SetRadius(q[7]);
        v->SetMaxSpeed(q[8]);
        v->SetWarmup((unsigned)q[13]);
        io = v;
      }
      io->SetNormal(util::Vector3D<double>(q[1], q[2], q[3]));
      io->SetPosition(LatticePosition(q[4], q[5], q[6]));
      store.emplace_back(io);
      bv.iolets.push_back(io);
    }
  }

  // The slice of lb::LBM<Traits> that touches the streamers.  The six streamer types are formed as
  // lb.h:90-107 forms them from the Traits' STREAMER / WALL_BOUNDARY / INLET_BOUNDARY /
  // OUTLET_BOUNDARY template template parameters.
  template <class COLLISION, class WALL, class INLET, class OUTLET>
  struct HostLBM {
    using tMidFluid = g::Bulk<COLLISION>;
    using tWall = typename g::Wall<WALL>::template type<COLLISION>;
    using tInlet = typename g::Inlet<INLET>::template type<COLLISION>;
    using tOutlet = typename g::Outlet<OUTLET>::template type<COLLISION>;
    using tInletWall = typename lb::CombineWallAndIoletStreamers<tWall, tInlet>::type;
    using tOutletWall = typename lb::CombineWallAndIoletStreamers<tWall, tOutlet>::type;

    geometry::FieldData* latDat;
    lb::LbmParameters params;
    lb::BoundaryValues *inletValues, *outletValues;
    lb::MacroscopicPropertyCache& cache;
    std::unique_ptr<tMidFluid> midFluid;
    std::unique_ptr<tWall> wall;
    std::unique_ptr<tInlet> inlet;
    std::unique_ptr<tOutlet> outlet;
    std::unique_ptr<tInletWall> inletWall;
    std::unique_ptr<tOutletWall> outletWall;

    HostLBM(geometry::FieldData* fd, const lb::LbmParameters& p, lb::BoundaryValues* in, lb::BoundaryValues* out,
            lb::MacroscopicPropertyCache& c, geometry::neighbouring::NeighbouringDataManager* ndm = nullptr) :
        latDat(fd), params(p), inletValues(in), outletValues(out), cache(c) {
      PrepareBoundaryObjects();
      // LBM::InitCollisions (lb.hpp:75-114): every streamer is told its mid-domain and its
      // domain-edge range
      auto& dom = fd->GetDomain();
      lb::InitParams ip;
      ip.latDat = &dom;
      ip.lbmParams = &params;
      ip.neighbouringDataManager = ndm;
      ip.siteRanges.resize(2);
      ip.siteRanges[0].first = 0;
      ip.siteRanges[1].first = dom.GetMidDomainSiteCount();
      auto advance = [&](unsigned t) {
        ip.siteRanges[0].second = ip.siteRanges[0].first + dom.GetMidDomainCollisionCount(t);
        ip.siteRanges[1].second = ip.siteRanges[1].first + dom.GetDomainEdgeCollisionCount(t);
      };
      auto next = [&]() { ip.siteRanges[0].first = ip.siteRanges[0].second; ip.siteRanges[1].first = ip.siteRanges[1].second; };
      ip.boundaryObject = nullptr;
      advance(0);
      midFluid = std::make_unique<tMidFluid>(ip);
      next(); advance(1);
      wall = std::make_unique<tWall>(ip);
      next(); advance(2);
      ip.boundaryObject = inletValues;
      inlet = std::make_unique<tInlet>(ip);
      next(); advance(3);
      ip.boundaryObject = outletValues;
      outlet = std::make_unique<tOutlet>(ip);
      next(); advance(4);
      ip.boundaryObject = inletValues;
      inletWall = std::make_unique<tInletWall>(ip);
      next(); advance(5);
      ip.boundaryObject = outletValues;
      outletWall = std::make_unique<tOutletWall>(ip);
    }

    void PrepareBoundaryObjects() {  // lb.hpp:128-152
      distribn_t lowest = std::numeric_limits<distribn_t>::max();
      for (auto* bv : {inletValues, outletValues})
        for (unsigned i = 0; i < bv->GetLocalIoletCount(); ++i) lowest = std::min(lowest, bv->GetLocalIolet(i)->GetDensityMin());
      for (auto* bv : {inletValues, outletValues})
        for (unsigned i = 0; i < bv->GetLocalIoletCount(); ++i) bv->GetLocalIolet(i)->SetMinimumSimulationDensity(lowest);
    }

    template <class F> void SixRanges(bool edge, F&& call) {
      auto& dom = latDat->GetDomain();
      site_t offset = edge ? dom.GetMidDomainSiteCount() : 0;
      auto count = [&](unsigned t) { return edge ? dom.GetDomainEdgeCollisionCount(t) : dom.GetMidDomainCollisionCount(t); };
      call(*midFluid, offset, count(0)); offset += count(0);
      call(*wall, offset, count(1)); offset += count(1);
      call(*inlet, offset, count(2)); offset += count(2);
      call(*outlet, offset, count(3)); offset += count(3);
      call(*inletWall, offset, count(4)); offset += count(4);
      call(*outletWall, offset, count(5));
    }
    void RequestComms() { latDat->SendAndReceive(nullptr); }
    void PreSend() { SixRanges(true, [&](auto& s, site_t a, site_t n) { s.StreamAndCollide(a, n, &params, *latDat, cache); }); }
    void PreReceive() { SixRanges(false, [&](auto& s, site_t a, site_t n) { s.StreamAndCollide(a, n, &params, *latDat, cache); }); }
    void PostReceive() {
      latDat->CopyReceived();
      SixRanges(true, [&](auto& s, site_t a, site_t n) { s.PostStep(a, n, &params, *latDat, cache); });
      SixRanges(false, [&](auto& s, site_t a, site_t n) { s.PostStep(a, n, &params, *latDat, cache); });
    }
    void EndIteration() {}
  };

  template <class COLLISION, class WALL, class INLET, class OUTLET>
  int run(const Case& c, const char* outPath) {
    using Lattice = typename COLLISION::LatticeType;
    const site_t N = c.N();
    const int Q = c.Q();
    auto dom = std::make_shared<geometry::Domain>(Lattice::GetLatticeInfo());
    HostDomainFiller::Fill(*dom, c);
    geometry::FieldData fd(dom);
    lb::SimulationState state{c.phys[0], 1000000000ul};
    lb::LbmParameters params(c.phys[0], c.phys[1], c.phys[2], c.phys[3]);
    std::vector<std::unique_ptr<lb::InOutLet>> inStore, outStore;
    lb::BoundaryValues inletValues, outletValues;
    make_iolets(c.inletRec, inStore, inletValues, &state);
    make_iolets(c.outletRec, outStore, outletValues, &state);
    lb::MacroscopicPropertyCache cache(state, *dom);
    // several ranks: the NeighbouringDataManager (hemelb_b200/host's stand-in) over a stand-in net
    // between the harness processes -- SimBuilder.h:153-160 constructs it, :235 shares the needs
    std::unique_ptr<net::InterfaceDelegationNet> ndmNet;
    std::unique_ptr<geometry::neighbouring::NeighbouringDataManager> ndm;
    // (only where a streamer registers needs: ShareNeeds is collective, and the call-sequence tests run
    // their ranks one after the other)
    if (c.head[25] > 1 && getenv("HLB_HOST_ID_FILE") && std::is_same_v<WALL, g::GuoZhengShi>) {
      ndmNet = std::make_unique<net::InterfaceDelegationNet>((int)c.head[24], (int)c.head[25], getenv("HLB_HOST_ID_FILE"));
      ndm = std::make_unique<geometry::neighbouring::NeighbouringDataManager>(fd, fd.GetNeighbouringData(), *ndmNet);
    }
    HostLBM<COLLISION, WALL, INLET, OUTLET> lbm(&fd, params, &inletValues, &outletValues, cache, ndm.get());
    if (ndm) {
      ndm->ShareNeeds();
      ndm->TransferNonFieldDependentInformation();
    }

    // the initial condition is written through the host view, as lb::InitialCondition does
    for (size_t i = 0; i < c.f0.size(); ++i) {
      *fd.GetFOld(i) = c.f0[i];
      *fd.GetFNew(i) = c.f0[i];
    }
    const int64_t steps = c.head[22];
    const unsigned want = (unsigned)c.head[23];
    // HLB_HOST_TIMING=1: wall-clock MLUPS of steps 2..K as this host drives them (stderr); step 1
    // builds the engine and uploads the tables
    const bool timing = getenv("HLB_HOST_TIMING") != nullptr && steps > 1;
    // HLB_HOST_STABILITY=1 (2: with the velocity convergence check): an lb::StabilityTester assesses
    // every step, after the streaming and before the swap, as SimulationMaster's step manager runs it
    const int stabilityMode = getenv("HLB_HOST_STABILITY") ? atoi(getenv("HLB_HOST_STABILITY")) : 0;
    reporting::Timers timers;
    configuration::MonitoringConfig monitoring;
    monitoring.doConvergenceCheck = stabilityMode == 2;
    monitoring.convergenceVariable = extraction::source::Velocity{};
    monitoring.convergenceReferenceValue = 0.01;
    monitoring.convergenceRelativeTolerance = 1e-9;
    std::unique_ptr<lb::StabilityTester<Lattice>> tester;
    if (stabilityMode)
      tester = std::make_unique<lb::StabilityTester<Lattice>>(
          std::shared_ptr<const geometry::FieldData>(&fd, [](const geometry::FieldData*) {}), nullptr, &state, timers,
          monitoring);
    // HLB_HOST_INCOMPRESSIBILITY=1: an lb::IncompressibilityChecker beside it, constructed as
    // configuration/SimBuilder.h:216-226 constructs it (with the Domain and the property cache)
    using Checker = lb::IncompressibilityChecker<net::PhasedBroadcastRegular<>>;
    std::unique_ptr<Checker> checker;
    if (getenv("HLB_HOST_INCOMPRESSIBILITY"))
      checker = std::make_unique<Checker>(&fd.GetDomain(), nullptr, &state, cache, timers, 0.05);
    auto t0 = std::chrono::steady_clock::now();
    for (int64_t s = 0; s < steps; ++s) {
      if (timing && s == 1) {
        geometry::FieldData::Check(hlb_gpu_sync(fd.Engine()));
        t0 = std::chrono::steady_clock::now();
      }
      cache.ResetRequirements();
      if (s == steps - 1) {  // a PropertyActor asking for output on the last step
        if (want & 1) cache.densityCache.SetRefreshFlag();
        if (want & 2) cache.velocityCache.SetRefreshFlag();
      }
      if (ndm) ndm->RequestComms();  // phase 0 (SimBuilder.h:160)
      lbm.RequestComms();
      lbm.PreSend();
      lbm.PreReceive();
      lbm.PostReceive();
      lbm.EndIteration();
      if (tester) {
        tester->RunCycle();
        fprintf(stderr, "host_lbm_run: step %lld stability %d\n", (long long)s, (int)state.GetStability());
      }
      if (checker) {
        checker->RunCycle();
        fprintf(stderr, "host_lbm_run: step %lld densities %.17g %.17g speed %.17g within %d\n", (long long)s,
                checker->GetGlobalSmallestDensity(), checker->GetGlobalLargestDensity(),
                checker->GetGlobalLargestVelocityMagnitude(), (int)checker->IsDensityDiffWithinRange());
      }
      fd.SwapOldAndNew();
      state.Increment();
    }
    if (timing) {
      geometry::FieldData::Check(hlb_gpu_sync(fd.Engine()));
      const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      fprintf(stderr, "host_lbm_run: %lld sites, %lld timed steps, %.6f s, %.1f MLUPS\n", (long long)N,
              (long long)(steps - 1), dt, N * double(steps - 1) / dt / 1e6);
    }
    FILE* fh = fopen(outPath, "wb");
    if (!fh) throw std::runtime_error("cannot write the result");
    const distribn_t* f = const_cast<geometry::FieldData const&>(fd).GetFOld(0);
    fwrite(f, sizeof(double), N * Q, fh);
    if (want & 1) for (site_t i = 0; i < N; ++i) { double v = cache.densityCache.Get(i); fwrite(&v, sizeof(double), 1, fh); }
    if (want & 2) for (site_t i = 0; i < N; ++i) { auto v = cache.velocityCache.Get(i); double w[3] = {v[0], v[1], v[2]}; fwrite(w, sizeof(double), 3, fh); }
    fclose(fh);
    return 0;
  }
}

// defined by each HemeLB executable (util/Vector3D.h:28-33)
hemelb::util::Vector3DBase::HandlerFunction* hemelb::util::Vector3DBase::handler = nullptr;

// link-time stub: lb/SimulationState.cc references reporting::Dict (ctemplate wrapper)
namespace hemelb::reporting {
  Dict::Dict(const std::string&) : raw(nullptr, [](ctemplate::TemplateDictionary*) {}) {}
  Dict::Dict(ctemplate::TemplateDictionary*) : raw(nullptr, [](ctemplate::TemplateDictionary*) {}) {}
  Dict Dict::AddSectionDictionary(const std::string&) { return Dict(std::string()); }
  void Dict::SetValue(const std::string&, const std::string&) {}
  void Dict::SetIntValue(const std::string&, long) {}
  void Dict::SetBoolValue(const std::string&, bool) {}
  template <typename T> void Dict::SetFormattedValue(const std::string&, const char*, const T&) {}
  template void Dict::SetFormattedValue<double>(const std::string&, const char*, const double&);
}

int main(int argc, char** argv) {
  if (argc != 3) {
    fprintf(stderr, "usage: host_lbm_run <case.bin> <out.bin>\n");
    return 2;
  }
  try {
    const Case c = read_case(argv[1]);
    const int Q = c.Q(), kernel = (int)c.head[2], wall = (int)c.head[3], in = (int)c.head[4], out = (int)c.head[5];
    using LBGK15 = lb::Normal<lb::LBGK<lb::D3Q15>>;
    using LBGK19 = lb::Normal<lb::LBGK<lb::D3Q19>>;
    using MRT19 = lb::Normal<lb::MRT<lb::DHumieresD3Q19MRTBasis>>;
    using LBGK27 = lb::Normal<lb::LBGK<lb::D3Q27>>;
    if (Q == 15 && kernel == 0 && wall == 0 && in == 0 && out == 0)
      return run<LBGK15, g::SimpleBounceBack, g::NashZerothOrderPressure, g::NashZerothOrderPressure>(c, argv[2]);
    if (Q == 19 && kernel == 0 && wall == 1 && in == 0 && out == 0)
      return run<LBGK19, g::BouzidiFirdaousLallemand, g::NashZerothOrderPressure, g::NashZerothOrderPressure>(c, argv[2]);
    if (Q == 19 && kernel == 1 && wall == 1 && in == 1 && out == 0)
      return run<MRT19, g::BouzidiFirdaousLallemand, g::LaddIolet, g::NashZerothOrderPressure>(c, argv[2]);
    if (Q == 19 && kernel == 0 && wall == 2 && in == 1 && out == 0)
      return run<LBGK19, g::GuoZhengShi, g::LaddIolet, g::NashZerothOrderPressure>(c, argv[2]);
    if (Q == 27 && kernel == 0 && wall == 0 && in == 0 && out == 0)
      return run<LBGK27, g::SimpleBounceBack, g::NashZerothOrderPressure, g::NashZerothOrderPressure>(c, argv[2]);
    fprintf(stderr, "policy combination not instantiated in this harness\n");
    return 2;
  } catch (std::exception& e) {
    fprintf(stderr, "host_lbm_run: %s\n", e.what());
    return 1;
  }
}
