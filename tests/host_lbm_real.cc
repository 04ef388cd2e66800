// The reference's own lb::LBM<Traits> (Code/lb/lb.h, lb.hpp -- unmodified), constructed and stepped as
// SimulationMaster does, over the reference's own geometry::Domain (Code/geometry/Domain.cc), net::Net,
// lb::BoundaryValues, SimulationState, LbmParameters and reporting::Timers -- all compiled unmodified --
// with hemelb_b200/host's policy classes named in the Traits and its device-backed geometry::FieldData
// in place of the reference's.  The R ranks are threads of this process (oracle/fake_mpi.cc stands in
// for MPI, as for oracle/_ref/libhemelb_refdom.so); the C ABI behind the policy classes is the recording
// stand-in (tests/host_mock_abi.cc), so the test (tests/test_host_lbm.py) reads, rank by rank, which
// engine calls a HemeLB build would make and with which tables.
//
// What is shadowed, and why: Traits.h (the reference's drags in every streamer it has, one of which needs
// boost::ublas) and build_info.h (cmake-generated) under tests/host_shim_lbm/; the MPI / boost / logger
// stand-ins of oracle/ref_shim_dom/.  Test infrastructure only.
#include <mpi.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <memory>
#include <string>
#include <vector>

#include "ref_domain_build.h"
#include "debug/Debugger.h"
#include "geometry/Domain.h"
#include "geometry/FieldData.h"  // hemelb_b200/host
#include "geometry/neighbouring/NeighbouringDataManager.h"  // hemelb_b200/host
#include "lb/lb.hpp"
#include "lb/IncompressibilityChecker.hpp"  // hemelb_b200/host, over the reference's own net::PhasedBroadcastRegular
#include "lb/StabilityTester.h"             // hemelb_b200/host, likewise
#include "configuration/MonitoringConfig.h"
#include "lb/iolets/BoundaryValues.h"
#include "lb/iolets/InOutLetCosine.h"
#include "lb/iolets/InOutLetParabolicVelocity.h"
#include "lb/SimulationState.h"
#include "net/net.h"
#include "reporting/Dict.h"
#include "reporting/Timers.h"
#include "util/UnitConverter.h"

#ifdef HLB_REAL_ENGINE
// linked against hemelb_b200/libhemelb_b200.so: one rank, the engine's results come back for the oracle
static void hlb_mock_set_rank(int) {}
#else
extern "C" void hlb_mock_set_rank(int r);
#endif

namespace hemelb::reporting {
  // reporting/Dict.cc wraps ctemplate (absent here); nothing on this path writes a report
  Dict::Dict(const std::string&) : raw(nullptr, [](ctemplate::TemplateDictionary*) {}) {}
  Dict::Dict(ctemplate::TemplateDictionary*) : raw(nullptr, [](ctemplate::TemplateDictionary*) {}) {}
  Dict Dict::AddSectionDictionary(const std::string& s) { return Dict(s); }
  void Dict::SetValue(const std::string&, const std::string&) {}
  void Dict::SetIntValue(const std::string&, long) {}
  void Dict::SetBoolValue(const std::string&, bool) {}
  template <typename T> void Dict::SetFormattedValue(const std::string&, const char*, const T&) {}
  template void Dict::SetFormattedValue<double>(const std::string&, const char*, const double&);
}
namespace hemelb::debug {
  namespace {
    struct NoDebugger : Debugger {
      NoDebugger() : Debugger(nullptr, net::MpiCommunicator()) {}
      void BreakHere() override {}
      void Print(const char*, ...) override {}
      void Attach() override {}
    };
  }
  Debugger::Debugger(const char*, net::MpiCommunicator c) : mCommunicator(c) {}
  Debugger* Debugger::Get() {
    static NoDebugger none;
    return &none;
  }
}

namespace {
  using namespace hemelb;

  struct Job {
    refdom::GeometryArrays g;
    double dt = 0, dx = 0;
    int nIn = 0, nOut = 0;
    const double *inRec = nullptr, *outRec = nullptr;  // HLB_IOLET_RECORD_DOUBLES per iolet, lattice units
    int64_t steps = 0;
    int variant = 0;             // 0: D3Q19 BFL Nash; 1: D3Q15 SBB Nash; 2: D3Q27 BFL Nash; 3: D3Q19 BFL, Ladd velocity inlet
    int monitors = 0;            // 1: an lb::StabilityTester and an lb::IncompressibilityChecker are actions of every step
    double* monitorOut = nullptr;  // per rank {stability, smallest density, largest density, largest speed, available}
    int wall = 1;                // HLB_WALL_BFL, or HLB_WALL_GZS: GuoZhengShi walls + the NeighbouringDataManager
    const double* f0 = nullptr;  // rank 0's initial distributions (N * Q, the Domain's site order) or null: 0.05 everywhere
    double* fOut = nullptr;      // rank 0's distributions after the last step, or null
  };

  std::vector<util::clone_ptr<lb::InOutLet>> make_iolets(int n, const double* rec) {
    std::vector<util::clone_ptr<lb::InOutLet>> out;
    for (int i = 0; i < n; ++i) {
      const double* q = rec + (size_t)i * HLB_IOLET_RECORD_DOUBLES;
      if ((int)q[0] == 0) {
        auto c = util::make_clone_ptr<lb::InOutLetCosine>();
        c->SetDensityMean(q[9]);
        c->SetDensityAmp(q[10]);
        c->SetPhase(q[11]);
        c->SetPeriod(q[12]);
        c->SetWarmup((unsigned)q[13]);
        c->SetNormal(util::Vector3D<double>(q[1], q[2], q[3]));
        c->SetPosition(LatticePosition(q[4], q[5], q[6]));
        out.emplace_back(std::move(c));
      } else {
        auto v = util::make_clone_ptr<lb::InOutLetParabolicVelocity>();
        v->SetRadius(q[7]);
        v->SetMaxSpeed(q[8]);
        v->SetWarmup((unsigned)q[13]);
        v->SetNormal(util::Vector3D<double>(q[1], q[2], q[3]));
        v->SetPosition(LatticePosition(q[4], q[5], q[6]));
        out.emplace_back(std::move(v));
      }
    }
    return out;
  }

  // tests/host_shim_lbm/Traits.h: D3Q19 LBGK + the gpu:: streamers (BFL, Nash) by default
  using BflTraits = hemelb::Traits<>;
  using GzsTraits = hemelb::Traits<lb::D3Q19, lb::LBGK, lb::Normal, lb::gpu::Bulk,
                                   lb::gpu::Wall<lb::gpu::GuoZhengShi>::template type,
                                   lb::gpu::Inlet<lb::gpu::NashZerothOrderPressure>::template type,
                                   lb::gpu::Outlet<lb::gpu::NashZerothOrderPressure>::template type>;
  // other lattices and link rules through the same unmodified lb::LBM (Job::variant)
  using Q15SbbTraits = hemelb::Traits<lb::D3Q15, lb::LBGK, lb::Normal, lb::gpu::Bulk,
                                      lb::gpu::Wall<lb::gpu::SimpleBounceBack>::template type,
                                      lb::gpu::Inlet<lb::gpu::NashZerothOrderPressure>::template type,
                                      lb::gpu::Outlet<lb::gpu::NashZerothOrderPressure>::template type>;
  using Q27BflTraits = hemelb::Traits<lb::D3Q27, lb::LBGK, lb::Normal, lb::gpu::Bulk,
                                      lb::gpu::Wall<lb::gpu::BouzidiFirdaousLallemand>::template type,
                                      lb::gpu::Inlet<lb::gpu::NashZerothOrderPressure>::template type,
                                      lb::gpu::Outlet<lb::gpu::NashZerothOrderPressure>::template type>;
  using LaddTraits = hemelb::Traits<lb::D3Q19, lb::LBGK, lb::Normal, lb::gpu::Bulk,
                                    lb::gpu::Wall<lb::gpu::BouzidiFirdaousLallemand>::template type,
                                    lb::gpu::Inlet<lb::gpu::LaddIolet>::template type,
                                    lb::gpu::Outlet<lb::gpu::NashZerothOrderPressure>::template type>;

  template <class TraitsT> void run_rank(int rank, const Job& job) {
    hlb_mock_set_rank(rank);
    using Lattice = typename TraitsT::Lattice;
    auto const& info = Lattice::GetLatticeInfo();
    net::IOCommunicator comms{net::MpiCommunicator::World()};
    {
      geometry::GmyReadResult read = refdom::BuildReadResult(job.g, info, comms, nullptr);
      auto dom = std::make_shared<geometry::Domain>(info, read, comms);
      geometry::FieldData fd(dom);  // device-backed (hemelb_b200/host/geometry/FieldData.h)

      lb::SimulationState state{job.dt, 1000000000ul};
      lb::LbmParameters params(job.dt, job.dx);
      util::UnitConverter units(job.dt, job.dx, PhysicalPosition(0, 0, 0), DEFAULT_FLUID_DENSITY_Kg_per_m3, 0.0);
      auto inlets = make_iolets(job.nIn, job.inRec), outlets = make_iolets(job.nOut, job.outRec);
      lb::BoundaryValues inletValues(geometry::INLET_TYPE, *dom, inlets, &state, comms, units);
      lb::BoundaryValues outletValues(geometry::OUTLET_TYPE, *dom, outlets, &state, comms, units);
      reporting::Timers timers(comms);
      net::Net net(comms);

      // configuration/SimBuilder.h:153-160: the manager exists before the LBM, whose streamers register
      // their needs with it (GuoZhengShi); :235-236 shares them once everything is constructed
      std::unique_ptr<geometry::neighbouring::NeighbouringDataManager> ndm;
      if (job.wall == HLB_WALL_GZS)
        ndm = std::make_unique<geometry::neighbouring::NeighbouringDataManager>(fd, fd.GetNeighbouringData(), net);
      lb::LBM<TraitsT> lbm(params, &net, &fd, &state, timers, ndm.get());
      lbm.Initialise(&inletValues, &outletValues);
      if (ndm) {
        ndm->ShareNeeds();
        ndm->TransferNonFieldDependentInformation();
      }

      // configuration/SimBuilder.h:207-226: the monitors, actions of the same step manager
      using Checker = lb::IncompressibilityChecker<net::PhasedBroadcastRegular<>>;
      // (with the convergence check on: without it the root of the tree, which never looks at sites of its own,
      // hands down UndefinedStability for a stable run -- StabilityTester.h:192-239 -- and there is nothing to see)
      configuration::MonitoringConfig monitoring;
      monitoring.doConvergenceCheck = true;
      monitoring.convergenceVariable = extraction::source::Velocity{};
      monitoring.convergenceReferenceValue = 0.01;
      monitoring.convergenceRelativeTolerance = 1e-9;
      std::unique_ptr<lb::StabilityTester<Lattice>> tester;
      std::unique_ptr<Checker> checker;
      if (job.monitors) {
        tester = std::make_unique<lb::StabilityTester<Lattice>>(
            std::shared_ptr<const geometry::FieldData>(&fd, [](const geometry::FieldData*) {}), &net, &state, timers, monitoring);
        checker = std::make_unique<Checker>(dom.get(), &net, &state, lbm.GetPropertyCache(), timers, 0.05);
      }
      std::vector<net::IteratedAction*> actions;
      if (ndm) actions.push_back(ndm.get());
      actions.push_back(&inletValues);
      actions.push_back(&outletValues);
      actions.push_back(&lbm);
      if (tester) actions.push_back(tester.get());
      if (checker) actions.push_back(checker.get());

      // an initial condition written through the host view, as lb::InitialCondition does
      const site_t n = dom->GetLocalFluidSiteCount() * Lattice::NUMVECTORS;
      for (site_t i = 0; i < n; ++i) {
        const double v = (job.f0 && rank == 0) ? job.f0[i] : 0.05;
        *fd.GetFOld(i) = v;
        *fd.GetFNew(i) = v;
      }
      // net::phased::StepManager's order for one time step: every action's RequestComms, then PreSend,
      // PreReceive, PostReceive, EndIteration (Code/net/phased/StepManager.cc); SimulationMaster then swaps
      // the arrays and advances the state (SimulationMaster.impl.h:218-223)
      for (int64_t s = 0; s < job.steps; ++s) {
        for (auto* a : actions) a->RequestComms();
        net.Dispatch();  // (one phase here: the step manager sends, receives and waits between the calls below)
        for (auto* a : actions) a->PreSend();
        for (auto* a : actions) a->PreReceive();
        for (auto* a : actions) a->PostReceive();
        for (auto* a : actions) a->EndIteration();
        fd.SwapOldAndNew();
        state.Increment();
      }
      if (job.monitorOut) {
        double* o = job.monitorOut + 5 * rank;
        o[0] = (double)state.GetStability();
        o[4] = checker && checker->AreDensitiesAvailable() ? 1.0 : 0.0;
        if (o[4] != 0.0) {
          o[1] = checker->GetGlobalSmallestDensity();
          o[2] = checker->GetGlobalLargestDensity();
          o[3] = checker->GetGlobalLargestVelocityMagnitude();
        }
      }
      if (job.fOut && rank == 0) {
        const distribn_t* f = const_cast<geometry::FieldData const&>(fd).GetFOld(0);
        std::copy(f, f + n, job.fOut);
      }
    }  // (the Domain's windows are freed collectively here)
    hlb_mock_set_rank(-1);
  }

  void rank_body(int rank, void* arg) {
    const Job& job = *static_cast<const Job*>(arg);
    if (job.wall == HLB_WALL_GZS) run_rank<GzsTraits>(rank, job);
    else if (job.variant == 1) run_rank<Q15SbbTraits>(rank, job);
    else if (job.variant == 2) run_rank<Q27BflTraits>(rank, job);
    else if (job.variant == 3) run_rank<LaddTraits>(rank, job);
    else run_rank<BflTraits>(rank, job);
  }
}

extern "C" int hreal_run(int R, const int32_t* blockDims, int blockSize, int64_t N, const int32_t* coords, int64_t nb,
                         const int64_t* bsite, const uint8_t* btype, const int32_t* biolet, const float* bdist,
                         const uint8_t* bnavail, const float* bnormal, const int32_t* siteRank, double dt, double dx,
                         int nIn, const double* inRec, int nOut, const double* outRec, int64_t steps, const double* f0,
                         double* fOut, int wall, int monitors, double* monitorOut, int variant) {
  Job job;
  job.g.blockSize = blockSize;
  for (int k = 0; k < 3; ++k) job.g.bd[k] = blockDims[k];
  job.g.N = N;
  job.g.nb = nb;
  job.g.coords = coords;
  job.g.bsite = bsite;
  job.g.btype = btype;
  job.g.biolet = biolet;
  job.g.bdist = bdist;
  job.g.bnavail = bnavail;
  job.g.bnormal = bnormal;
  job.g.siteRank = siteRank;
  job.dt = dt;
  job.dx = dx;
  job.nIn = nIn;
  job.inRec = inRec;
  job.nOut = nOut;
  job.outRec = outRec;
  job.steps = steps;
  job.f0 = f0;
  job.fOut = fOut;
  job.wall = wall;
  job.monitors = monitors;
  job.monitorOut = monitorOut;
  job.variant = variant;
  fakempi_run(R, rank_body, &job);
  return 0;
}
