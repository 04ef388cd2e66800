// TEST INFRASTRUCTURE ONLY -- never linked into hemelb_b200/, never timed as the product.
//
// oracle/_ref/libhemelb_reflbm.so: the WHOLE hot path as the reference itself runs it.  Every class on the path is the
// reference's own, compiled unmodified from /root/reference/Code by oracle/Makefile:
//   lb::LBM<Traits>                 lb/lb.h, lb/lb.hpp                    (the phase schedule, lb.hpp:162-314)
//   BulkStreamer / StreamerTypeFactory<BFL|SBB|GZS, Nash|Ladd>           (the CPU streamers, lb/streamers/*.h)
//   geometry::FieldData             geometry/FieldData.{h,cc}            (f_old / f_new, SendAndReceive, CopyReceived)
//   geometry::Domain                geometry/Domain.cc                   (index tables)
//   NeighbouringDataManager         geometry/neighbouring/*.cc           (GuoZhengShi's site halo)
//   lb::BoundaryValues, InOutLet*   lb/iolets/*.cc
//   EquilibriumInitialCondition     lb/InitialCondition.{h,hpp,cc}
//   net::Net, net::phased::StepManager + NetConcern (two phases, separated concerns: SimBuilder.h:254-268)
// R ranks are threads of this process over oracle/fake_mpi.cc (as for libhemelb_refdom.so).  Actors are registered
// in the order and phases of configuration/SimBuilder.h:153-236; one time step is SimulationMaster::DoTimeStep's
// stepManager->CallActions(); fieldData->SwapOldAndNew(); simulationState->Increment() (SimulationMaster.impl.h:190-218).
// Shadowed: Traits.h (the reference's pulls in JunkYang.h -> boost::ublas) and the cmake-generated build_info.h,
// under oracle/ref_shim_lbm/; the MPI / boost / logger stand-ins of oracle/ref_shim_dom/.
//
// What it is for: tests/test_oracle_vs_ref_lbm.py holds the restatement (oracle/hemelb_oracle.cc, on the same
// emulated ranks) against it after several steps -- schedule, halo exchange, iolet densities over time and the
// initial condition included, not only the streamers one range at a time (that is libhemelb_ref.so's job).
#include <mpi.h>

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <memory>
#include <string>
#include <vector>

#include "ref_domain_build.h"
#include "debug/Debugger.h"
#include "geometry/Domain.h"
#include "geometry/FieldData.h"
#include "geometry/neighbouring/NeighbouringDataManager.h"
#include "lb/lb.hpp"
#include "lb/iolets/BoundaryValues.h"
#include "lb/iolets/InOutLetCosine.h"
#include "lb/iolets/InOutLetParabolicVelocity.h"
#include "lb/SimulationState.h"
#include "net/net.h"
#include "net/phased/NetConcern.h"
#include "net/phased/StepManager.h"
#include "reporting/Dict.h"
#include "reporting/Timers.h"
#include "util/UnitConverter.h"

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
    struct Silent : Debugger {
      Silent() : Debugger(nullptr, net::MpiCommunicator()) {}
      void BreakHere() override {}
      void Print(const char*, ...) override {}
      void Attach() override {}
    };
  }
  Debugger::Debugger(const char*, net::MpiCommunicator c) : mCommunicator(c) {}
  Debugger* Debugger::Get() {
    static Silent s;
    return &s;
  }
}

namespace {
  using namespace hemelb;

  // record layout of include/hemelb_b200.h (HLB_IOLET_RECORD_DOUBLES = 16), restated so this library needs nothing
  // of the product: {kind, normal xyz, position xyz, radius, maxSpeed, densityMean, densityAmp, phase, period, warmUp}
  constexpr int kRec = 16;

  struct Run {
    refdom::GeometryArrays g;
    double dt = 0, dx = 0;
    int nIn = 0, nOut = 0;
    const double *inRec = nullptr, *outRec = nullptr;
    int64_t steps = 0;
    int variant = 0;           // see rank_body
    int init = 0;              // 0: f0 given per rank; 1: EquilibriumInitialCondition(rho, m) through LBM::SetInitialConditions' callee
    double rho = 1.0, m[3] = {0, 0, 0};
    const double* f0 = nullptr;      // all ranks' initial f_old, rank r at fOff[r] (N_r * Q doubles, the Domain's site order)
    const int64_t* fOff = nullptr;   // R + 1 offsets (in doubles)
    double* fOut = nullptr;          // all ranks' f_old after the last swap, same offsets
    int64_t* nLocal = nullptr;       // per rank: the reference Domain's local fluid site count
    double* densities = nullptr;     // rank 0: GetBoundaryDensity of inlet 0.., outlet 0.. at each step (steps * (nIn+nOut)) or null
    double* loopSeconds = nullptr;   // rank 0's wall clock around the step loop (barrier on both sides), or null
  };

  std::vector<util::clone_ptr<lb::InOutLet>> iolets_from(int n, const double* rec) {
    std::vector<util::clone_ptr<lb::InOutLet>> out;
    for (int i = 0; i < n; ++i) {
      const double* q = rec + (size_t)i * kRec;
      const util::Vector3D<double> normal(q[1], q[2], q[3]);
      const LatticePosition where(q[4], q[5], q[6]);
      if ((int)q[0] == 0) {
        auto c = util::make_clone_ptr<lb::InOutLetCosine>();
        c->SetDensityMean(q[9]);
        c->SetDensityAmp(q[10]);
        c->SetPhase(q[11]);
        c->SetPeriod(q[12]);
        c->SetWarmup((unsigned)q[13]);
        c->SetNormal(normal);
        c->SetPosition(where);
        out.emplace_back(std::move(c));
      } else {
        auto v = util::make_clone_ptr<lb::InOutLetParabolicVelocity>();
        v->SetRadius(q[7]);
        v->SetMaxSpeed(q[8]);
        v->SetWarmup((unsigned)q[13]);
        v->SetNormal(normal);
        v->SetPosition(where);
        out.emplace_back(std::move(v));
      }
    }
    return out;
  }

  // MRT<MomentBasis> under the one-parameter name Traits wants (the basis fixes the lattice)
  template <lb::lattice_type> using Mrt15 = lb::MRT<lb::DHumieresD3Q15MRTBasis>;
  template <lb::lattice_type> using Mrt19 = lb::MRT<lb::DHumieresD3Q19MRTBasis>;

  template <class L, template <class> class W, template <class> class I, template <class> class O = lb::cpu::NashIolet,
            template <lb::lattice_type> class K = lb::LBGK>
  using T_ = hemelb::Traits<L, K, lb::Normal, lb::BulkStreamer, W, I, O>;

  template <class TraitsT> void one_rank(int rank, const Run& run) {
    using Lattice = typename TraitsT::Lattice;
    auto const& info = Lattice::GetLatticeInfo();
    net::IOCommunicator comms{net::MpiCommunicator::World()};
    {
      geometry::GmyReadResult read = refdom::BuildReadResult(run.g, info, comms, nullptr);
      auto dom = std::make_shared<geometry::Domain>(info, read, comms);
      auto fd = std::make_shared<geometry::FieldData>(dom);

      lb::SimulationState state{run.dt, 1000000000ul};
      lb::LbmParameters params(run.dt, run.dx);
      util::UnitConverter units(run.dt, run.dx, PhysicalPosition(0, 0, 0), DEFAULT_FLUID_DENSITY_Kg_per_m3, 0.0);
      reporting::Timers timers(comms);
      net::Net net(comms);

      // configuration/SimBuilder.h:153-236, in its order; (actor, phase)
      std::vector<std::pair<net::IteratedAction*, unsigned>> actors;
      geometry::neighbouring::NeighbouringDataManager ndm(*fd, fd->GetNeighbouringData(), net);
      actors.emplace_back(&ndm, 0u);
      lb::LBM<TraitsT> lbm(params, &net, fd.get(), &state, timers, &ndm);
      actors.emplace_back(&lbm, 1u);
      lb::BoundaryValues inletValues(geometry::INLET_TYPE, *dom, iolets_from(run.nIn, run.inRec), &state, comms, units);
      actors.emplace_back(&inletValues, 1u);
      lb::BoundaryValues outletValues(geometry::OUTLET_TYPE, *dom, iolets_from(run.nOut, run.outRec), &state, comms, units);
      actors.emplace_back(&outletValues, 1u);

      lbm.Initialise(&inletValues, &outletValues);
      const site_t nf = dom->GetLocalFluidSiteCount() * (site_t)Lattice::NUMVECTORS;
      if (run.nLocal) run.nLocal[rank] = dom->GetLocalFluidSiteCount();
      if (run.init == 1) {
        // what lb::InitialCondition's visitor reaches for an <equilibrium> initial condition (InitialCondition.hpp)
        lb::EquilibriumInitialCondition ic(std::nullopt, run.rho, run.m[0], run.m[1], run.m[2]);
        ic.SetFs<Lattice>(fd.get(), comms);
        ic.SetTime(&state);
      } else {
        const double* src = run.f0 + run.fOff[rank];
        if (run.fOff[rank + 1] - run.fOff[rank] != nf) {
          std::fprintf(stderr, "ref_lbm_driver: rank %d holds %ld doubles, caller laid out %ld\n", rank, (long)nf,
                       (long)(run.fOff[rank + 1] - run.fOff[rank]));
          MPI_Abort(MPI_COMM_WORLD, 2);
        }
        for (site_t i = 0; i < nf; ++i) {
          *fd->GetFOld(i) = src[i];
          *fd->GetFNew(i) = src[i];
        }
      }
      ndm.ShareNeeds();
      ndm.TransferNonFieldDependentInformation();

      net::phased::NetConcern netConcern(net);
      net::phased::StepManager stepManager(2, &timers, net::separate_communications);
      for (auto [a, phase] : actors) stepManager.RegisterIteratedActorSteps(*a, phase);
      stepManager.RegisterCommsForAllPhases(netConcern);

      MPI_Barrier(MPI_COMM_WORLD);
      const auto t0 = std::chrono::steady_clock::now();
      for (int64_t s = 0; s < run.steps; ++s) {
        if (run.densities && rank == 0) {
          double* d = run.densities + s * (run.nIn + run.nOut);
          for (int i = 0; i < run.nIn; ++i) d[i] = inletValues.GetBoundaryDensity(i);
          for (int i = 0; i < run.nOut; ++i) d[run.nIn + i] = outletValues.GetBoundaryDensity(i);
        }
        stepManager.CallActions();
        fd->SwapOldAndNew();
        state.Increment();
      }
      MPI_Barrier(MPI_COMM_WORLD);
      if (run.loopSeconds && rank == 0)
        *run.loopSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      if (run.fOut) {
        const distribn_t* f = const_cast<geometry::FieldData const&>(*fd).GetFOld(0);
        std::copy(f, f + nf, run.fOut + run.fOff[rank]);
      }
    }  // (the Domain's one-sided windows are freed collectively here)
  }

  void rank_body(int rank, void* arg) {
    const Run& run = *static_cast<const Run*>(arg);
    using namespace lb;
    switch (run.variant) {
      case 0: one_rank<T_<D3Q19, cpu::BflWall, cpu::NashIolet>>(rank, run); break;   // the headline bundle
      case 1: one_rank<T_<D3Q15, cpu::SbbWall, cpu::NashIolet>>(rank, run); break;
      case 2: one_rank<T_<D3Q27, cpu::BflWall, cpu::NashIolet>>(rank, run); break;
      case 3: one_rank<T_<D3Q19, cpu::BflWall, cpu::LaddIolet>>(rank, run); break;   // velocity inlet, pressure outlets
      case 4: one_rank<T_<D3Q19, cpu::GzsWall, cpu::NashIolet>>(rank, run); break;   // site halo through the NDM
      case 5: one_rank<T_<D3Q19, cpu::SbbWall, cpu::NashIolet>>(rank, run); break;
      case 6: one_rank<T_<D3Q19, cpu::GzsWall, cpu::LaddIolet>>(rank, run); break;   // configs[3]'s link rules with LBGK
      case 7: one_rank<T_<D3Q15, cpu::BflWall, cpu::NashIolet>>(rank, run); break;
      case 8: one_rank<T_<D3Q27, cpu::SbbWall, cpu::NashIolet>>(rank, run); break;
      // MRT: with velocity iolets on both sides (MRT + Nash does not compile in the reference, MRT.h:73-86)
      case 9: one_rank<T_<D3Q19, cpu::BflWall, cpu::LaddIolet, cpu::LaddIolet, Mrt19>>(rank, run); break;
      case 10: one_rank<T_<D3Q15, cpu::SbbWall, cpu::LaddIolet, cpu::LaddIolet, Mrt15>>(rank, run); break;
      case 11: one_rank<T_<D3Q19, cpu::BflWall, cpu::LaddIolet, cpu::LaddIolet>>(rank, run); break;  // (LBGK beside case 9)
      default: std::fprintf(stderr, "ref_lbm_driver: unknown variant %d\n", run.variant); MPI_Abort(MPI_COMM_WORLD, 3);
    }
  }
}

extern "C" int hreflbm_run(int R, const int32_t* blockDims, int blockSize, int64_t N, const int32_t* coords, int64_t nb,
                           const int64_t* bsite, const uint8_t* btype, const int32_t* biolet, const float* bdist,
                           const uint8_t* bnavail, const float* bnormal, const int32_t* siteRank, double dt, double dx,
                           int nIn, const double* inRec, int nOut, const double* outRec, int64_t steps, int variant,
                           int init, double rho, const double* momentum, const double* f0, const int64_t* fOff,
                           double* fOut, int64_t* nLocal, double* densities, double* loopSeconds) {
  Run run;
  run.g.blockSize = blockSize;
  for (int k = 0; k < 3; ++k) run.g.bd[k] = blockDims[k];
  run.g.N = N;
  run.g.nb = nb;
  run.g.coords = coords;
  run.g.bsite = bsite;
  run.g.btype = btype;
  run.g.biolet = biolet;
  run.g.bdist = bdist;
  run.g.bnavail = bnavail;
  run.g.bnormal = bnormal;
  run.g.siteRank = siteRank;
  run.dt = dt;
  run.dx = dx;
  run.nIn = nIn;
  run.inRec = inRec;
  run.nOut = nOut;
  run.outRec = outRec;
  run.steps = steps;
  run.variant = variant;
  run.init = init;
  run.rho = rho;
  if (momentum) std::copy(momentum, momentum + 3, run.m);
  run.f0 = f0;
  run.fOff = fOff;
  run.fOut = fOut;
  run.nLocal = nLocal;
  run.densities = densities;
  run.loopSeconds = loopSeconds;
  if (init == 0 && (!f0 || !fOff)) return 1;
  if (fOut && !fOff) return 1;
  fakempi_run(R, rank_body, &run);
  return 0;
}
