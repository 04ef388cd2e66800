// lb/StabilityTester.h -- lb::StabilityTester for a build whose distributions live on the B200.
//
// Stands in for Code/lb/StabilityTester.h (same class template, constructor and PhasedBroadcast role,
// so configuration/SimBuilder.h:207-214 and SimulationMaster.h:115 compile unchanged) when
// hemelb_b200/host precedes Code/ on the include path.  The reference's tester walks every local site
// on the host each cycle: *GetFNew(i * Q + l) for the "value > 0" test and, with the convergence
// check on, GetFNew<L>(i) against GetSite(i).GetFOld<L>() (StabilityTester.h:97-141, 156-180).
// Through the device-backed geometry::FieldData that walk would pull both distribution arrays to the
// host every time step.  Here the local verdict comes from one device reduction
// (hlb_gpu_stability: 16 bytes leave the GPU); how the verdicts of the ranks meet -- up and down
// the PhasedBroadcast tree, ints of lb::Stability -- is the reference's own protocol and untouched.
#ifndef HEMELB_LB_STABILITYTESTER_H
#define HEMELB_LB_STABILITYTESTER_H

#include <algorithm>
#include <array>
#include <memory>
#include <variant>

#include "net/PhasedBroadcastRegular.h"
#include "geometry/Domain.h"
#include "geometry/FieldData.h"
#include "configuration/MonitoringConfig.h"
#include "reporting/Timers.h"

namespace hemelb::lb
{
  namespace gpu
  {
    // What a node reports upwards once it knows its own sites' verdict and its children's
    // (StabilityTester.h:192-239): Unstable wins; with the convergence check on, a node that is the
    // tree's root or itself converged takes StableAndConverged from a converged child, and any child
    // that is stable-but-unconverged brings the node back to Stable.
    template <class Children>
    inline int MergeStability(int own, Children const& children, bool convergenceCheck, bool isRoot)
    {
      if (own == Unstable) return Unstable;
      auto any = [&](int what) { return std::find(children.begin(), children.end(), what) != children.end(); };
      if (any(Unstable)) return Unstable;
      if (!convergenceCheck) return own;
      if (any(StableAndConverged) && (own == StableAndConverged || isRoot)) own = StableAndConverged;
      if (any(Stable)) own = Stable;
      return own;
    }
  }

  template<class LatticeType>
  class StabilityTester : public net::PhasedBroadcastRegular<>
  {
    public:
      StabilityTester(std::shared_ptr<const geometry::FieldData> latDat, net::Net* net, SimulationState* simState,
                      reporting::Timers& timers, const hemelb::configuration::MonitoringConfig& config) :
          net::PhasedBroadcastRegular<>(net, simState, kSpread), fieldData(std::move(latDat)), state(simState),
          timers(timers), config(config)
      {
        Reset();
      }

      bool ShouldTerminateWhenConverged() const { return config.convergenceTerminate; }

      void Reset()
      {
        up = down = UndefinedStability;
        state->SetStability(UndefinedStability);
        children.fill(UndefinedStability);
      }

    protected:
      void ProgressFromChildren(unsigned long) override { ReceiveFromChildren<int>(children.data(), 1); }
      void ProgressFromParent(unsigned long) override { ReceiveFromParent<int>(&down, 1); }
      void ProgressToChildren(unsigned long) override { SendToChildren<int>(&down, 1); }
      void ProgressToParent(unsigned long) override { SendToParent<int>(&up, 1); }

      // The local sites' verdict, taken here (not in ProgressToParent) so that the step has finished
      // streaming -- as in the reference -- and before SimulationMaster swaps the arrays.
      void PostSendToParent(unsigned long) override
      {
        timers[hemelb::reporting::Timers::monitoring].Start();
        if (up != Unstable)
        {
          if (config.doConvergenceCheck
              && !std::holds_alternative<extraction::source::Velocity>(config.convergenceVariable))
            throw Exception() << "Convergence check based on requested variable currently not available";
          double verdict[2] = { 0.0, 0.0 };
          // (a rank without an engine yet has not streamed anything: nothing to test)
          if (hlb_gpu_t engine = const_cast<geometry::FieldData&>(*fieldData).EngineIfBuilt())
            geometry::FieldData::Check(hlb_gpu_stability(engine, config.doConvergenceCheck ? 1 : 0, verdict));
          if (verdict[0] != 0.0)
            up = Unstable;
          else
          {
            // some site whose |u_new - u_old| / reference exceeds the tolerance <=> the largest one does
            const bool unconverged = config.doConvergenceCheck
                && verdict[1] / config.convergenceReferenceValue > config.convergenceRelativeTolerance;
            up = (config.doConvergenceCheck && !unconverged) ? StableAndConverged : Stable;
          }
        }
        timers[hemelb::reporting::Timers::monitoring].Stop();
      }

      void TopNodeAction() override { down = up; }

      void PostReceiveFromChildren(unsigned long) override
      {
        timers[hemelb::reporting::Timers::monitoring].Start();
        up = gpu::MergeStability(up, children, config.doConvergenceCheck, GetParent() == NOPARENT);
        timers[hemelb::reporting::Timers::monitoring].Stop();
      }

      void Effect() override { state->SetStability((Stability) down); }

    private:
      static constexpr unsigned kSpread = 10;  // the reference's tree width (StabilityTester.h:252)
      std::shared_ptr<const geometry::FieldData> fieldData;
      int up, down;
      std::array<int, kSpread> children;
      lb::SimulationState* state;
      reporting::Timers& timers;
      const hemelb::configuration::MonitoringConfig& config;
  };
}
#endif
