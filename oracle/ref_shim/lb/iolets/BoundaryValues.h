// Shadow header (oracle/_ref build only): the three members of lb::BoundaryValues the iolet
// link streamers call (Code/lb/iolets/BoundaryValues.h:50-70, .cc:162-165).
#pragma once
#include <vector>
#include "lb/iolets/InOutLet.h"
#include "lb/SimulationState.h"
namespace hemelb::lb {
  class BoundaryValues {
  public:
    std::vector<InOutLet*> iolets;
    std::vector<int> localIoletIDs;  // empty: every iolet is local (one rank)
    SimulationState* state = nullptr;
    LatticeDensity GetBoundaryDensity(int i) { return iolets[i]->GetDensity(state->Get0IndexedTimeStep()); }
    InOutLet* GetLocalIolet(unsigned i) { return localIoletIDs.empty() ? iolets[i] : iolets[localIoletIDs[i]]; }
    unsigned GetLocalIoletCount() const { return localIoletIDs.empty() ? iolets.size() : localIoletIDs.size(); }
    InOutLet* GetGlobalIolet(unsigned i) { return iolets[i]; }
    unsigned GetGlobalIoletCount() const { return iolets.size(); }
    LatticeTimeStep GetTimeStep() const { return state->GetTimeStep(); }
  };
}
