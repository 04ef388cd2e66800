// Test stand-in for net::PhasedBroadcastRegular (Code/net/PhasedBroadcastRegular.h,
// Code/net/PhasedBroadcast.h): a one-node tree.  The hooks a derived class overrides are the
// reference's; RunCycle() calls them in the order one full up-and-down cycle of the real class does on
// a rank that is both leaf and root.  Test infrastructure only.
#pragma once
#include "lb/SimulationState.h"
namespace hemelb::net {
  class Net;
  template <bool initialAction = false, unsigned splay = 1, unsigned overlap = 0, bool goDown = true, bool goUp = true>
  class PhasedBroadcastRegular {
  public:
    PhasedBroadcastRegular(Net*, const lb::SimulationState*, unsigned) {}
    virtual ~PhasedBroadcastRegular() = default;
    void RunCycle() {
      ProgressToParent(0);
      PostSendToParent(0);
      ProgressFromChildren(0);
      PostReceiveFromChildren(0);
      TopNodeAction();
      ProgressToChildren(0);
      ProgressFromParent(0);
      Effect();
    }
  protected:
    virtual void ProgressFromChildren(unsigned long) {}
    virtual void ProgressFromParent(unsigned long) {}
    virtual void ProgressToChildren(unsigned long) {}
    virtual void ProgressToParent(unsigned long) {}
    virtual void PostSendToParent(unsigned long) {}
    virtual void PostReceiveFromChildren(unsigned long) {}
    virtual void TopNodeAction() {}
    virtual void Effect() {}
    static const int NOPARENT = -1;  // (PhasedBroadcast.h:380)
    int GetParent() const { return NOPARENT; }
    template <class T> void ReceiveFromChildren(T*, int) {}
    template <class T> void ReceiveFromParent(T*, int) {}
    template <class T> void SendToChildren(T*, int) {}
    template <class T> void SendToParent(T*, int) {}
  };
}
