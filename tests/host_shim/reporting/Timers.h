// Test stand-in for reporting::Timers (Code/reporting/Timers.h): the indexing and Start/Stop surface
// the monitoring classes use.  Test infrastructure only.
#pragma once
namespace hemelb::reporting {
  struct Timer { void Start() {} void Stop() {} };
  class Timers {
  public:
    enum TimerName { total = 0, monitoring = 17, last = 32 };
    Timer& operator[](unsigned) { return one; }
  private:
    Timer one;
  };
}
