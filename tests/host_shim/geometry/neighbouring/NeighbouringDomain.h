// Test-only stand-in (tests/test_host_headers.py) with the constructor shape of the reference's
// NeighbouringFieldData (Code/geometry/neighbouring/NeighbouringDomain.h:97-125).
#pragma once
#include <memory>
#include "units.h"
namespace hemelb::geometry::neighbouring {
  class NeighbouringDomain {};
  class NeighbouringFieldData {
  public:
    NeighbouringFieldData() = default;
    explicit NeighbouringFieldData(std::shared_ptr<NeighbouringDomain> d) : dom(d) {}
    std::shared_ptr<NeighbouringDomain> dom;
  };
}
