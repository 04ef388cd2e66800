// Test stand-in for net/net.h (Code/net/net.h defines net::Net from the build's mixins over MPI): the harness
// that makes lb::LBM's calls itself (tests/host_lbm_run.cc) only needs the InterfaceDelegationNet face.  The
// harness around the reference's own lb::LBM (tests/host_lbm_real.cc) uses the real header.  Test infrastructure.
#pragma once
#include "net/mixins/InterfaceDelegationNet.h"
namespace hemelb::net { class Net; }
