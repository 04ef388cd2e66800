// Shadow header (oracle/_ref build only): geometry/neighbouring/RequiredSiteInformation.cc includes
// the MPI wrapper without using it.  Intentionally empty.
#pragma once
