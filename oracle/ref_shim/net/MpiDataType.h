// Shadow header (oracle/_ref build only): the reference's geometry/SiteData.h pulls this in for
// MPI datatype traits that the hot path never uses.  Intentionally empty.
#pragma once
