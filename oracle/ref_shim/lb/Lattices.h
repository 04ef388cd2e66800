// Shadow header (oracle/_ref build only): the reference's lb/Lattices.h selects a build-time
// default lattice from the generated build_info.h; the extraction sources that include it use
// nothing from it.
#pragma once
#include "lb/lattices/D3Q15.h"
#include "lb/lattices/D3Q19.h"
#include "lb/lattices/D3Q27.h"
