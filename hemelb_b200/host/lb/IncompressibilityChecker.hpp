// lb/IncompressibilityChecker.hpp -- the reference keeps the member definitions here and its users include
// this file (SimulationMaster.h:22, configuration/SimBuilder.h:16); the stand-in is header-only.
#ifndef HEMELB_LB_INCOMPRESSIBILITYCHECKER_HPP
#define HEMELB_LB_INCOMPRESSIBILITYCHECKER_HPP
#include "lb/IncompressibilityChecker.h"
#endif
