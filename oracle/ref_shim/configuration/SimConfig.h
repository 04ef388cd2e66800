// Shadow header (oracle/_ref build only): included by lb/iolets/InOutLet*.cc, unused by the
// GetDensity / GetVelocity bodies the oracle needs.
#pragma once
