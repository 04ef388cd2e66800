// Shadow header (oracle/_ref build only): included by lb/iolets/InOutLetCosine.cc, unused there.
#pragma once
