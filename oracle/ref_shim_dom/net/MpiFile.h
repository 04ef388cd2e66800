// Shadow of net/MpiFile.h (oracle/_ref Domain build only): geometry::Domain does no file I/O.
#pragma once
namespace hemelb::net { class MpiFile {}; }
