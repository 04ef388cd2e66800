// Internal seam between the device-side Domain builder (domain_builder.cu) and the engine
// (abi.cu): raw views of a not-yet-finalised engine handle so that the builder can write the
// tables device-to-device instead of round-tripping N*Q int64 through the host.  Not part of the
// C ABI (include/hemelb_b200.h); both translation units live in the same shared library.
#pragma once
#include <cstdint>

#include "../../include/hemelb_b200.h"

struct hlb_gpu_raw {
  // device
  uint32_t* nbr;        // (Q-1) planes of `stride`: pre-renumbering internal targets
  int32_t* coordsAll;   // 3 planes of `stride` (allocated on demand when cfg.reorder), else null
  int64_t stride;
  // host staging of the boundary tables (plane-major over bStride), uploaded by hlb_gpu_finalise
  uint32_t* hWall;
  uint32_t* hIolet;
  int32_t* hIoletId;
  float* hCut;          // (Q-1) planes
  double* hNormal;      // 3 planes
  int32_t* hCoords;     // 3 planes
  int64_t bStride, NB;
};

// fills `out`; allocates coordsAll when the handle was created with reorder = 1
int hlb_gpu_internal_raw(hlb_gpu_t h, hlb_gpu_raw* out);
// the builder wrote neighbour table, boundary tables and (if reorder) every site's coordinates
int hlb_gpu_internal_mark_installed(hlb_gpu_t h);
// error text shared with hlb_gpu_last_error()
int hlb_internal_fail(const char* msg);
