// Internal seam between the device-side Domain builder (domain_builder.cu) and the engine
// (abi.cu): raw views of a not-yet-finalised engine handle so that the builder can write the
// tables device-to-device instead of round-tripping N*Q int64 through the host.  Not part of the
// C ABI (include/hemelb_b200.h); both translation units live in the same shared library.
#pragma once
#include <cstdint>
#include <vector_types.h>

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

// Read/write view of a FINALISED engine handle for the extraction / checkpoint kernels
// (extraction.cu): where the distributions, caches and boundary tables live on the device.
struct hlb_gpu_view {
  int Q, device, rank, nranks;
  int64_t N, stride, bStride;
  const uint2* bInfo;         // per 32 internal sites {bitmap of boundary-typed sites, ordinal of the first}
  double* f[2];               // f[0] = current f_old, f[1] = current f_new (SoA, `stride`)
  const uint32_t* perm;       // reference site -> internal site, or null (identity)
  const uint32_t* wallMask;   // by boundary ordinal of the INTERNAL site
  const double* wallNormal;   // 3 planes of bStride
  double* cache[8];           // MacroscopicPropertyCache arrays (reference site-major), null if never requested
  void* computeStream;        // cudaStream_t
};
int hlb_gpu_internal_view(hlb_gpu_t h, hlb_gpu_view* out);
int hlb_gpu_internal_count_launch(hlb_gpu_t h, int64_t n);

// device-resident site coordinates of a built Domain: 3 planes of N int32, reference site order
struct hlb_dom_handle;
int hlb_dom_internal_coords(hlb_dom_handle* d, const int32_t** planes, int64_t* n_sites, int* device);
