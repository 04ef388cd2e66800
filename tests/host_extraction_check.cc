// Compile-only check (tests/test_host_headers.py): the C++ face of the device extraction path
// builds against the reference's own extraction headers (PropertyOutputFile, OutputField, the
// GeometrySelector subclasses).
#include "extraction/GpuPropertyEncoder.h"

using namespace hemelb;

std::vector<char> example(hlb_gpu_t engine, const extraction::PropertyOutputFile& spec,
                          const std::vector<util::Vector3D<site_t>>& coords, std::vector<char>& buffer) {
  extraction::gpu::Units u{1e-4, 1e-4, PhysicalPosition(0, 0, 0), 1000.0, 80.0};
  extraction::gpu::PropertyEncoder enc(engine, spec, u, coords);
  buffer.resize(enc.CountWrittenSitesOnRank() * enc.CalcSiteWriteLen());
  enc.Encode(buffer);
  (void)enc.RequiredCaches();
  return enc.PrepareHeader(enc.CountWrittenSitesOnRank());
}
