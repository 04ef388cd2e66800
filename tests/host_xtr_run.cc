// Runnable check of extraction/GpuPropertyEncoder.h: the reference's own PropertyOutputFile /
// OutputField / GeometrySelector objects go in, and what reaches the C ABI (hlb_xtr_spec) is
// recorded by tests/host_mock_abi.cc for tests/test_host_lbm.py to compare.  One spec per selector
// class, every extraction::source among the fields.  Test infrastructure (needs /root/reference).
#include <cstdio>
#include <vector>

#include "extraction/GpuPropertyEncoder.h"

using namespace hemelb;
namespace ex = hemelb::extraction;

hemelb::util::Vector3DBase::HandlerFunction* hemelb::util::Vector3DBase::handler = nullptr;

static ex::OutputField field(const char* name, ex::source::Type src, ex::code::Type tc, std::vector<double> off = {}) {
  return ex::OutputField{name, src, tc, (std::uint32_t)off.size(), off};
}

int main() {
  using V = util::Vector3D<float>;
  std::vector<util::Vector3D<site_t>> coords = {{3, 4, 5}, {6, 7, 8}, {9, 10, 11}};
  ex::gpu::Units units{1e-4, 2e-4, PhysicalPosition(0.1, 0.2, 0.3), 1000.0, 80.0};
  std::vector<ex::PropertyOutputFile> specs;
  {
    ex::PropertyOutputFile f;
    f.filename = "whole.xtr";
    f.frequency = 10;
    f.geometry.reset(new ex::WholeGeometrySelector());
    f.fields = {field("pressure", ex::source::Pressure{}, float{}, {80.0}), field("velocity", ex::source::Velocity{}, float{}),
                field("distributions", ex::source::Distributions{}, double{})};
    specs.push_back(std::move(f));
  }
  {
    ex::PropertyOutputFile f;
    f.geometry.reset(new ex::GeometrySurfaceSelector());
    f.fields = {field("shearstress", ex::source::ShearStress{}, float{}), field("traction", ex::source::Traction{}, double{})};
    specs.push_back(std::move(f));
  }
  {
    ex::PropertyOutputFile f;
    f.geometry.reset(new ex::PlaneGeometrySelector(V(0.001f, 0.002f, 0.003f), V(0.f, 0.f, 2.f), 0.004f));
    f.fields = {field("velocity", ex::source::Velocity{}, double{}), field("stresstensor", ex::source::StressTensor{}, float{}),
                field("rank", ex::source::MpiRank{}, std::int32_t{})};
    specs.push_back(std::move(f));
  }
  {
    ex::PropertyOutputFile f;
    f.geometry.reset(new ex::PlaneGeometrySelector(V(0.5f, 0.f, 0.f), V(3.f, 0.f, 4.f)));
    f.fields = {field("pressure", ex::source::Pressure{}, double{})};
    specs.push_back(std::move(f));
  }
  {
    ex::PropertyOutputFile f;
    f.geometry.reset(new ex::StraightLineGeometrySelector(V(0.f, 0.f, 0.f), V(0.f, 0.f, 0.01f)));
    f.fields = {field("vonmisesstress", ex::source::VonMisesStress{}, float{}), field("shearrate", ex::source::ShearRate{}, float{})};
    specs.push_back(std::move(f));
  }
  {
    ex::PropertyOutputFile f;
    f.geometry.reset(new ex::SurfacePointSelector(V(0.01f, 0.02f, 0.03f)));
    f.fields = {field("tangentialprojectiontraction", ex::source::TangentialProjectionTraction{}, float{})};
    specs.push_back(std::move(f));
  }
  try {
    for (auto const& spec : specs) {
      ex::gpu::PropertyEncoder enc(nullptr, spec, units, coords);
      std::vector<char> header = enc.PrepareHeader(1234);
      std::vector<char> records(enc.CountWrittenSitesOnRank() * enc.CalcSiteWriteLen());
      enc.Encode(records);
      printf("sites=%llu site_len=%llu header_len=%llu header0=%c records0=%c caches=%u\n",
             (unsigned long long)enc.CountWrittenSitesOnRank(), (unsigned long long)enc.CalcSiteWriteLen(),
             (unsigned long long)enc.HeaderLength(), header[0], records[0], enc.RequiredCaches());
    }
  } catch (std::exception& e) {
    fprintf(stderr, "host_xtr_run: %s\n", e.what());
    return 1;
  }
  return 0;
}
