// Host-side C++ face of the device extraction path (include/hemelb_b200.h, hlb_xtr_*): turns the
// reference's own extraction::PropertyOutputFile (Code/extraction/PropertyOutputFile.h:24-31,
// OutputField.h:105-112, the GeometrySelector subclasses) into an hlb_xtr_spec and hands
// LocalPropertyOutput what its per-site loops produced: the local site count
// (CountWrittenSitesOnRank, LocalPropertyOutput.cc:133-145), the site length (CalcSiteWriteLen,
// :147-165), the header bytes (PrepareHeader, :178-213) and, per write, the record bytes that go
// into `buffer` behind the IO rank's time stamp (Write, :283-350).  Everything else of
// LocalPropertyOutput -- AllReduce / Scan of the lengths, the .off file, MpiFile::WriteAt -- stays
// as it is.  Header-only; errors become hemelb::Exception like everywhere in the reference.
#ifndef HEMELB_B200_HOST_EXTRACTION_GPUPROPERTYENCODER_H
#define HEMELB_B200_HOST_EXTRACTION_GPUPROPERTYENCODER_H

#include <cstdint>
#include <span>
#include <string>
#include <type_traits>
#include <variant>
#include <vector>

#include "Exception.h"
#include "extraction/GeometrySelectors.h"
#include "extraction/PropertyOutputFile.h"
#include "hemelb_b200.h"
#include "units.h"
#include "util/Vector3D.h"
#include "util/variant.h"

namespace hemelb::extraction::gpu
{
  // The C ABI's enums are the reference's orders (OutputField.h:33-44, io/formats/extraction.h:50-57);
  // pinned here so that a reordering upstream breaks the build, not the files.
  namespace detail {
    template <int I, class T> constexpr bool source_is = std::is_same_v<std::variant_alternative_t<I, source::Type>, T>;
    static_assert(source_is<HLB_XTR_PRESSURE, source::Pressure> && source_is<HLB_XTR_VELOCITY, source::Velocity> &&
                  source_is<HLB_XTR_SHEARSTRESS, source::ShearStress> && source_is<HLB_XTR_VONMISESSTRESS, source::VonMisesStress> &&
                  source_is<HLB_XTR_SHEARRATE, source::ShearRate> && source_is<HLB_XTR_STRESSTENSOR, source::StressTensor> &&
                  source_is<HLB_XTR_TRACTION, source::Traction> &&
                  source_is<HLB_XTR_TANGENTIALPROJECTIONTRACTION, source::TangentialProjectionTraction> &&
                  source_is<HLB_XTR_DISTRIBUTIONS, source::Distributions> && source_is<HLB_XTR_MPIRANK, source::MpiRank>);
    static_assert(std::variant_size_v<source::Type> == 10);
    using io::formats::extraction::TypeCode;
    static_assert(int(TypeCode::FLOAT) == HLB_XTR_FLOAT && int(TypeCode::DOUBLE) == HLB_XTR_DOUBLE &&
                  int(TypeCode::INT32) == HLB_XTR_INT32 && int(TypeCode::UINT32) == HLB_XTR_UINT32 &&
                  int(TypeCode::INT64) == HLB_XTR_INT64 && int(TypeCode::UINT64) == HLB_XTR_UINT64);
  }

  // the util::UnitConverter constructor arguments (Code/util/UnitConverter.cc:14-24) as SimBuilder has them
  struct Units
  {
    PhysicalTime timeStep;
    PhysicalDistance voxelSize;
    PhysicalPosition latticeOrigin;
    PhysicalDensity fluidDensity;
    PhysicalPressure referencePressure;
  };

  class PropertyEncoder
  {
  public:
    // globalSiteCoords: Domain::globalSiteCoords of this rank (Code/geometry/Domain.h:496-530)
    PropertyEncoder(hlb_gpu_t engine, const PropertyOutputFile& spec, const Units& units,
                    std::span<const util::Vector3D<site_t>> globalSiteCoords)
    {
      std::vector<hlb_xtr_field> fields;
      fields.reserve(spec.fields.size());
      for (auto const& f : spec.fields)
      {
        hlb_xtr_field x;
        x.name = f.name.c_str();
        x.source = static_cast<int>(f.src.index());      // extraction::source::Type order = HLB_XTR_* order
        x.typecode = static_cast<int>(code::type_to_enum(f.typecode));
        x.n_offsets = f.noffsets;
        x.offsets = f.offset.data();
        fields.push_back(x);
      }
      hlb_xtr_spec s{};
      SetSelector(*spec.geometry, s);
      s.n_fields = static_cast<int>(fields.size());
      s.fields = fields.data();
      s.time_step = units.timeStep;
      s.voxel_size = units.voxelSize;
      for (int k = 0; k < 3; ++k) s.origin[k] = units.latticeOrigin[k];
      s.fluid_density = units.fluidDensity;
      s.reference_pressure = units.referencePressure;
      static_assert(sizeof(util::Vector3D<site_t>) == 3 * sizeof(std::int64_t));
      Check(hlb_xtr_create(engine, &s, reinterpret_cast<const std::int64_t*>(globalSiteCoords.data()), &handle));
      Check(hlb_xtr_sizes(handle, &localSiteCount, &siteLength, &headerLength));
    }
    PropertyEncoder(const PropertyEncoder&) = delete;
    PropertyEncoder& operator=(const PropertyEncoder&) = delete;
    ~PropertyEncoder() { hlb_xtr_destroy(handle); }

    std::uint64_t CountWrittenSitesOnRank() const { return localSiteCount; }
    std::uint64_t CalcSiteWriteLen() const { return siteLength; }
    std::uint64_t HeaderLength() const { return headerLength; }

    // which MacroscopicPropertyCache members a step must refresh (PropertyActor.cc:22-75)
    std::uint32_t RequiredCaches() const
    {
      std::uint32_t mask = 0;
      Check(hlb_xtr_required_caches(handle, &mask));
      return mask;
    }

    std::vector<char> PrepareHeader(std::uint64_t globalSiteCount) const
    {
      std::vector<char> ans(headerLength);
      Check(hlb_xtr_header(handle, globalSiteCount, ans.data(), ans.size()));
      return ans;
    }

    // the site records of this rank into LocalPropertyOutput::buffer (after the time stamp)
    void Encode(std::span<char> records) const
    {
      Check(hlb_xtr_encode(handle, 0, localSiteCount, records.data(), records.size()));
    }

  private:
    static void Check(int rc)
    {
      if (rc != 0) throw Exception() << "hemelb_b200: " << hlb_gpu_last_error();
    }

    static void SetSelector(const GeometrySelector& g, hlb_xtr_spec& s)
    {
      auto put = [&](int at, const util::Vector3D<float>& v) {
        for (int k = 0; k < 3; ++k) s.selector_params[at + k] = v[k];
      };
      if (dynamic_cast<const WholeGeometrySelector*>(&g)) {
        s.selector = HLB_XTR_WHOLE;
      } else if (dynamic_cast<const GeometrySurfaceSelector*>(&g)) {
        s.selector = HLB_XTR_SURFACE;
      } else if (auto p = dynamic_cast<const PlaneGeometrySelector*>(&g)) {
        s.selector = HLB_XTR_PLANE_NORMALISED;  // GetNormal() is the constructor's normalised vector
        put(0, p->GetPoint());
        put(3, p->GetNormal());
        s.selector_params[6] = p->GetRadius();
      } else if (auto l = dynamic_cast<const StraightLineGeometrySelector*>(&g)) {
        s.selector = HLB_XTR_LINE;
        put(0, l->GetEndpoint1());
        put(3, l->GetEndpoint2());
      } else if (auto sp = dynamic_cast<const SurfacePointSelector*>(&g)) {
        s.selector = HLB_XTR_SURFACEPOINT;
        put(0, sp->GetPoint());
      } else {
        throw Exception() << "hemelb_b200: unknown GeometrySelector";
      }
    }

    hlb_xtr_t handle = nullptr;
    std::uint64_t localSiteCount = 0, siteLength = 0, headerLength = 0;
  };
}
#endif
