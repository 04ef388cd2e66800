// ---------------------------------------------------------------------------------------------
// ref_xtr.h -- TEST INFRASTRUCTURE ONLY (part of oracle/_ref/libhemelb_ref.so; included by
// ref_driver.cc).
//
// Drives the UNMODIFIED reference extraction / checkpoint sources
//   extraction/{LocalPropertyOutput,LbDataSourceIterator,LocalDistributionInput,GeometrySelector,
//               WholeGeometrySelector,GeometrySurfaceSelector,PlaneGeometrySelector,
//               StraightLineGeometrySelector,SurfacePointSelector,IterableDataSource}.cc,
//   io/writers/Xdr*.cc, io/readers/Xdr*.cc, util/UnitConverter.cc
// over the emulated ranks of a RefSim: one std::thread per rank, collectives and file I/O through
// the shadow net/IOCommunicator.h and net/MpiFile.h.  What the reference writes to disk here is
// what the GPU extraction path (hlb_xtr_*) has to reproduce byte for byte.
// ---------------------------------------------------------------------------------------------
#pragma once
#include <optional>
#include <string>
#include <thread>

#include "extraction/GeometrySelectors.h"
#include "extraction/LbDataSourceIterator.h"
#include "extraction/LocalDistributionInput.h"
#include "extraction/LocalPropertyOutput.h"
#include "util/UnitConverter.h"

namespace {

struct XtrSession {
  SimBase* S = nullptr;
  std::shared_ptr<net::EmulatedWorld> world;
  std::shared_ptr<util::UnitConverter> conv;
  std::vector<std::unique_ptr<net::IOCommunicator>> comms;
  std::vector<std::unique_ptr<extraction::LbDataSourceIterator>> src;
  std::vector<std::unique_ptr<extraction::LocalPropertyOutput>> out;
  std::string error;
};

// run fn(rank) on one thread per emulated rank; a throwing rank leaves the barrier so that the
// others cannot hang, and its message is kept
template <class F>
std::string RunRanks(int R, net::EmulatedWorld& world, F fn) {
  std::vector<std::string> errs(R);
  std::vector<std::thread> th;
  for (int r = 0; r < R; ++r)
    th.emplace_back([&, r] {
      try {
        fn(r);
      } catch (std::exception& e) {
        errs[r] = e.what();
        if (errs[r].empty()) errs[r] = "exception";
        world.sync.arrive_and_drop();
      }
    });
  for (auto& t : th) t.join();
  for (auto& e : errs)
    if (!e.empty()) return e;
  return {};
}

extraction::source::Type SourceOf(int k) {
  using namespace extraction::source;
  switch (k) {
    case 0: return Pressure{};
    case 1: return Velocity{};
    case 2: return ShearStress{};
    case 3: return VonMisesStress{};
    case 4: return ShearRate{};
    case 5: return StressTensor{};
    case 6: return Traction{};
    case 7: return TangentialProjectionTraction{};
    case 8: return Distributions{};
    default: return MpiRank{};
  }
}

thread_local std::string g_xtrError;

}  // namespace

extern "C" {

const char* href_xtr_last_error() { return g_xtrError.c_str(); }

// selKind: 0 whole, 1 surface, 2 plane (point[3], normal[3], radius), 3 line (p1[3], p2[3]),
// 4 surface point (p[3]).  typecode as io::formats::extraction::TypeCode.
void* href_xtr_open(void* sp, const char* path, uint64_t frequency, int singleTimestepFiles,
                    int selKind, const float* sel, int nFields, const char* const* names,
                    const int* srcKind, const int* typecode, const int* noffsets,
                    const double* offsets, double dt, double dx, const double* origin,
                    double fluidDensity, double refPressure) {
  SimBase* S = (SimBase*)sp;
  auto X = std::make_unique<XtrSession>();
  X->S = S;
  X->world = std::make_shared<net::EmulatedWorld>(S->R);
  X->conv = std::make_shared<util::UnitConverter>(dt, dx, PhysicalPosition(origin[0], origin[1], origin[2]),
                                                  fluidDensity, refPressure);
  extraction::PropertyOutputFile spec;
  spec.filename = path;
  spec.frequency = frequency;
  using V = util::Vector3D<float>;
  switch (selKind) {
    case 0: spec.geometry = util::make_clone_ptr<extraction::WholeGeometrySelector>(); break;
    case 1: spec.geometry = util::make_clone_ptr<extraction::GeometrySurfaceSelector>(); break;
    case 2:
      if (sel[6] > 0.f)
        spec.geometry = util::make_clone_ptr<extraction::PlaneGeometrySelector>(V(sel[0], sel[1], sel[2]), V(sel[3], sel[4], sel[5]), sel[6]);
      else
        spec.geometry = util::make_clone_ptr<extraction::PlaneGeometrySelector>(V(sel[0], sel[1], sel[2]), V(sel[3], sel[4], sel[5]));
      break;
    case 3: spec.geometry = util::make_clone_ptr<extraction::StraightLineGeometrySelector>(V(sel[0], sel[1], sel[2]), V(sel[3], sel[4], sel[5])); break;
    default: spec.geometry = util::make_clone_ptr<extraction::SurfacePointSelector>(V(sel[0], sel[1], sel[2])); break;
  }
  const double* off = offsets;
  for (int i = 0; i < nFields; ++i) {
    extraction::OutputField f;
    f.name = names[i];
    f.src = SourceOf(srcKind[i]);
    f.typecode = extraction::code::enum_to_type((io::formats::extraction::TypeCode)typecode[i]);
    f.noffsets = noffsets[i];
    f.offset.assign(off, off + noffsets[i]);
    off += noffsets[i];
    spec.fields.push_back(f);
  }
  if (singleTimestepFiles) spec.ts_mode = extraction::single_timestep_files{};
  else spec.ts_mode = extraction::multi_timestep_file{};
  X->comms.resize(S->R);
  X->src.resize(S->R);
  X->out.resize(S->R);
  for (int r = 0; r < S->R; ++r) {
    RankState& rs = *S->ranks[r];
    X->comms[r] = std::make_unique<net::IOCommunicator>(X->world, r);
    X->src[r] = std::make_unique<extraction::LbDataSourceIterator>(*rs.cache, rs.fd, r, X->conv);
  }
  XtrSession* x = X.get();
  g_xtrError = RunRanks(S->R, *X->world, [&](int r) {
    x->out[r] = std::make_unique<extraction::LocalPropertyOutput>(*x->src[r], spec, *x->comms[r]);
  });
  if (!g_xtrError.empty()) return nullptr;
  return X.release();
}

int href_xtr_write(void* xp, uint64_t timestep, uint64_t totalSteps) {
  XtrSession* X = (XtrSession*)xp;
  g_xtrError = RunRanks(X->S->R, *X->world, [&](int r) { X->out[r]->Write(timestep, totalSteps); });
  return g_xtrError.empty() ? 0 : 1;
}

void href_xtr_close(void* xp) { delete (XtrSession*)xp; }

// extraction::LocalDistributionInput::LoadDistribution (CheckpointInitialCondition::SetFs,
// lb/InitialCondition.hpp).  target < 0: "use the last time step in the file".
int href_load_checkpoint(void* sp, const char* xtrPath, const char* offPath, int64_t target, uint64_t* timeOut) {
  SimBase* S = (SimBase*)sp;
  auto world = std::make_shared<net::EmulatedWorld>(S->R);
  std::vector<uint64_t> t(S->R, 0);
  g_xtrError = RunRanks(S->R, *world, [&](int r) {
    net::IOCommunicator comm(world, r);
    std::optional<std::filesystem::path> off;
    if (offPath && offPath[0]) off = std::filesystem::path(offPath);
    extraction::LocalDistributionInput in(xtrPath, off, comm);
    std::optional<LatticeTimeStep> tt;
    if (target >= 0) tt = (LatticeTimeStep)target;
    in.LoadDistribution(&S->ranks[r]->fd, tt);
    t[r] = *tt;
  });
  if (timeOut) *timeOut = t[0];
  return g_xtrError.empty() ? 0 : 1;
}

// io::formats::extraction helpers pinned by LocalPropertyOutputTests.cc:127-150
uint64_t href_xtr_string_length(const char* s) { return io::formats::extraction::GetStoredLengthOfString(s); }
uint64_t href_xtr_field_header_length(const char* name, uint32_t noff, int typecode) {
  return io::formats::extraction::GetFieldHeaderLength(name, noff, (io::formats::extraction::TypeCode)typecode);
}

}  // extern "C"
