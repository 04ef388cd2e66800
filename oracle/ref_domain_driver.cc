// TEST INFRASTRUCTURE ONLY.  Drives the reference's UNMODIFIED geometry::Domain (Code/geometry/Domain.cc),
// octree::LookupTree / DistributedStore (Code/geometry/LookupTree.cc) and decomposition::BasicDecomposition
// (Code/geometry/decomposition/BasicDecomposition.cc) over R emulated ranks (threads; oracle/fake_mpi.cc
// stands in for MPI, the reference's own net:: classes run on top of it) so that the index tables of
// the hot path -- site order, neighbourIndices, neighbouring processors, shared-distribution slots,
// streaming indices of received distributions -- can be compared with the oracle's restatement and
// with hemelb_b200's builders bit for bit.
//
// What this file does itself is what geometry::GeometryReader (Code/geometry/GeometryReader.cc:85-160,
// 556-650) does between the .gmy file and the Domain constructor: fill a GmyReadResult (one
// GeometrySite per lattice site of every non-empty block, links matched to the lattice in use),
// build the block octree, run the basic decomposition, create the distributed store, assign
// targetProcessor.  File reading and ParMETIS are not involved.
#include <mpi.h>

#include <array>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "geometry/Domain.h"
#include "geometry/GmyReadResult.h"
#include "geometry/LookupTree.h"
#include "geometry/decomposition/BasicDecomposition.h"
#include "debug/Debugger.h"
#include "io/formats/geometry.h"
#include "lb/lattices/D3Q15.h"
#include "lb/lattices/D3Q19.h"
#include "lb/lattices/D3Q27.h"
#include "net/IOCommunicator.h"
#include "reporting/Dict.h"
#include "ref_domain_build.h"

namespace hemelb::reporting {
  // reporting/Dict.cc wraps ctemplate (absent here); geometry::Domain::Report is not on the path
  Dict::Dict(const std::string&) : raw(nullptr, [](ctemplate::TemplateDictionary*) {}) {}
  Dict::Dict(ctemplate::TemplateDictionary*) : raw(nullptr, [](ctemplate::TemplateDictionary*) {}) {}
  Dict Dict::AddSectionDictionary(const std::string& s) { return Dict(s); }
  void Dict::SetValue(const std::string&, const std::string&) {}
  void Dict::SetIntValue(const std::string&, long) {}
  void Dict::SetBoolValue(const std::string&, bool) {}
}

namespace hemelb::debug {
  // net/MpiError.cc asks the debugger to break before it throws; here: nothing to attach to
  namespace {
    struct NoDebugger : Debugger {
      NoDebugger() : Debugger(nullptr, net::MpiCommunicator()) {}
      void BreakHere() override {}
      void Print(const char*, ...) override {}
      void Attach() override {}
    };
  }
  Debugger::Debugger(const char*, net::MpiCommunicator c) : mCommunicator(c) {}
  Debugger* Debugger::Get() {
    static NoDebugger none;
    return &none;
  }
}

namespace hemelb::tests::helpers {
  // the reference's own friend hook for tests (geometry/Domain.h:39-42,58)
  class LatticeDataAccess {
   public:
    explicit LatticeDataAccess(geometry::Domain const& d) : dom(d) {}
    auto const& NeighbourIndices() const { return dom.neighbourIndices; }
    auto const& NeighbouringProcs() const { return dom.neighbouringProcs; }
    auto const& StreamingIndices() const { return dom.streamingIndicesForReceivedDistributions; }
    auto const& DistanceToWall() const { return dom.distanceToWall; }
    auto const& WallNormals() const { return dom.wallNormalAtSite; }
    auto const& SiteDatas() const { return dom.siteData; }
    auto const& GlobalCoords() const { return dom.globalSiteCoords; }
    site_t TotalSharedFs() const { return dom.totalSharedFs; }
   private:
    geometry::Domain const& dom;
  };
}

namespace {
using namespace hemelb;

struct RankTables {
  std::vector<int64_t> counts, neighbourIndices, globalCoords, streamingIndices, procs;
  std::vector<uint32_t> wallMask, ioletMask;
  std::vector<int32_t> siteType, ioletId;
  std::vector<double> distanceToWall, wallNormal;
  int64_t N = 0, totalSharedFs = 0;
};
struct Run {
  int Q = 0, R = 0;
  refdom::GeometryArrays g;
  std::vector<RankTables> out;
  std::vector<int32_t> blockRank;     // per .gmy block: rank given by BasicDecomposition, or SITE_OR_BLOCK_SOLID
  std::string error;
};

lb::LatticeInfo const& lattice_info(int Q) {
  switch (Q) {
    case 15: return lb::D3Q15::GetLatticeInfo();
    case 19: return lb::D3Q19::GetLatticeInfo();
    default: return lb::D3Q27::GetLatticeInfo();
  }
}

void rank_body(int rank, void* arg) {
  Run& run = *static_cast<Run*>(arg);
  auto const& info = lattice_info(run.Q);
  net::IOCommunicator comms{net::MpiCommunicator::World()};

  std::vector<proc_t> procForEachBlock;
  geometry::GmyReadResult read = refdom::BuildReadResult(run.g, info, comms, &procForEachBlock);
  if (rank == 0) run.blockRank.assign(procForEachBlock.begin(), procForEachBlock.end());

  {
    geometry::Domain domain(info, read, comms);
    tests::helpers::LatticeDataAccess access(domain);
    RankTables& T = run.out[rank];
    const int Q = run.Q;
    T.N = domain.GetLocalFluidSiteCount();
    T.totalSharedFs = access.TotalSharedFs();
    for (unsigned t = 0; t < COLLISION_TYPES; ++t) T.counts.push_back(domain.GetMidDomainCollisionCount(t));
    for (unsigned t = 0; t < COLLISION_TYPES; ++t) T.counts.push_back(domain.GetDomainEdgeCollisionCount(t));
    T.neighbourIndices.assign(access.NeighbourIndices().begin(), access.NeighbourIndices().end());
    T.streamingIndices.assign(access.StreamingIndices().begin(), access.StreamingIndices().end());
    for (auto const& p : access.NeighbouringProcs()) {
      T.procs.push_back(p.Rank);
      T.procs.push_back(p.SharedDistributionCount);
      T.procs.push_back(p.FirstSharedDistribution);
    }
    T.distanceToWall.assign(access.DistanceToWall().begin(), access.DistanceToWall().end());
    for (site_t i = 0; i < T.N; ++i) {
      auto const& sd = access.SiteDatas()[i];
      T.wallMask.push_back(sd.GetWallIntersectionData());
      T.ioletMask.push_back(sd.GetIoletIntersectionData());
      T.siteType.push_back((int32_t)sd.GetSiteType());
      T.ioletId.push_back(sd.GetIoletId());
      for (int k = 0; k < 3; ++k) {
        T.wallNormal.push_back(access.WallNormals()[i][k]);
        T.globalCoords.push_back(access.GlobalCoords()[i][k]);
      }
    }
    (void)Q;
  }  // (the Domain's windows are freed collectively here)
}

}  // namespace

extern "C" {

void* hrefdom_run(int Q, int R, const int32_t* blockDims, int blockSize, int64_t N, const int32_t* coords, int64_t nb,
                  const int64_t* bsite, const uint8_t* btype, const int32_t* biolet, const float* bdist,
                  const uint8_t* bnavail, const float* bnormal, const int32_t* siteRank) {
  {
    // decomposition::BasicDecomposition throws when there are more ranks than non-empty blocks
    // (BasicDecomposition.cc:66-67); answered here, before any thread exists that could be left at a barrier
    std::vector<char> seen((size_t)blockDims[0] * blockDims[1] * blockDims[2], 0);
    int64_t nonEmpty = 0;
    for (int64_t s = 0; s < N; ++s) {
      const size_t b = ((size_t)(coords[3 * s] / blockSize) * blockDims[1] + coords[3 * s + 1] / blockSize) * blockDims[2] +
                       coords[3 * s + 2] / blockSize;
      if (!seen[b]) {
        seen[b] = 1;
        ++nonEmpty;
      }
    }
    if (nonEmpty < R) return nullptr;
  }
  Run* run = new Run;
  run->Q = Q;
  run->R = R;
  run->g.blockSize = blockSize;
  for (int k = 0; k < 3; ++k) run->g.bd[k] = blockDims[k];
  run->g.N = N;
  run->g.nb = nb;
  run->g.coords = coords;
  run->g.bsite = bsite;
  run->g.btype = btype;
  run->g.biolet = biolet;
  run->g.bdist = bdist;
  run->g.bnavail = bnavail;
  run->g.bnormal = bnormal;
  run->g.siteRank = siteRank;
  run->out.resize(R);
  fakempi_run(R, rank_body, run);
  return run;
}

// name -> number of elements; data copied to `out` when it is not null
int64_t hrefdom_get(void* h, int rank, const char* name, void* out) {
  Run& run = *static_cast<Run*>(h);
  const std::string n(name);
  auto give = [&](auto const& v) -> int64_t {
    if (out && !v.empty()) std::memcpy(out, v.data(), v.size() * sizeof(v[0]));
    return (int64_t)v.size();
  };
  if (n == "blockRank") return give(run.blockRank);
  RankTables& T = run.out[rank];
  if (n == "N") return T.N;
  if (n == "totalSharedFs") return T.totalSharedFs;
  if (n == "counts") return give(T.counts);
  if (n == "neighbourIndices") return give(T.neighbourIndices);
  if (n == "streamingIndices") return give(T.streamingIndices);
  if (n == "procs") return give(T.procs);
  if (n == "wallMask") return give(T.wallMask);
  if (n == "ioletMask") return give(T.ioletMask);
  if (n == "siteType") return give(T.siteType);
  if (n == "ioletId") return give(T.ioletId);
  if (n == "distanceToWall") return give(T.distanceToWall);
  if (n == "wallNormal") return give(T.wallNormal);
  if (n == "globalCoords") return give(T.globalCoords);
  return -1;
}

void hrefdom_destroy(void* h) { delete static_cast<Run*>(h); }

}  // extern "C"
