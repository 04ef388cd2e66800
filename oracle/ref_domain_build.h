// TEST INFRASTRUCTURE ONLY.  What geometry::GeometryReader (Code/geometry/GeometryReader.cc:85-160, 556-650) does
// between the .gmy file and the geometry::Domain constructor, from arrays instead of a file: fill a
// GmyReadResult (one GeometrySite per lattice site of every non-empty block, links matched to the lattice
// in use), build the block octree, run the reference's BasicDecomposition, create the DistributedStore,
// assign targetProcessor.  Used by oracle/ref_domain_driver.cc (tables against the reference's Domain)
// and tests/host_lbm_real.cc (the real lb::LBM over the real Domain driving the GPU policy classes).
#pragma once
#include <cstdint>
#include <memory>
#include <vector>

#include "geometry/GmyReadResult.h"
#include "geometry/LookupTree.h"
#include "geometry/decomposition/BasicDecomposition.h"
#include "io/formats/geometry.h"
#include "lb/lattices/LatticeInfo.h"
#include "net/IOCommunicator.h"

namespace refdom {
  using namespace hemelb;

  struct GeometryArrays {
    int blockSize = 0;
    int bd[3] = {0, 0, 0};
    int64_t N = 0, nb = 0;
    const int32_t* coords = nullptr;   // (N, 3) global voxel coordinates, .gmy order
    const int64_t* bsite = nullptr;    // (nb) input site of every boundary record
    const uint8_t* btype = nullptr;    // (nb, 26) cut types in the file's neighbourhood order
    const int32_t* biolet = nullptr;   // (nb, 26)
    const float* bdist = nullptr;      // (nb, 26)
    const uint8_t* bnavail = nullptr;  // (nb)
    const float* bnormal = nullptr;    // (nb, 3)
    const int32_t* siteRank = nullptr; // (N) or null: the reference's BasicDecomposition over blocks
  };

  // collective over comms (the DistributedStore allocates its window)
  inline geometry::GmyReadResult BuildReadResult(const GeometryArrays& run, const lb::LatticeInfo& info,
                                                 const net::IOCommunicator& comms, std::vector<proc_t>* blockRank) {
    using gmy = io::formats::geometry;
    const int B = run.blockSize;
    geometry::GmyReadResult read(Vec16(run.bd[0], run.bd[1], run.bd[2]), U16(B));
    read.Blocks.resize(read.GetBlockCount());
    std::vector<site_t> fluidSitesPerBlock(read.GetBlockCount(), 0);
    // which input site carries which boundary record
    std::vector<int64_t> recordOf(run.N, -1);
    for (int64_t k = 0; k < run.nb; ++k) recordOf[run.bsite[k]] = k;
    auto block_of = [&](int64_t s) {
      return read.GetBlockIdFromBlockCoordinates(run.coords[3 * s] / B, run.coords[3 * s + 1] / B, run.coords[3 * s + 2] / B);
    };
    for (int64_t s = 0; s < run.N; ++s) {
      const site_t blk = block_of(s);
      auto& sites = read.Blocks[blk].Sites;
      if (sites.empty()) sites.assign(read.GetSitesPerBlock(), geometry::GeometrySite(false));
      const site_t local = read.GetSiteIdFromSiteCoordinates(run.coords[3 * s] % B, run.coords[3 * s + 1] % B, run.coords[3 * s + 2] % B);
      geometry::GeometrySite site(true);
      site.links.resize(info.GetNumVectors() - 1);
      const int64_t rec = recordOf[s];
      if (rec >= 0) {
        int n = 0;
        for (auto&& dir : gmy::Neighbourhood) {  // the file's 26 directions, matched to the lattice in use
          geometry::GeometrySiteLink link;
          link.type = static_cast<gmy::CutType>(run.btype[rec * 26 + n]);
          if (link.type != gmy::CutType::NONE) {
            link.distanceToIntersection = run.bdist[rec * 26 + n];
            if (link.type != gmy::CutType::WALL) link.ioletId = run.biolet[rec * 26 + n];
          }
          for (Direction l = 1; l < info.GetNumVectors(); ++l)
            if (info.GetVector(l) == dir) {
              site.links[l - 1] = link;
              break;
            }
          ++n;
        }
        site.wallNormalAvailable = run.bnavail[rec] != 0;
        if (site.wallNormalAvailable)
          site.wallNormal = util::Vector3D<float>(run.bnormal[3 * rec], run.bnormal[3 * rec + 1], run.bnormal[3 * rec + 2]);
      }
      sites[local] = site;
      ++fluidSitesPerBlock[blk];
    }

    auto blockTree = geometry::octree::build_block_tree(read.GetBlockDimensions().as<geometry::octree::U16>(), fluidSitesPerBlock);
    std::vector<proc_t> procForEachBlock(read.GetBlockCount());
    geometry::decomposition::BasicDecomposition basic(read, comms.Size());
    auto procForBlockOct = basic.Decompose(blockTree, procForEachBlock);
    read.block_store = std::make_unique<geometry::octree::DistributedStore>(read.GetSitesPerBlock(), std::move(blockTree),
                                                                                procForBlockOct, comms);
    for (int64_t s = 0; s < run.N; ++s) {
      const site_t blk = block_of(s);
      const site_t local = read.GetSiteIdFromSiteCoordinates(run.coords[3 * s] % B, run.coords[3 * s + 1] % B, run.coords[3 * s + 2] % B);
      read.Blocks[blk].Sites[local].targetProcessor = run.siteRank ? run.siteRank[s] : procForEachBlock[blk];
    }
    if (blockRank) *blockRank = procForEachBlock;
    return read;
  }
}
