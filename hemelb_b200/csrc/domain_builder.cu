// Device-side construction of one rank's geometry::Domain tables (sm_100a).
//
// What the reference does on the host in Code/geometry/Domain.cc:69-580 (site order, collision-type
// buckets, neighbourIndices, neighbouringProcs, halo send slots, streamingIndicesForReceived-
// Distributions) for the sites GeometryReader hands it (Code/geometry/GeometryReader.cc:556-652),
// done here with O(N) kernels over a dense voxel window in HBM, so that a 1e8-site rank is ready in
// seconds and the N*Q int64 table never exists on the host.  Two site sources feed the same table
// code:
//   * explicit -- the .gmy-level site list + cut-link records (any order, any site -> rank map);
//   * analytic -- a union of capsules clipped by flat iolet caps, voxelised on the device with the
//                 same link model as hemelb_b200/geometry.py:voxelise (wall crossing by 30-step
//                 bisection, plane crossing analytically, distances rounded to float32 as a .gmy
//                 stores them), with a site -> rank rule (slabs or a block table).
// Only the O(halo) part -- ordering the cut links into the per-neighbour send slices with the
// reference's pair protocol (lower rank's list is authoritative, Domain.cc:530-565) -- runs on the
// host, from the compact list of remote links the device emits.
//
// Grid codes (int32 per voxel of the window): -1 solid, -2-r fluid owned by another rank r,
// >= 0 an own site (first its traversal index, finally its local site id).
#include <cuda_runtime.h>
#include <cub/cub.cuh>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "../../include/hemelb_b200.h"
#include "engine_internal.h"
#include "lattice.cuh"

using namespace hlb;

namespace {

int fail(const std::string& m) { return hlb_internal_fail(m.c_str()); }
#define CU(call)                                                                     \
  do {                                                                               \
    cudaError_t e_ = (call);                                                         \
    if (e_ != cudaSuccess) return fail(std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

enum { CUT_NONE = 0, CUT_WALL = 1, CUT_INLET = 2, CUT_OUTLET = 3 };
enum { PART_SINGLE = 0, PART_SLABS = 1, PART_BLOCKS = 2, PART_EXPLICIT = 3 };
constexpr int kMaxSiteCand = 8;
constexpr int kMaxBlockCand = 32;

struct LatticeTab {
  int Q;
  int8_t c[27][3];
  int8_t link[27];  // index into the .gmy 26-neighbourhood (Code/io/formats/geometry.h:120-156)
  int8_t inv[27];
};

__host__ __device__ inline int link_index(int i, int j, int k) {
  int idx = (i + 1) * 9 + (j + 1) * 3 + (k + 1);
  return idx > 13 ? idx - 1 : idx;
}
__host__ __device__ inline void link_vector(int l, int& i, int& j, int& k) {
  const int idx = l >= 13 ? l + 1 : l;
  i = idx / 9 - 1;
  j = (idx / 3) % 3 - 1;
  k = idx % 3 - 1;
}

template <int Q> void fill_lattice(LatticeTab& T) {
  T.Q = Q;
  for (int d = 0; d < Q; ++d) {
    T.c[d][0] = (int8_t)Lat<Q>::cx(d);
    T.c[d][1] = (int8_t)Lat<Q>::cy(d);
    T.c[d][2] = (int8_t)Lat<Q>::cz(d);
    T.inv[d] = (int8_t)inv_dir(d);
    T.link[d] = d ? (int8_t)link_index(T.c[d][0], T.c[d][1], T.c[d][2]) : 0;
  }
}

struct Win {  // dense voxel window: grid index of global (x,y,z)
  int64_t org[3], dim[3];
  __host__ __device__ int64_t key(int64_t x, int64_t y, int64_t z) const {
    return ((x - org[0]) * dim[1] + (y - org[1])) * dim[2] + (z - org[2]);
  }
  __host__ __device__ bool holds(int64_t x, int64_t y, int64_t z) const {
    return x >= org[0] && x < org[0] + dim[0] && y >= org[1] && y < org[1] + dim[1] && z >= org[2] &&
           z < org[2] + dim[2];
  }
  __host__ __device__ int64_t offset(int cx, int cy, int cz) const { return ((int64_t)cx * dim[1] + cy) * dim[2] + cz; }
};

struct Box3 { int64_t lo[3], dim[3]; };  // a box of blocks

struct Part {
  int mode, me, nranks, axis, B;
  int64_t bd[3];
  const int64_t* first;        // SLABS: nranks + 1 ascending coordinates along `axis`
  const int16_t* rankOfBlock;  // BLOCKS: .gmy block index -> rank
  __device__ int rank_of(int64_t x, int64_t y, int64_t z) const {
    if (mode == PART_SINGLE) return 0;
    if (mode == PART_SLABS) {
      const int64_t v = axis == 0 ? x : (axis == 1 ? y : z);
      int r = 0;
      while (r + 1 < nranks && v >= first[r + 1]) ++r;
      return r;
    }
    const int64_t b = ((x / B) * bd[1] + (y / B)) * bd[2] + (z / B);
    return rankOfBlock[b];
  }
};

struct Capsule { double a[3], ab[3], L2, r, rough; };  // rough: amplitude of the wall roughness (voxels), 0 = smooth
struct IoletPlane { int kind, index; double pos[3], n[3], radius; };
struct Shape {
  const Capsule* caps;
  int nCaps;
  const IoletPlane* iolets;
  int nIolets;
  // seeded value noise on an ng^3 grid stretched over the lattice (geometry.py:sac): the surface of a
  // rough capsule is displaced by rough * noise(p); wall normals then come from central differences
  const double* noise;
  int ng;
  double ext[3];
  double maxRough;
};

struct Explicit {
  const int32_t* coords;      // n x 3
  const int32_t* rankOfSite;  // n, or null (all on rank 0)
  const int32_t* recOfInput;  // n: record index or -1
  const uint8_t* type;        // nrec x 26
  const int32_t* iolet;       // nrec x 26
  const float* dist;          // nrec x 26
  const uint8_t* navail;      // nrec
  const float* normal;        // nrec x 3
  int64_t n;
};

// ------------------------------------------------------------------------------- analytic shape
// signed distance to the smooth capsule
__device__ __forceinline__ double capsule_phi(const Capsule& c, double x, double y, double z) {
  const double d0 = x - c.a[0], d1 = y - c.a[1], d2 = z - c.a[2];
  double t = 0.0;
  if (c.L2 > 0.0) {
    t = (d0 * c.ab[0] + d1 * c.ab[1] + d2 * c.ab[2]) / c.L2;
    t = t < 0.0 ? 0.0 : (t > 1.0 ? 1.0 : t);
  }
  const double e0 = d0 - t * c.ab[0], e1 = d1 - t * c.ab[1], e2 = d2 - t * c.ab[2];
  return sqrt(e0 * e0 + e1 * e1 + e2 * e2) - c.r;
}
// geometry.py:sac.value_noise -- trilinear interpolation of the seeded grid, corners in its order
__device__ __forceinline__ double value_noise(const Shape& S, double x, double y, double z) {
  const double p[3] = {x, y, z};
  int i0[3];
  double f[3];
  for (int k = 0; k < 3; ++k) {
    const double q = (p[k] / S.ext[k]) * (S.ng - 1);
    int i = (int)floor(q);
    i = i < 0 ? 0 : (i > S.ng - 2 ? S.ng - 2 : i);
    i0[k] = i;
    f[k] = q - i;
  }
  double acc = 0.0;
  for (int dx = 0; dx < 2; ++dx)
    for (int dy = 0; dy < 2; ++dy)
      for (int dz = 0; dz < 2; ++dz) {
        const double w = (dx ? f[0] : 1 - f[0]) * (dy ? f[1] : 1 - f[1]) * (dz ? f[2] : 1 - f[2]);
        acc += w * S.noise[((i0[0] + dx) * S.ng + i0[1] + dy) * S.ng + i0[2] + dz];
      }
  return acc;
}
// the capsule's implicit function: signed distance, displaced by the roughness
__device__ __forceinline__ double capsule_phi(const Shape& S, const Capsule& c, double x, double y, double z) {
  double v = capsule_phi(c, x, y, z);
  if (c.rough != 0.0) v = v - c.rough * value_noise(S, x, y, z);
  return v;
}
// min over a candidate list (nCand < 0: every capsule)
__device__ __forceinline__ double shape_phi(const Shape& S, const int* cand, int nCand, double x, double y, double z) {
  double out = INFINITY;
  if (nCand < 0) {
    for (int k = 0; k < S.nCaps; ++k) out = fmin(out, capsule_phi(S, S.caps[k], x, y, z));
  } else {
    for (int k = 0; k < nCand; ++k) out = fmin(out, capsule_phi(S, S.caps[cand[k]], x, y, z));
  }
  return out;
}
// geometry.py:voxelise.clipped -- first iolet cap whose outside half-space holds p, or -1
__device__ __forceinline__ int clipped(const Shape& S, double x, double y, double z) {
  for (int k = 0; k < S.nIolets; ++k) {
    const IoletPlane& io = S.iolets[k];
    const double d0 = x - io.pos[0], d1 = y - io.pos[1], d2 = z - io.pos[2];
    const double h = d0 * io.n[0] + d1 * io.n[1] + d2 * io.n[2];
    const double r2 = (d0 * d0 + d1 * d1 + d2 * d2) - h * h;
    const double rr = io.radius * 1.5;
    if (h <= 0.0 && r2 <= rr * rr && h > -4.0 - io.radius) return k;
  }
  return -1;
}
__device__ __forceinline__ int site_candidates(const Shape& S, double x, double y, double z, int* cand) {
  int n = 0;
  for (int k = 0; k < S.nCaps; ++k)
    if (capsule_phi(S.caps[k], x, y, z) < 2.5 + fabs(S.caps[k].rough)) {
      if (n == kMaxSiteCand) return -1;
      cand[n++] = k;
    }
  return n;
}
struct LinkRes { int type, id; float dist; };
// one cut link of a fluid site a towards the non-fluid voxel a + c (geometry.py:voxelise)
__device__ LinkRes analytic_link(const Shape& S, const int* cand, int nCand, double ax, double ay, double az, int cx,
                                 int cy, int cz) {
  const double bx = ax + cx, by = ay + cy, bz = az + cz;
  double tWall = INFINITY;
  if (shape_phi(S, cand, nCand, bx, by, bz) >= 0.0) {
    double lo = 0.0, hi = 1.0;
    for (int it = 0; it < 30; ++it) {
      const double mid = 0.5 * (lo + hi);
      const bool inside = shape_phi(S, cand, nCand, ax + mid * cx, ay + mid * cy, az + mid * cz) < 0.0;
      lo = inside ? mid : lo;
      hi = inside ? hi : mid;
    }
    tWall = hi;
  }
  double tIo = INFINITY;
  const int k = clipped(S, bx, by, bz);
  if (k >= 0) {
    const IoletPlane& io = S.iolets[k];
    const double h0 = (ax - io.pos[0]) * io.n[0] + (ay - io.pos[1]) * io.n[1] + (az - io.pos[2]) * io.n[2];
    const double dh = cx * io.n[0] + cy * io.n[1] + cz * io.n[2];
    tIo = h0 / (-dh);
  }
  const bool isIo = tIo < tWall;
  double t = isIo ? tIo : tWall;
  t = t < 1e-6 ? 1e-6 : (t > 1.0 ? 1.0 : t);
  LinkRes r;
  r.type = isIo ? S.iolets[k].kind : CUT_WALL;
  r.id = isIo ? S.iolets[k].index : -1;
  r.dist = (float)t;
  return r;
}
// wall normal at a site: the exact gradient of the signed distance to the nearest capsule (radially
// away from its axis), the analytic counterpart of geometry.py's finite-difference default
__device__ void analytic_normal(const Shape& S, const int* cand, int nCand, double x, double y, double z, float* out) {
  if (S.maxRough != 0.0) {
    // geometry.py:voxelise's default normal_fn: central differences of phi over +-0.25 voxel
    double g[3];
    g[0] = shape_phi(S, cand, nCand, x + 0.25, y, z) - shape_phi(S, cand, nCand, x - 0.25, y, z);
    g[1] = shape_phi(S, cand, nCand, x, y + 0.25, z) - shape_phi(S, cand, nCand, x, y - 0.25, z);
    g[2] = shape_phi(S, cand, nCand, x, y, z + 0.25) - shape_phi(S, cand, nCand, x, y, z - 0.25);
    double len = sqrt(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
    if (len == 0.0) len = 1.0;
    for (int k = 0; k < 3; ++k) out[k] = (float)(g[k] / len);
    return;
  }
  double best = INFINITY;
  int kb = 0;
  const int n = nCand < 0 ? S.nCaps : nCand;
  for (int i = 0; i < n; ++i) {
    const int k = nCand < 0 ? i : cand[i];
    const double p = capsule_phi(S.caps[k], x, y, z);
    if (p < best) { best = p; kb = k; }
  }
  const Capsule& c = S.caps[kb];
  const double d0 = x - c.a[0], d1 = y - c.a[1], d2 = z - c.a[2];
  double t = 0.0;
  if (c.L2 > 0.0) {
    t = (d0 * c.ab[0] + d1 * c.ab[1] + d2 * c.ab[2]) / c.L2;
    t = t < 0.0 ? 0.0 : (t > 1.0 ? 1.0 : t);
  }
  const double g[3] = {d0 - t * c.ab[0], d1 - t * c.ab[1], d2 - t * c.ab[2]};
  double len = sqrt(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
  if (len == 0.0) len = 1.0;
  for (int k = 0; k < 3; ++k) out[k] = (float)(g[k] / len);
}

__host__ __device__ inline uint64_t spread21(uint64_t v) {
  v &= 0x1fffffull;
  v = (v | v << 32) & 0x1f00000000ffffull;
  v = (v | v << 16) & 0x1f0000ff0000ffull;
  v = (v | v << 8) & 0x100f00f00f00f00full;
  v = (v | v << 4) & 0x10c30c30c30c30c3ull;
  v = (v | v << 2) & 0x1249249249249249ull;
  return v;
}
// octree id of block coordinates, x most significant (Code/geometry/LookupTree.h:92-97)
__host__ __device__ inline uint64_t morton3(int64_t x, int64_t y, int64_t z) {
  return (spread21((uint64_t)x) << 2) | (spread21((uint64_t)y) << 1) | spread21((uint64_t)z);
}

// ------------------------------------------------------------------------------- kernels
// analytic source, one CTA per block of `box`: fluid test per voxel; COUNT_ONLY: fluid sites per
// block (any owner); else: grid code of every fluid voxel inside the window + own sites per block
template <bool COUNT_ONLY>
__global__ void __launch_bounds__(512) classify_blocks_kernel(Shape S, Part P, Win W, Box3 box, int32_t* __restrict__ grid,
                                                             int32_t* __restrict__ blockCount,
                                                             int32_t* __restrict__ boundaryCount = nullptr,
                                                             const LatticeTab* __restrict__ lat = nullptr) {
  __shared__ int cand[kMaxBlockCand];
  __shared__ int nCandS;
  const int B = P.B;
  const int64_t bi = blockIdx.x;
  const int64_t bx = box.lo[0] + bi / (box.dim[1] * box.dim[2]);
  const int64_t by = box.lo[1] + (bi / box.dim[2]) % box.dim[1];
  const int64_t bz = box.lo[2] + bi % box.dim[2];
  const int t = threadIdx.x;
  if (t == 0) nCandS = 0;
  __syncthreads();
  const double h = 0.5 * (B - 1);
  const double cxx = bx * B + h, cyy = by * B + h, czz = bz * B + h;
  // (+ one voxel of rim when the lattice neighbours of the block's sites are classified too)
  const double reach = 1.7320508075688772 * (h + (COUNT_ONLY && boundaryCount ? 1.0 : 0.0)) + 0.5;
  for (int k = t; k < S.nCaps; k += blockDim.x)
    if (capsule_phi(S.caps[k], cxx, cyy, czz) < reach + fabs(S.caps[k].rough)) {
      const int i = atomicAdd(&nCandS, 1);
      if (i < kMaxBlockCand) cand[i] = k;
    }
  __syncthreads();
  int nCand = nCandS;
  if (nCand == 0) {
    if (t == 0) {
      blockCount[bi] = 0;
      if (COUNT_ONLY && boundaryCount) boundaryCount[bi] = 0;
    }
    return;
  }
  if (nCand > kMaxBlockCand) nCand = -1;
  bool flag = false;
  if (t < B * B * B) {
    const int64_t x = bx * B + t / (B * B), y = by * B + (t / B) % B, z = bz * B + t % B;
    const bool inLattice = x >= 0 && y >= 0 && z >= 0 && x < P.bd[0] * B && y < P.bd[1] * B && z < P.bd[2] * B;
    if (inLattice && shape_phi(S, cand, nCand, (double)x, (double)y, (double)z) < 0.0 &&
        clipped(S, (double)x, (double)y, (double)z) < 0) {
      if (COUNT_ONLY) {
        flag = true;
      } else {
        const int r = P.rank_of(x, y, z);
        flag = r == P.me;
        if (W.holds(x, y, z)) grid[W.key(x, y, z)] = -2 - r;
      }
    }
  }
  const int c = __syncthreads_count(flag);
  if (t == 0) blockCount[bi] = c;
  if constexpr (COUNT_ONLY) {
    if (boundaryCount) {
      // boundary-typed = some lattice link leaves the fluid (wall or iolet cut): Domain.cc:186-207
      bool cut = false;
      if (flag) {
        const int64_t x = bx * B + t / (B * B), y = by * B + (t / B) % B, z = bz * B + t % B;
        for (int d = 1; d < lat->Q && !cut; ++d) {
          const double nx = (double)(x + lat->c[d][0]), ny = (double)(y + lat->c[d][1]), nz = (double)(z + lat->c[d][2]);
          const bool inL = nx >= 0 && ny >= 0 && nz >= 0 && nx < P.bd[0] * B && ny < P.bd[1] * B && nz < P.bd[2] * B;
          cut = !(inL && shape_phi(S, cand, nCand, nx, ny, nz) < 0.0 && clipped(S, nx, ny, nz) < 0);
        }
      }
      const int cb = __syncthreads_count(cut);
      if (t == 0) boundaryCount[bi] = cb;
    }
  }
}

// explicit source: drop every uploaded site inside the window into the grid (its input index) and
// count own sites per block of `box`
__global__ void mark_sites_kernel(Explicit E, Part P, Win W, Box3 box, int32_t* __restrict__ grid,
                                  int32_t* __restrict__ blockCount, int* __restrict__ bad) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= E.n) return;
  const int64_t x = E.coords[3 * i], y = E.coords[3 * i + 1], z = E.coords[3 * i + 2];
  const int r = E.rankOfSite ? E.rankOfSite[i] : 0;
  if (!W.holds(x, y, z)) {
    if (r == P.me) atomicExch(bad, 1);
    return;
  }
  const int32_t prev = atomicExch(&grid[W.key(x, y, z)], (int32_t)i);
  if (prev != -1) atomicExch(bad, 2);  // two sites at one voxel
  if (r == P.me) {
    const int64_t bx = x / P.B - box.lo[0], by = y / P.B - box.lo[1], bz = z / P.B - box.lo[2];
    atomicAdd(&blockCount[(bx * box.dim[1] + by) * box.dim[2] + bz], 1);
  }
}

__global__ void block_keys_kernel(Box3 box, Part P, const int32_t* __restrict__ blockCount, uint64_t* __restrict__ keys,
                                  uint32_t* __restrict__ vals, int64_t nb) {
  const int64_t bi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (bi >= nb) return;
  const int64_t bx = box.lo[0] + bi / (box.dim[1] * box.dim[2]);
  const int64_t by = box.lo[1] + (bi / box.dim[2]) % box.dim[1];
  const int64_t bz = box.lo[2] + bi % box.dim[2];
  const bool in = bx >= 0 && by >= 0 && bz >= 0 && bx < P.bd[0] && by < P.bd[1] && bz < P.bd[2];
  keys[bi] = (in && blockCount[bi] > 0) ? morton3(bx, by, bz) : ~0ull;
  vals[bi] = (uint32_t)bi;
}
__global__ void gather_counts_kernel(const uint32_t* __restrict__ sortedBlocks, const int32_t* __restrict__ blockCount,
                                     int64_t* __restrict__ out, int64_t nb) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nb) out[i] = blockCount[sortedBlocks[i]];
}
__global__ void scatter_starts_kernel(const uint32_t* __restrict__ sortedBlocks, const int64_t* __restrict__ scanned,
                                      int64_t* __restrict__ blockStart, int64_t nb) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nb) blockStart[sortedBlocks[i]] = scanned[i];
}

// one CTA per block of `box`: own sites of the block, in x-major / z-fastest order
// (Code/geometry/VolumeTraverser.cc:27-52), get consecutive traversal indices from blockStart
template <bool ANALYTIC>
__global__ void __launch_bounds__(512) compact_kernel(Explicit E, Part P, Win W, Box3 box, int32_t* __restrict__ grid,
                                                     const int64_t* __restrict__ blockStart, int64_t nTrav,
                                                     int32_t* __restrict__ coordsTrav, int32_t* __restrict__ inputOfTrav) {
  typedef cub::BlockScan<int, 512> Scan;
  __shared__ typename Scan::TempStorage tmp;
  const int B = P.B;
  const int64_t bi = blockIdx.x;
  const int64_t bx = box.lo[0] + bi / (box.dim[1] * box.dim[2]);
  const int64_t by = box.lo[1] + (bi / box.dim[2]) % box.dim[1];
  const int64_t bz = box.lo[2] + bi % box.dim[2];
  const int t = threadIdx.x;
  int64_t x = 0, y = 0, z = 0, key = -1;
  int32_t g = -1;
  if (t < B * B * B) {
    x = bx * B + t / (B * B);
    y = by * B + (t / B) % B;
    z = bz * B + t % B;
    if (W.holds(x, y, z)) {
      key = W.key(x, y, z);
      g = grid[key];
    }
  }
  bool own;
  int r = -1;
  if (ANALYTIC) {
    own = g == -2 - P.me;
  } else {
    if (g >= 0) r = E.rankOfSite ? E.rankOfSite[g] : 0;
    own = g >= 0 && r == P.me;
  }
  int pos = 0;
  Scan(tmp).ExclusiveSum(own ? 1 : 0, pos);
  if (own) {
    const int64_t tr = blockStart[bi] + pos;
    coordsTrav[tr] = (int32_t)x;
    coordsTrav[nTrav + tr] = (int32_t)y;
    coordsTrav[2 * nTrav + tr] = (int32_t)z;
    if (!ANALYTIC) inputOfTrav[tr] = g;
    grid[key] = (int32_t)tr;
  } else if (!ANALYTIC && g >= 0) {
    grid[key] = -2 - r;
  }
}

struct SiteMasks { uint32_t wall, iolet; bool hadIn, hadOut; };
__device__ __forceinline__ int collision_type(const SiteMasks& m) {
  // SiteDataBare.cc:23-140 -> the six collision types of lb::LBM (lb.h:102-107)
  const int type = m.hadIn ? 2 : (m.hadOut ? 3 : 1);
  if (m.wall == 0) return type == 1 ? 0 : (type == 2 ? 2 : 3);
  return type == 1 ? 1 : (type == 2 ? 4 : 5);
}

// per own site (traversal order): edge test + collision type -> bucket; histogram; remote links
template <bool ANALYTIC>
__global__ void __launch_bounds__(256) classify_sites_kernel(Shape S, Explicit E, Part P, Win W, LatticeTab L,
                                                            const int32_t* __restrict__ grid,
                                                            const int32_t* __restrict__ coordsTrav,
                                                            const int32_t* __restrict__ inputOfTrav, int64_t nTrav,
                                                            uint8_t* __restrict__ bucket,
                                                            unsigned long long* __restrict__ hist /*12 + 1*/) {
  __shared__ unsigned int sh[13];
  if (threadIdx.x < 13) sh[threadIdx.x] = 0;
  __syncthreads();
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < nTrav) {
    const int64_t x = coordsTrav[t], y = coordsTrav[nTrav + t], z = coordsTrav[2 * nTrav + t];
    const int64_t key = W.key(x, y, z);
    SiteMasks m = {0u, 0u, false, false};
    bool edge = false;
    int nRemote = 0;
    int64_t rec = -1;
    if (!ANALYTIC) rec = E.recOfInput[inputOfTrav[t]];
    int cand[kMaxSiteCand];
    int nCand = -2;  // not computed yet
    for (int d = 1; d < L.Q; ++d) {
      const int32_t g = grid[key + W.offset(L.c[d][0], L.c[d][1], L.c[d][2])];
      if (g <= -2 && g != -2 - P.me) {
        edge = true;
        ++nRemote;
      }
      int ty = CUT_NONE;
      if (ANALYTIC) {
        if (g == -1) {
          if (nCand == -2) nCand = site_candidates(S, (double)x, (double)y, (double)z, cand);
          ty = analytic_link(S, cand, nCand, (double)x, (double)y, (double)z, L.c[d][0], L.c[d][1], L.c[d][2]).type;
        }
      } else if (rec >= 0) {
        ty = E.type[rec * 26 + L.link[d]];
      }
      const uint32_t bit = 1u << (d - 1);
      if (ty == CUT_WALL) m.wall |= bit;
      else if (ty == CUT_INLET) { m.iolet |= bit; m.hadIn = true; }
      else if (ty == CUT_OUTLET) { m.iolet |= bit; m.hadOut = true; }
    }
    const int b = (edge ? 6 : 0) + collision_type(m);
    bucket[t] = (uint8_t)b;
    atomicAdd(&sh[b], 1u);
    if (nRemote) atomicAdd(&sh[12], (unsigned)nRemote);
  }
  __syncthreads();
  if (threadIdx.x < 13 && sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], (unsigned long long)sh[threadIdx.x]);
}

__global__ void iota_kernel(uint32_t* __restrict__ v, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = (uint32_t)i;
}
// local order: gather coordinates (and input indices) and write the final grid codes
__global__ void finalise_sites_kernel(Win W, const uint32_t* __restrict__ travOfLocal, const int32_t* __restrict__ coordsTrav,
                                      const int32_t* __restrict__ inputOfTrav, int64_t N, int32_t* __restrict__ coordsLocal,
                                      int32_t* __restrict__ inputOfLocal, int32_t* __restrict__ grid) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= N) return;
  const int64_t t = travOfLocal[s];
  const int32_t x = coordsTrav[t], y = coordsTrav[N + t], z = coordsTrav[2 * N + t];
  coordsLocal[s] = x;
  coordsLocal[N + s] = y;
  coordsLocal[2 * N + s] = z;
  if (inputOfTrav) inputOfLocal[s] = inputOfTrav[t];
  grid[W.key(x, y, z)] = (int32_t)s;
}

struct RemoteLink { int32_t site, dir, rank, nx, ny, nz; uint32_t trav, pad; };
__global__ void emit_remote_kernel(Win W, LatticeTab L, const int32_t* __restrict__ grid, const int32_t* __restrict__ coordsLocal,
                                   const uint32_t* __restrict__ travOfLocal, int64_t N, int64_t firstEdge,
                                   RemoteLink* __restrict__ out, unsigned long long* __restrict__ counter,
                                   unsigned long long cap) {
  const int64_t s = firstEdge + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= N) return;
  const int64_t x = coordsLocal[s], y = coordsLocal[N + s], z = coordsLocal[2 * N + s];
  const int64_t key = W.key(x, y, z);
  for (int d = 1; d < L.Q; ++d) {
    const int32_t g = grid[key + W.offset(L.c[d][0], L.c[d][1], L.c[d][2])];
    if (g <= -2) {
      const unsigned long long i = atomicAdd(counter, 1ull);
      if (i < cap) {
        RemoteLink r;
        r.site = (int32_t)s;
        r.dir = d;
        r.rank = -2 - g;
        r.nx = (int32_t)(x + L.c[d][0]);
        r.ny = (int32_t)(y + L.c[d][1]);
        r.nz = (int32_t)(z + L.c[d][2]);
        r.trav = travOfLocal[s];
        r.pad = 0;
        out[i] = r;
      }
    }
  }
}

// Domain::neighbourIndices entry in reference form (Domain.cc:425-505, 548-580)
__device__ __forceinline__ int64_t nbr_ref_value(const Win& W, const LatticeTab& L, const int32_t* __restrict__ grid,
                                                 const int32_t* __restrict__ coordsLocal, int64_t N, int64_t s, int d,
                                                 const int64_t* __restrict__ sendKey, const int64_t* __restrict__ sendSlot,
                                                 int64_t S) {
  if (d == 0) return s * L.Q;
  const int64_t key = W.key(coordsLocal[s], coordsLocal[N + s], coordsLocal[2 * N + s]);
  const int32_t g = grid[key + W.offset(L.c[d][0], L.c[d][1], L.c[d][2])];
  if (g >= 0) return (int64_t)g * L.Q + d;
  if (g == -1) return N * L.Q;  // the rubbish site
  const int64_t want = s * L.Q + d;
  int64_t lo = 0, hi = S;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (sendKey[mid] < want) lo = mid + 1; else hi = mid;
  }
  return sendSlot[lo];
}
__global__ void nbr_ref_kernel(Win W, LatticeTab L, const int32_t* __restrict__ grid, const int32_t* __restrict__ coordsLocal,
                               int64_t N, const int64_t* __restrict__ sendKey, const int64_t* __restrict__ sendSlot, int64_t S,
                               int64_t first, int64_t n, int64_t* __restrict__ out) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= n * L.Q) return;
  const int d = (int)(tid / n);
  const int64_t s = first + tid % n;
  out[(s - first) * L.Q + d] = nbr_ref_value(W, L, grid, coordsLocal, N, s, d, sendKey, sendSlot, S);
}
// straight into the engine's pre-renumbering plane layout (what convert_nbr_kernel in abi.cu makes)
__global__ void nbr_engine_kernel(Win W, LatticeTab L, const int32_t* __restrict__ grid, const int32_t* __restrict__ coordsLocal,
                                  int64_t N, const int64_t* __restrict__ sendKey, const int64_t* __restrict__ sendSlot,
                                  int64_t S, int64_t stride, uint32_t* __restrict__ nbr) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= N * (L.Q - 1)) return;
  const int d = (int)(tid / N) + 1;
  const int64_t s = tid % N;
  const int64_t v = nbr_ref_value(W, L, grid, coordsLocal, N, s, d, sendKey, sendSlot, S);
  const int64_t NQ = N * L.Q;
  const int64_t internal = v < NQ ? (v % L.Q) * stride + v / L.Q : v - NQ + (int64_t)L.Q * stride;
  nbr[(int64_t)(d - 1) * stride + s] = (uint32_t)internal;
}
__global__ void coords_planes_kernel(const int32_t* __restrict__ coordsLocal, int64_t N, int64_t stride,
                                     int32_t* __restrict__ out) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= 3 * N) return;
  out[(tid / N) * stride + tid % N] = coordsLocal[tid];
}

// all 26 links + normal of one site
struct FullRecord { uint8_t type[26]; int32_t id[26]; float dist[26]; uint8_t navail; float normal[3]; bool any; };
template <bool ANALYTIC>
__device__ void full_record(const Shape& S, const Explicit& E, const Win& W, const int32_t* __restrict__ grid, int64_t x,
                            int64_t y, int64_t z, int64_t input, FullRecord& R) {
  R.any = false;
  R.navail = 0;
  R.normal[0] = R.normal[1] = R.normal[2] = 0.f;
  if (!ANALYTIC) {
    const int64_t rec = E.recOfInput[input];
    for (int l = 0; l < 26; ++l) {
      R.type[l] = rec >= 0 ? E.type[rec * 26 + l] : 0;
      R.id[l] = rec >= 0 ? E.iolet[rec * 26 + l] : -1;
      R.dist[l] = rec >= 0 ? E.dist[rec * 26 + l] : -1.f;
      if (R.type[l]) R.any = true;
    }
    if (rec >= 0) {
      R.navail = E.navail[rec];
      for (int k = 0; k < 3; ++k) R.normal[k] = E.normal[rec * 3 + k];
      if (R.navail) R.any = true;
    }
    return;
  }
  const int64_t key = W.key(x, y, z);
  int cand[kMaxSiteCand];
  int nCand = -2;
  for (int l = 0; l < 26; ++l) {
    int i, j, k;
    link_vector(l, i, j, k);
    R.type[l] = 0;
    R.id[l] = -1;
    R.dist[l] = -1.f;
    if (grid[key + W.offset(i, j, k)] != -1) continue;
    if (nCand == -2) nCand = site_candidates(S, (double)x, (double)y, (double)z, cand);
    const LinkRes r = analytic_link(S, cand, nCand, (double)x, (double)y, (double)z, i, j, k);
    R.type[l] = (uint8_t)r.type;
    R.id[l] = r.id;
    R.dist[l] = r.dist;
    R.any = true;
    if (r.type == CUT_WALL) R.navail = 1;
  }
  if (R.navail) analytic_normal(S, cand, nCand, (double)x, (double)y, (double)z, R.normal);
}

// boundary tables of the boundary-typed sites (the two contiguous local-id ranges), by ordinal
template <bool ANALYTIC>
__global__ void __launch_bounds__(128) boundary_tables_kernel(Shape S, Explicit E, Win W, LatticeTab L,
                                                             const int32_t* __restrict__ grid,
                                                             const int32_t* __restrict__ coordsLocal,
                                                             const int32_t* __restrict__ inputOfLocal, int64_t N, int64_t NB,
                                                             int64_t midBulk, int64_t midTotal, int64_t edgeBulk,
                                                             uint32_t* __restrict__ bWall, uint32_t* __restrict__ bIolet,
                                                             int32_t* __restrict__ bIoletId, float* __restrict__ bDist,
                                                             float* __restrict__ bNormal) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= NB) return;
  const int64_t nbMid = midTotal - midBulk;
  const int64_t s = b < nbMid ? b + midBulk : b - nbMid + midTotal + edgeBulk;
  FullRecord R;
  full_record<ANALYTIC>(S, E, W, grid, coordsLocal[s], coordsLocal[N + s], coordsLocal[2 * N + s],
                        ANALYTIC ? 0 : inputOfLocal[s], R);
  uint32_t wall = 0, iol = 0;
  int32_t id = -1;
  for (int d = 1; d < L.Q; ++d) {
    const int l = L.link[d];
    const int ty = R.type[l];
    if (ty == CUT_WALL) wall |= 1u << (d - 1);
    else if (ty != CUT_NONE) {
      iol |= 1u << (d - 1);
      id = R.id[l];  // the LAST iolet link in direction order wins (SiteDataBare.cc)
    }
    bDist[(int64_t)(d - 1) * NB + b] = ty != CUT_NONE ? R.dist[l] : -1.f;
  }
  bWall[b] = wall;
  bIolet[b] = iol;
  bIoletId[b] = id;
  for (int k = 0; k < 3; ++k) bNormal[(int64_t)k * NB + b] = R.navail ? R.normal[k] : INFINITY;
}

// GuoZhengShi's remote needs (GuoZhengShi.h:36-104): wall link d of a domain-edge site whose opposite
// direction is neither wall nor iolet and leads to a site of another rank -> {site, opposite
// direction, that rank, the neighbour's coordinates}
struct GzsNeed { int64_t site; int32_t dir, rank; int64_t x, y, z; };
__global__ void gzs_needs_kernel(Win W, LatticeTab L, const int32_t* __restrict__ grid,
                                 const int32_t* __restrict__ coordsLocal, int64_t N, int64_t NB, int64_t nbMid,
                                 int64_t midTotal, int64_t edgeBulk, int me, const uint32_t* __restrict__ bWall,
                                 const uint32_t* __restrict__ bIolet, unsigned long long* __restrict__ counter,
                                 unsigned long long cap, GzsNeed* __restrict__ out) {
  const int64_t b = nbMid + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // domain-edge boundary sites only
  if (b >= NB) return;
  const uint32_t wall = bWall[b], iol = bIolet[b];
  if (!wall) return;
  const int64_t s = b - nbMid + midTotal + edgeBulk;
  const int64_t x = coordsLocal[s], y = coordsLocal[N + s], z = coordsLocal[2 * N + s];
  const int64_t key = W.key(x, y, z);
  for (int d = 1; d < L.Q; ++d) {
    if (!((wall >> (d - 1)) & 1u)) continue;
    const int o = L.inv[d];
    if (((wall >> (o - 1)) & 1u) || ((iol >> (o - 1)) & 1u)) continue;
    const int32_t g = grid[key + W.offset(L.c[o][0], L.c[o][1], L.c[o][2])];
    if (g > -2 || g == -2 - me) continue;
    const unsigned long long k = atomicAdd(counter, 1ull);
    if (k < cap) out[k] = GzsNeed{s, o, -2 - g, x + L.c[o][0], y + L.c[o][1], z + L.c[o][2]};
  }
}
__global__ void lookup_sites_kernel(Win W, const int32_t* __restrict__ grid, int64_t n, const int64_t* __restrict__ coords,
                                    int64_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t x = coords[3 * i], y = coords[3 * i + 1], z = coords[3 * i + 2];
  int64_t v = -1;
  if (W.holds(x, y, z)) {
    const int32_t g = grid[W.key(x, y, z)];
    if (g >= 0) v = g;
  }
  out[i] = v;
}

// geometry download (analytic source): one record per own site with a non-fluid 26-neighbour
__global__ void __launch_bounds__(128) emit_records_kernel(Shape S, Explicit E, Win W, const int32_t* __restrict__ grid,
                                                          const int32_t* __restrict__ coordsLocal,
                                                          const uint32_t* __restrict__ travOfLocal, int64_t N,
                                                          unsigned long long* __restrict__ counter, unsigned long long cap,
                                                          int64_t* __restrict__ recSite, uint8_t* __restrict__ type,
                                                          int32_t* __restrict__ iolet, float* __restrict__ dist,
                                                          uint8_t* __restrict__ navail, float* __restrict__ normal) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= N) return;
  const int64_t x = coordsLocal[s], y = coordsLocal[N + s], z = coordsLocal[2 * N + s];
  const int64_t key = W.key(x, y, z);
  bool cut = false;
  for (int l = 0; l < 26 && !cut; ++l) {
    int i, j, k;
    link_vector(l, i, j, k);
    cut = grid[key + W.offset(i, j, k)] == -1;
  }
  if (!cut) return;
  const unsigned long long r = atomicAdd(counter, 1ull);
  if (r >= cap) return;
  FullRecord R;
  full_record<true>(S, E, W, grid, x, y, z, 0, R);
  recSite[r] = travOfLocal[s];
  for (int l = 0; l < 26; ++l) {
    type[r * 26 + l] = R.type[l];
    iolet[r * 26 + l] = R.id[l];
    dist[r * 26 + l] = R.dist[l];
  }
  navail[r] = R.navail;
  for (int k = 0; k < 3; ++k) normal[r * 3 + k] = R.normal[k];
}

inline unsigned blocks_for(int64_t n, int per = 256) { return (unsigned)((n + per - 1) / per); }

template <typename T> int dmalloc(T*& p, int64_t n) {
  p = nullptr;
  CU(cudaMalloc(&p, sizeof(T) * (size_t)std::max<int64_t>(n, 1)));
  return 0;
}
template <typename T> int upload(T*& p, const T* host, int64_t n) {
  if (dmalloc(p, n)) return 1;
  if (n) CU(cudaMemcpy(p, host, sizeof(T) * (size_t)n, cudaMemcpyHostToDevice));
  return 0;
}

}  // namespace

struct hlb_dom_handle {
  hlb_dom_config cfg;
  LatticeTab L;
  int source = -1;  // 0 explicit, 1 analytic
  int partMode = PART_SINGLE;
  int slabAxis = 2;
  // explicit source (device)
  Explicit E = {};
  std::vector<void*> owned;  // device allocations freed at destroy
  int64_t ownLo[3] = {0, 0, 0}, ownHi[3] = {0, 0, 0};  // voxel bounding box of own sites (explicit), [lo, hi)
  bool anyOwn = false;
  // analytic source
  Capsule* caps = nullptr;
  IoletPlane* iolets = nullptr;
  int nCaps = 0, nIolets = 0;
  std::vector<Capsule> hCaps;
  int64_t* slabFirst = nullptr;
  std::vector<int64_t> hSlabFirst;
  int16_t* rankOfBlock = nullptr;
  std::vector<int32_t> hRankOfBlock;
  // products
  bool built = false;
  Win W = {};
  int32_t* grid = nullptr;
  int64_t N = 0, S = 0, NB = 0;
  int64_t mid[6] = {0, 0, 0, 0, 0, 0}, edge[6] = {0, 0, 0, 0, 0, 0};
  int32_t* coordsLocal = nullptr;   // 3 planes of N
  int32_t* inputOfLocal = nullptr;  // explicit source only
  uint32_t* travOfLocal = nullptr;
  int64_t *sendKey = nullptr, *sendSlot = nullptr;
  std::vector<int> procRank;
  std::vector<int64_t> procCount, procFirst, streaming;
  uint32_t *bWall = nullptr, *bIolet = nullptr;
  int32_t* bIoletId = nullptr;
  float *bDist = nullptr, *bNormal = nullptr;
  // geometry download
  int64_t nRecords = -1;
  int64_t* gRecSite = nullptr;
  uint8_t *gType = nullptr, *gNavail = nullptr;
  int32_t* gIolet = nullptr;
  float *gDist = nullptr, *gNormal = nullptr;
  double buildSeconds = 0;

  Part part() const {
    Part P;
    P.mode = partMode;
    P.me = cfg.rank;
    P.nranks = cfg.nranks;
    P.axis = slabAxis;
    P.B = cfg.block_size;
    for (int k = 0; k < 3; ++k) P.bd[k] = cfg.block_dims[k];
    P.first = slabFirst;
    P.rankOfBlock = rankOfBlock;
    return P;
  }
  double* noise = nullptr;
  int noiseGrid = 0;
  double noiseExtent[3] = {1, 1, 1}, maxRough = 0.0;
  Shape shape() const {
    return Shape{caps, nCaps, iolets, nIolets, noise, noiseGrid, {noiseExtent[0], noiseExtent[1], noiseExtent[2]}, maxRough};
  }
};

namespace {

void free_products(hlb_dom_t d) {
  void* ps[] = {d->grid, d->coordsLocal, d->inputOfLocal, d->travOfLocal, d->sendKey, d->sendSlot, d->bWall, d->bIolet,
                d->bIoletId, d->bDist, d->bNormal, d->gRecSite, d->gType, d->gNavail, d->gIolet, d->gDist, d->gNormal};
  for (void* p : ps) cudaFree(p);
  d->grid = nullptr; d->coordsLocal = nullptr; d->inputOfLocal = nullptr; d->travOfLocal = nullptr;
  d->sendKey = d->sendSlot = nullptr; d->bWall = d->bIolet = nullptr; d->bIoletId = nullptr;
  d->bDist = d->bNormal = nullptr; d->gRecSite = nullptr; d->gType = d->gNavail = nullptr; d->gIolet = nullptr;
  d->gDist = d->gNormal = nullptr;
  d->nRecords = -1;
  d->built = false;
}

// the blocks that can hold own sites: [lo, hi) in block coordinates
int own_block_box(hlb_dom_t d, int64_t lo[3], int64_t hi[3]) {
  const int B = d->cfg.block_size;
  for (int k = 0; k < 3; ++k) { lo[k] = 0; hi[k] = d->cfg.block_dims[k]; }
  if (d->source == 0) {
    if (!d->anyOwn) { for (int k = 0; k < 3; ++k) hi[k] = lo[k] = 0; return 0; }
    for (int k = 0; k < 3; ++k) { lo[k] = d->ownLo[k] / B; hi[k] = (d->ownHi[k] - 1) / B + 1; }
    return 0;
  }
  // analytic: the capsules' bounding box ...
  double clo[3] = {1e300, 1e300, 1e300}, chi[3] = {-1e300, -1e300, -1e300};
  for (const Capsule& c : d->hCaps)
    for (int k = 0; k < 3; ++k) {
      const double e = c.a[k] + c.ab[k];
      clo[k] = std::min(clo[k], std::min(c.a[k], e) - c.r - std::fabs(c.rough) - 1.0);
      chi[k] = std::max(chi[k], std::max(c.a[k], e) + c.r + std::fabs(c.rough) + 1.0);
    }
  for (int k = 0; k < 3; ++k) {
    const int64_t vlo = (int64_t)std::floor(std::max(clo[k], 0.0));
    const int64_t vhi = (int64_t)std::ceil(std::min(chi[k], (double)(d->cfg.block_dims[k] * B - 1))) + 1;
    lo[k] = std::max<int64_t>(lo[k], vlo / B);
    hi[k] = std::min<int64_t>(hi[k], (std::max<int64_t>(vhi, 1) - 1) / B + 1);
    if (hi[k] < lo[k]) hi[k] = lo[k];
  }
  // ... cut down by the site -> rank rule
  if (d->partMode == PART_SLABS) {
    const int a = d->slabAxis, r = d->cfg.rank;
    const int64_t f0 = d->hSlabFirst[r], f1 = d->hSlabFirst[r + 1];
    if (r > 0) lo[a] = std::max<int64_t>(lo[a], std::max<int64_t>(f0, 0) / B);
    if (r + 1 < d->cfg.nranks) hi[a] = std::min<int64_t>(hi[a], (std::max<int64_t>(f1, 1) - 1) / B + 1);
    if (hi[a] < lo[a]) hi[a] = lo[a];
  } else if (d->partMode == PART_BLOCKS) {
    int64_t mlo[3] = {INT64_MAX, INT64_MAX, INT64_MAX}, mhi[3] = {-1, -1, -1};
    const int64_t* bd = d->cfg.block_dims;
    for (int64_t x = lo[0]; x < hi[0]; ++x)
      for (int64_t y = lo[1]; y < hi[1]; ++y)
        for (int64_t z = lo[2]; z < hi[2]; ++z)
          if (d->hRankOfBlock[(x * bd[1] + y) * bd[2] + z] == d->cfg.rank) {
            const int64_t c[3] = {x, y, z};
            for (int k = 0; k < 3; ++k) { mlo[k] = std::min(mlo[k], c[k]); mhi[k] = std::max(mhi[k], c[k] + 1); }
          }
    for (int k = 0; k < 3; ++k) { lo[k] = mhi[k] < 0 ? 0 : mlo[k]; hi[k] = mhi[k] < 0 ? 0 : mhi[k]; }
  }
  return 0;
}

}  // namespace

extern "C" {

int hlb_dom_create(const hlb_dom_config* cfg, hlb_dom_t* out) {
  if (!cfg || !out) return fail("null argument");
  const int Q = cfg->lattice;
  if (Q != 15 && Q != 19 && Q != 27) return fail("lattice must be 15, 19 or 27");
  if (cfg->block_size < 1 || cfg->block_size > 8) return fail("block size must be 1..8 for the device builder");
  if (cfg->nranks < 1 || cfg->rank < 0 || cfg->rank >= cfg->nranks) return fail("bad rank / nranks");
  if (cfg->nranks > 32000) return fail("too many ranks");
  for (int k = 0; k < 3; ++k)
    if (cfg->block_dims[k] < 1 || cfg->block_dims[k] > (1 << 21)) return fail("block_dims out of range");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail("no CUDA device: the Domain builder has no CPU fallback");
  if (cfg->device < 0 || cfg->device >= ndev) return fail("CUDA device ordinal out of range");
  CU(cudaSetDevice(cfg->device));
  hlb_dom_handle* d = new hlb_dom_handle();
  d->cfg = *cfg;
  if (Q == 15) fill_lattice<15>(d->L);
  else if (Q == 19) fill_lattice<19>(d->L);
  else fill_lattice<27>(d->L);
  *out = d;
  return 0;
}

int hlb_dom_destroy(hlb_dom_t d) {
  if (!d) return 0;
  cudaSetDevice(d->cfg.device);
  free_products(d);
  for (void* p : d->owned) cudaFree(p);
  cudaFree(d->caps);
  cudaFree(d->iolets);
  cudaFree(d->slabFirst);
  cudaFree(d->rankOfBlock);
  delete d;
  return 0;
}

int hlb_dom_set_sites(hlb_dom_t d, int64_t n_sites, const int32_t* coords, const int32_t* rank_of_site, int64_t n_records,
                      const int64_t* record_site, const uint8_t* link_type, const int32_t* link_iolet,
                      const float* link_dist, const uint8_t* normal_available, const float* normal) {
  if (!d || (n_sites && !coords)) return fail("null argument");
  if (n_records && (!record_site || !link_type || !link_iolet || !link_dist || !normal_available || !normal))
    return fail("null argument");
  if (d->source >= 0) return fail("site source already set");
  if (n_sites >= ((int64_t)1 << 31)) return fail("too many sites for one rank's builder");
  CU(cudaSetDevice(d->cfg.device));
  const int B = d->cfg.block_size;
  // bounding box of own sites; reject coordinates outside the block lattice
  int64_t lo[3] = {INT64_MAX, INT64_MAX, INT64_MAX}, hi[3] = {INT64_MIN, INT64_MIN, INT64_MIN};
  for (int64_t i = 0; i < n_sites; ++i) {
    for (int k = 0; k < 3; ++k) {
      const int64_t v = coords[3 * i + k];
      if (v < 0 || v >= d->cfg.block_dims[k] * B) return fail("site coordinates outside the block lattice");
    }
    const int r = rank_of_site ? rank_of_site[i] : 0;
    if (r < 0 || r >= d->cfg.nranks) return fail("site rank outside the communicator");
    if (r != d->cfg.rank) continue;
    d->anyOwn = true;
    for (int k = 0; k < 3; ++k) {
      lo[k] = std::min<int64_t>(lo[k], coords[3 * i + k]);
      hi[k] = std::max<int64_t>(hi[k], coords[3 * i + k] + 1);
    }
  }
  for (int k = 0; k < 3; ++k) { d->ownLo[k] = d->anyOwn ? lo[k] : 0; d->ownHi[k] = d->anyOwn ? hi[k] : 0; }
  std::vector<int32_t> recOf(std::max<int64_t>(n_sites, 1), -1);
  for (int64_t r = 0; r < n_records; ++r) {
    if (record_site[r] < 0 || record_site[r] >= n_sites) return fail("cut-link record names a site outside the list");
    recOf[record_site[r]] = (int32_t)r;
  }
  int32_t *dc = nullptr, *dr = nullptr, *drec = nullptr, *dio = nullptr;
  uint8_t *dty = nullptr, *dna = nullptr;
  float *ddi = nullptr, *dno = nullptr;
  if (upload(dc, coords, 3 * n_sites)) return 1;
  d->owned.push_back(dc);
  if (rank_of_site) {
    if (upload(dr, rank_of_site, n_sites)) return 1;
    d->owned.push_back(dr);
  }
  if (upload(drec, recOf.data(), n_sites)) return 1;
  d->owned.push_back(drec);
  if (upload(dty, link_type, 26 * n_records)) return 1;
  d->owned.push_back(dty);
  if (upload(dio, link_iolet, 26 * n_records)) return 1;
  d->owned.push_back(dio);
  if (upload(ddi, link_dist, 26 * n_records)) return 1;
  d->owned.push_back(ddi);
  if (upload(dna, normal_available, n_records)) return 1;
  d->owned.push_back(dna);
  if (upload(dno, normal, 3 * n_records)) return 1;
  d->owned.push_back(dno);
  d->E = Explicit{dc, dr, drec, dty, dio, ddi, dna, dno, n_sites};
  d->source = 0;
  d->partMode = PART_EXPLICIT;
  return 0;
}

int hlb_dom_set_shape(hlb_dom_t d, int n_capsules, const double* capsules, int n_iolets, const double* iolets) {
  if (!d || !capsules || n_capsules < 1 || (n_iolets && !iolets)) return fail("bad argument");
  if (d->source >= 0) return fail("site source already set");
  CU(cudaSetDevice(d->cfg.device));
  d->hCaps.resize(n_capsules);
  for (int k = 0; k < n_capsules; ++k) {
    const double* c = capsules + 7 * k;
    Capsule& o = d->hCaps[k];
    o.L2 = 0.0;
    for (int j = 0; j < 3; ++j) {
      o.a[j] = c[j];
      o.ab[j] = c[3 + j] - c[j];
      o.L2 += o.ab[j] * o.ab[j];
    }
    o.r = c[6];
    o.rough = 0.0;
    if (!(o.r > 0)) return fail("capsule radius must be positive");
  }
  std::vector<IoletPlane> io(std::max(n_iolets, 1));
  for (int k = 0; k < n_iolets; ++k) {
    const double* c = iolets + 9 * k;
    io[k].kind = (int)c[0];
    io[k].index = (int)c[1];
    if (io[k].kind != CUT_INLET && io[k].kind != CUT_OUTLET) return fail("iolet kind must be 2 (inlet) or 3 (outlet)");
    for (int j = 0; j < 3; ++j) { io[k].pos[j] = c[2 + j]; io[k].n[j] = c[5 + j]; }
    io[k].radius = c[8];
  }
  if (upload(d->caps, d->hCaps.data(), n_capsules)) return 1;
  if (upload(d->iolets, io.data(), n_iolets)) return 1;
  d->nCaps = n_capsules;
  d->nIolets = n_iolets;
  d->source = 1;
  return 0;
}

int hlb_dom_set_roughness(hlb_dom_t d, const double* amplitude, int grid, const double* noise, const double* extent) {
  if (!d || !amplitude || !noise || !extent || grid < 2) return fail("bad argument");
  if (d->source != 1) return fail("hlb_dom_set_roughness needs hlb_dom_set_shape first");
  CU(cudaSetDevice(d->cfg.device));
  d->maxRough = 0.0;
  for (int k = 0; k < d->nCaps; ++k) {
    d->hCaps[k].rough = amplitude[k];
    d->maxRough = std::max(d->maxRough, std::fabs(amplitude[k]));
  }
  if (upload(d->caps, d->hCaps.data(), d->nCaps)) return 1;
  if (upload(d->noise, noise, grid * grid * grid)) return 1;
  d->noiseGrid = grid;
  for (int k = 0; k < 3; ++k) {
    if (!(extent[k] > 0)) return fail("noise extent must be positive");
    d->noiseExtent[k] = extent[k];
  }
  return 0;
}

int hlb_dom_set_partition_slabs(hlb_dom_t d, int axis, const int64_t* first_coord) {
  if (!d || !first_coord || axis < 0 || axis > 2) return fail("bad argument");
  if (d->source == 0) return fail("an explicit site list carries its own site -> rank map");
  CU(cudaSetDevice(d->cfg.device));
  d->hSlabFirst.assign(first_coord, first_coord + d->cfg.nranks + 1);
  for (int r = 0; r < d->cfg.nranks; ++r)
    if (d->hSlabFirst[r] > d->hSlabFirst[r + 1]) return fail("slab boundaries must ascend");
  cudaFree(d->slabFirst);
  if (upload(d->slabFirst, d->hSlabFirst.data(), d->cfg.nranks + 1)) return 1;
  d->partMode = PART_SLABS;
  d->slabAxis = axis;
  return 0;
}

int hlb_dom_set_partition_blocks(hlb_dom_t d, const int32_t* rank_of_block) {
  if (!d || !rank_of_block) return fail("null argument");
  if (d->source == 0) return fail("an explicit site list carries its own site -> rank map");
  CU(cudaSetDevice(d->cfg.device));
  const int64_t nb = d->cfg.block_dims[0] * d->cfg.block_dims[1] * d->cfg.block_dims[2];
  d->hRankOfBlock.assign(rank_of_block, rank_of_block + nb);
  std::vector<int16_t> v(nb);
  for (int64_t i = 0; i < nb; ++i) {
    if (rank_of_block[i] < -1 || rank_of_block[i] >= d->cfg.nranks) return fail("block rank outside the communicator");
    v[i] = (int16_t)(rank_of_block[i] < 0 ? 0 : rank_of_block[i]);
  }
  cudaFree(d->rankOfBlock);
  if (upload(d->rankOfBlock, v.data(), nb)) return 1;
  d->partMode = PART_BLOCKS;
  return 0;
}

static int count_blocks(hlb_dom_t d, const int64_t* lo, const int64_t* hi, int32_t* counts, int32_t* boundary);

int hlb_dom_count_block_sites(hlb_dom_t d, const int64_t* lo, const int64_t* hi, int32_t* counts) {
  return count_blocks(d, lo, hi, counts, nullptr);
}

int hlb_dom_count_block_sites_typed(hlb_dom_t d, const int64_t* lo, const int64_t* hi, int32_t* counts,
                                    int32_t* boundary_counts) {
  if (!boundary_counts) return fail("null argument");
  return count_blocks(d, lo, hi, counts, boundary_counts);
}

static int count_blocks(hlb_dom_t d, const int64_t* lo, const int64_t* hi, int32_t* counts, int32_t* boundary) {
  if (!d || !lo || !hi || !counts) return fail("null argument");
  if (d->source != 1) return fail("block counting needs the analytic shape source");
  CU(cudaSetDevice(d->cfg.device));
  Box3 box;
  int64_t nb = 1;
  for (int k = 0; k < 3; ++k) {
    if (lo[k] < 0 || hi[k] > d->cfg.block_dims[k] || hi[k] < lo[k]) return fail("block box outside the lattice");
    box.lo[k] = lo[k];
    box.dim[k] = hi[k] - lo[k];
    nb *= box.dim[k];
  }
  if (nb == 0) return 0;
  if (nb >= ((int64_t)1 << 31)) return fail("block box too large");
  int32_t* dcnt = nullptr;
  if (dmalloc(dcnt, nb)) return 1;
  Win W = {};
  int32_t* dbnd = nullptr;
  LatticeTab* dlat = nullptr;
  if (boundary) {
    if (dmalloc(dbnd, nb)) return 1;
    if (upload(dlat, &d->L, 1)) return 1;
  }
  classify_blocks_kernel<true><<<(unsigned)nb, 512>>>(d->shape(), d->part(), W, box, nullptr, dcnt, dbnd, dlat);
  CU(cudaGetLastError());
  CU(cudaMemcpy(counts, dcnt, sizeof(int32_t) * nb, cudaMemcpyDeviceToHost));
  if (boundary) CU(cudaMemcpy(boundary, dbnd, sizeof(int32_t) * nb, cudaMemcpyDeviceToHost));
  cudaFree(dcnt);
  if (dbnd) cudaFree(dbnd);
  if (dlat) cudaFree(dlat);
  return 0;
}

int hlb_dom_build(hlb_dom_t d) {
  if (!d) return fail("null argument");
  if (d->source < 0) return fail("no site source: call hlb_dom_set_sites or hlb_dom_set_shape first");
  if (d->source == 1 && d->cfg.nranks > 1 && d->partMode == PART_SINGLE)
    return fail("several ranks but no site -> rank rule: call hlb_dom_set_partition_slabs / _blocks");
  CU(cudaSetDevice(d->cfg.device));
  free_products(d);
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  CU(cudaEventRecord(e0));
  const int B = d->cfg.block_size, Q = d->cfg.lattice;
  const bool analytic = d->source == 1;
  const Part P = d->part();
  const Shape S = d->shape();
  const Explicit E = d->E;
  const LatticeTab L = d->L;

  // ---- window: own blocks, one voxel of rim; block box one block wider
  int64_t blo[3], bhi[3];
  if (own_block_box(d, blo, bhi)) return 1;
  Win W;
  Box3 box;
  int64_t nvox = 1, nblocks = 1;
  for (int k = 0; k < 3; ++k) {
    W.org[k] = blo[k] * B - 1;
    W.dim[k] = (bhi[k] - blo[k]) * B + 2;
    box.lo[k] = blo[k] - 1;
    box.dim[k] = bhi[k] - blo[k] + 2;
    nvox *= W.dim[k];
    nblocks *= box.dim[k];
  }
  d->W = W;
  if (nvox > (int64_t)24e9) return fail("voxel window of this rank exceeds 24e9 voxels: use more ranks or a tighter partition");
  if (nblocks >= ((int64_t)1 << 31)) return fail("too many blocks in this rank's window");
  if (dmalloc(d->grid, nvox)) return 1;
  CU(cudaMemset(d->grid, 0xff, sizeof(int32_t) * (size_t)nvox));
  int32_t* blockCount = nullptr;
  if (dmalloc(blockCount, nblocks)) return 1;
  CU(cudaMemset(blockCount, 0, sizeof(int32_t) * (size_t)nblocks));
  if (analytic) {
    classify_blocks_kernel<false><<<(unsigned)nblocks, 512>>>(S, P, W, box, d->grid, blockCount);
    CU(cudaGetLastError());
  } else {
    int* bad = nullptr;
    if (dmalloc(bad, 1)) return 1;
    CU(cudaMemset(bad, 0, sizeof(int)));
    if (E.n) mark_sites_kernel<<<blocks_for(E.n), 256>>>(E, P, W, box, d->grid, blockCount, bad);
    CU(cudaGetLastError());
    int hb = 0;
    CU(cudaMemcpy(&hb, bad, sizeof(int), cudaMemcpyDeviceToHost));
    cudaFree(bad);
    if (hb == 2) return fail("two sites share one voxel");
    if (hb) return fail("internal error: own site outside the window");
  }

  // ---- traversal order: blocks by Morton index (LookupTree.cc:155), sites z-fastest inside
  uint64_t *keysIn = nullptr, *keysOut = nullptr;
  uint32_t *valsIn = nullptr, *sortedBlocks = nullptr;
  int64_t *cnt64 = nullptr, *scanned = nullptr, *blockStart = nullptr;
  if (dmalloc(keysIn, nblocks) || dmalloc(keysOut, nblocks) || dmalloc(valsIn, nblocks) || dmalloc(sortedBlocks, nblocks) ||
      dmalloc(cnt64, nblocks + 1) || dmalloc(scanned, nblocks + 1) || dmalloc(blockStart, nblocks))
    return 1;
  block_keys_kernel<<<blocks_for(nblocks), 256>>>(box, P, blockCount, keysIn, valsIn, nblocks);
  CU(cudaGetLastError());
  {
    void* tmp = nullptr;
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(tmp, tb, keysIn, keysOut, valsIn, sortedBlocks, (int)nblocks);
    CU(cudaMalloc(&tmp, tb + 16));
    cub::DeviceRadixSort::SortPairs(tmp, tb, keysIn, keysOut, valsIn, sortedBlocks, (int)nblocks);
    CU(cudaGetLastError());
    CU(cudaDeviceSynchronize());
    cudaFree(tmp);
  }
  gather_counts_kernel<<<blocks_for(nblocks), 256>>>(sortedBlocks, blockCount, cnt64, nblocks);
  CU(cudaMemset(cnt64 + nblocks, 0, sizeof(int64_t)));
  {
    void* tmp = nullptr;
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(tmp, tb, cnt64, scanned, (int)(nblocks + 1));
    CU(cudaMalloc(&tmp, tb + 16));
    cub::DeviceScan::ExclusiveSum(tmp, tb, cnt64, scanned, (int)(nblocks + 1));
    CU(cudaGetLastError());
    CU(cudaDeviceSynchronize());
    cudaFree(tmp);
  }
  int64_t N = 0;
  CU(cudaMemcpy(&N, scanned + nblocks, sizeof(int64_t), cudaMemcpyDeviceToHost));
  scatter_starts_kernel<<<blocks_for(nblocks), 256>>>(sortedBlocks, scanned, blockStart, nblocks);
  CU(cudaGetLastError());
  cudaFree(keysIn); cudaFree(keysOut); cudaFree(valsIn); cudaFree(sortedBlocks); cudaFree(cnt64); cudaFree(scanned);
  d->N = N;
  if (N >= ((int64_t)1 << 31)) return fail("too many sites on one rank");

  int32_t *coordsTrav = nullptr, *inputOfTrav = nullptr;
  if (dmalloc(coordsTrav, 3 * N)) return 1;
  if (!analytic && dmalloc(inputOfTrav, N)) return 1;
  if (analytic) compact_kernel<true><<<(unsigned)nblocks, 512>>>(E, P, W, box, d->grid, blockStart, N, coordsTrav, inputOfTrav);
  else compact_kernel<false><<<(unsigned)nblocks, 512>>>(E, P, W, box, d->grid, blockStart, N, coordsTrav, inputOfTrav);
  CU(cudaGetLastError());
  CU(cudaDeviceSynchronize());
  cudaFree(blockStart);
  cudaFree(blockCount);

  // ---- collision-type buckets (Domain.cc:287-357): mid-domain[0..5] then domain-edge[0..5]
  uint8_t *bucket = nullptr, *bucketOut = nullptr;
  unsigned long long* hist = nullptr;
  uint32_t* iota = nullptr;
  if (dmalloc(bucket, N) || dmalloc(bucketOut, N) || dmalloc(hist, 16) || dmalloc(iota, N) || dmalloc(d->travOfLocal, N)) return 1;
  CU(cudaMemset(hist, 0, sizeof(unsigned long long) * 16));
  if (N) {
    if (analytic)
      classify_sites_kernel<true><<<blocks_for(N), 256>>>(S, E, P, W, L, d->grid, coordsTrav, inputOfTrav, N, bucket, hist);
    else
      classify_sites_kernel<false><<<blocks_for(N), 256>>>(S, E, P, W, L, d->grid, coordsTrav, inputOfTrav, N, bucket, hist);
    CU(cudaGetLastError());
    iota_kernel<<<blocks_for(N), 256>>>(iota, N);
    void* tmp = nullptr;
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(tmp, tb, bucket, bucketOut, iota, d->travOfLocal, (int)N, 0, 4);
    CU(cudaMalloc(&tmp, tb + 16));
    cub::DeviceRadixSort::SortPairs(tmp, tb, bucket, bucketOut, iota, d->travOfLocal, (int)N, 0, 4);
    CU(cudaGetLastError());
    CU(cudaDeviceSynchronize());
    cudaFree(tmp);
  }
  unsigned long long hh[16];
  CU(cudaMemcpy(hh, hist, sizeof(hh), cudaMemcpyDeviceToHost));
  cudaFree(bucket); cudaFree(bucketOut); cudaFree(hist); cudaFree(iota);
  int64_t midTotal = 0;
  for (int t = 0; t < 6; ++t) { d->mid[t] = (int64_t)hh[t]; d->edge[t] = (int64_t)hh[6 + t]; midTotal += d->mid[t]; }
  const int64_t nRemote = (int64_t)hh[12];
  d->NB = N - d->mid[0] - d->edge[0];

  if (dmalloc(d->coordsLocal, 3 * N)) return 1;
  if (!analytic && dmalloc(d->inputOfLocal, N)) return 1;
  if (N) {
    finalise_sites_kernel<<<blocks_for(N), 256>>>(W, d->travOfLocal, coordsTrav, inputOfTrav, N, d->coordsLocal,
                                                  d->inputOfLocal, d->grid);
    CU(cudaGetLastError());
    CU(cudaDeviceSynchronize());
  }
  cudaFree(coordsTrav);
  cudaFree(inputOfTrav);

  // ---- halo tables (Domain.cc:247-285, 404-419, 507-580) from the compact list of remote links
  d->procRank.clear(); d->procCount.clear(); d->procFirst.clear(); d->streaming.clear();
  std::vector<int64_t> keys, slots;
  d->S = nRemote;
  if (nRemote) {
    RemoteLink* drem = nullptr;
    unsigned long long* dcounter = nullptr;
    if (dmalloc(drem, nRemote) || dmalloc(dcounter, 1)) return 1;
    CU(cudaMemset(dcounter, 0, sizeof(unsigned long long)));
    const int64_t nEdge = N - midTotal;
    emit_remote_kernel<<<blocks_for(nEdge), 256>>>(W, L, d->grid, d->coordsLocal, d->travOfLocal, N, midTotal, drem, dcounter,
                                                   (unsigned long long)nRemote);
    CU(cudaGetLastError());
    unsigned long long got = 0;
    CU(cudaMemcpy(&got, dcounter, sizeof(got), cudaMemcpyDeviceToHost));
    if ((int64_t)got != nRemote) return fail("internal error: remote link count changed between passes");
    std::vector<RemoteLink> rem(nRemote);
    CU(cudaMemcpy(rem.data(), drem, sizeof(RemoteLink) * nRemote, cudaMemcpyDeviceToHost));
    cudaFree(drem);
    cudaFree(dcounter);
    // this rank's own list order: traversal order of the site, then direction
    std::sort(rem.begin(), rem.end(), [](const RemoteLink& a, const RemoteLink& b) {
      return a.trav != b.trav ? a.trav < b.trav : a.dir < b.dir;
    });
    // neighbouringProcs: first-encounter order
    std::vector<int> slotOfRank(d->cfg.nranks, -1);
    for (const RemoteLink& r : rem) {
      if (r.rank < 0 || r.rank >= d->cfg.nranks) return fail("internal error: remote link to a rank outside the communicator");
      if (slotOfRank[r.rank] < 0) {
        slotOfRank[r.rank] = (int)d->procRank.size();
        d->procRank.push_back(r.rank);
        d->procCount.push_back(0);
      }
      d->procCount[slotOfRank[r.rank]]++;
    }
    int64_t fcount = N * Q;
    keys.reserve(nRemote);
    slots.reserve(nRemote);
    d->streaming.reserve(nRemote);
    std::vector<const RemoteLink*> mine;
    for (size_t pi = 0; pi < d->procRank.size(); ++pi) {
      const int p = d->procRank[pi];
      d->procFirst.push_back(fcount + 1);
      mine.clear();
      for (const RemoteLink& r : rem)
        if (r.rank == p) mine.push_back(&r);
      if (p < d->cfg.rank) {
        // the lower rank's list is authoritative (Domain.cc:530-542): its traversal order of ITS
        // site (our neighbour), then ITS direction (the inverse of ours)
        struct K { uint64_t m; int32_t in, dir; const RemoteLink* r; };
        std::vector<K> ks(mine.size());
        for (size_t i = 0; i < mine.size(); ++i) {
          const RemoteLink* r = mine[i];
          ks[i].m = morton3(r->nx / B, r->ny / B, r->nz / B);
          ks[i].in = ((r->nx % B) * B + (r->ny % B)) * B + (r->nz % B);
          ks[i].dir = L.inv[r->dir];
          ks[i].r = r;
        }
        std::stable_sort(ks.begin(), ks.end(), [](const K& a, const K& b) {
          if (a.m != b.m) return a.m < b.m;
          if (a.in != b.in) return a.in < b.in;
          return a.dir < b.dir;
        });
        for (size_t i = 0; i < mine.size(); ++i) mine[i] = ks[i].r;
      }
      for (const RemoteLink* r : mine) {
        ++fcount;
        keys.push_back((int64_t)r->site * Q + r->dir);
        slots.push_back(fcount);
        d->streaming.push_back((int64_t)r->site * Q + L.inv[r->dir]);
      }
    }
    // sorted (site*Q + dir) -> slot for the device-side table fill
    std::vector<int64_t> order(nRemote);
    for (int64_t i = 0; i < nRemote; ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return keys[a] < keys[b]; });
    std::vector<int64_t> k2(nRemote), s2(nRemote);
    for (int64_t i = 0; i < nRemote; ++i) { k2[i] = keys[order[i]]; s2[i] = slots[order[i]]; }
    if (upload(d->sendKey, k2.data(), nRemote) || upload(d->sendSlot, s2.data(), nRemote)) return 1;
  } else {
    if (dmalloc(d->sendKey, 1) || dmalloc(d->sendSlot, 1)) return 1;
  }

  // ---- boundary tables of the boundary-typed sites
  const int64_t NB = d->NB;
  if (dmalloc(d->bWall, NB) || dmalloc(d->bIolet, NB) || dmalloc(d->bIoletId, NB) || dmalloc(d->bDist, NB * (Q - 1)) ||
      dmalloc(d->bNormal, NB * 3))
    return 1;
  if (NB) {
    if (analytic)
      boundary_tables_kernel<true><<<blocks_for(NB, 128), 128>>>(S, E, W, L, d->grid, d->coordsLocal, d->inputOfLocal, N, NB,
                                                                  d->mid[0], midTotal, d->edge[0], d->bWall, d->bIolet,
                                                                  d->bIoletId, d->bDist, d->bNormal);
    else
      boundary_tables_kernel<false><<<blocks_for(NB, 128), 128>>>(S, E, W, L, d->grid, d->coordsLocal, d->inputOfLocal, N, NB,
                                                                   d->mid[0], midTotal, d->edge[0], d->bWall, d->bIolet,
                                                                   d->bIoletId, d->bDist, d->bNormal);
    CU(cudaGetLastError());
  }
  CU(cudaEventRecord(e1));
  CU(cudaEventSynchronize(e1));
  float ms = 0.f;
  CU(cudaEventElapsedTime(&ms, e0, e1));
  d->buildSeconds = ms * 1e-3;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  d->built = true;
  return 0;
}

int hlb_dom_get_counts(hlb_dom_t d, int64_t* n_sites, int64_t* mid6, int64_t* edge6, int64_t* total_shared_fs,
                       int* n_neighbours) {
  if (!d || !n_sites || !mid6 || !edge6 || !total_shared_fs || !n_neighbours) return fail("null argument");
  if (!d->built) return fail("domain not built");
  *n_sites = d->N;
  for (int t = 0; t < 6; ++t) { mid6[t] = d->mid[t]; edge6[t] = d->edge[t]; }
  *total_shared_fs = d->S;
  *n_neighbours = (int)d->procRank.size();
  return 0;
}

int hlb_dom_get_neighbours(hlb_dom_t d, int* rank, int64_t* count, int64_t* first) {
  if (!d) return fail("null argument");
  if (!d->built) return fail("domain not built");
  for (size_t i = 0; i < d->procRank.size(); ++i) {
    rank[i] = d->procRank[i];
    count[i] = d->procCount[i];
    first[i] = d->procFirst[i];
  }
  return 0;
}

int hlb_dom_get_streaming_indices(hlb_dom_t d, int64_t* idx) {
  if (!d) return fail("null argument");
  if (!d->built) return fail("domain not built");
  for (size_t i = 0; i < d->streaming.size(); ++i) idx[i] = d->streaming[i];
  return 0;
}

int hlb_dom_get_neighbour_indices(hlb_dom_t d, int64_t first, int64_t n, int64_t* idx) {
  if (!d || !idx) return fail("null argument");
  if (!d->built) return fail("domain not built");
  if (first < 0 || n < 0 || first + n > d->N) return fail("site range outside the local fluid sites");
  CU(cudaSetDevice(d->cfg.device));
  const int Q = d->cfg.lattice;
  const int64_t chunk = 1 << 20;
  int64_t* buf = nullptr;
  if (dmalloc(buf, chunk * Q)) return 1;
  for (int64_t s0 = 0; s0 < n; s0 += chunk) {
    const int64_t m = std::min(chunk, n - s0);
    nbr_ref_kernel<<<blocks_for(m * Q), 256>>>(d->W, d->L, d->grid, d->coordsLocal, d->N, d->sendKey, d->sendSlot, d->S,
                                               first + s0, m, buf);
    CU(cudaGetLastError());
    CU(cudaMemcpy(idx + s0 * Q, buf, sizeof(int64_t) * m * Q, cudaMemcpyDeviceToHost));
  }
  cudaFree(buf);
  return 0;
}

}  // extern "C"
int hlb_dom_internal_coords(hlb_dom_handle* d, const int32_t** planes, int64_t* n_sites, int* device) {
  if (!d || !planes || !n_sites) return fail("null argument");
  if (!d->built) return fail("domain not built");
  *planes = d->coordsLocal;
  *n_sites = d->N;
  if (device) *device = d->cfg.device;
  return 0;
}
extern "C" {

int hlb_dom_get_site_coords(hlb_dom_t d, int64_t first, int64_t n, int64_t* coords) {
  if (!d || !coords) return fail("null argument");
  if (!d->built) return fail("domain not built");
  if (first < 0 || n < 0 || first + n > d->N) return fail("site range outside the local fluid sites");
  CU(cudaSetDevice(d->cfg.device));
  std::vector<int32_t> tmp(std::max<int64_t>(n, 1));
  for (int k = 0; k < 3; ++k) {
    if (n) CU(cudaMemcpy(tmp.data(), d->coordsLocal + (int64_t)k * d->N + first, sizeof(int32_t) * n, cudaMemcpyDeviceToHost));
    for (int64_t i = 0; i < n; ++i) coords[3 * i + k] = tmp[i];
  }
  return 0;
}

int hlb_dom_get_input_index(hlb_dom_t d, int64_t first, int64_t n, int64_t* idx) {
  if (!d || !idx) return fail("null argument");
  if (!d->built) return fail("domain not built");
  if (!d->inputOfLocal) return fail("input indices exist only for an explicit site list");
  if (first < 0 || n < 0 || first + n > d->N) return fail("site range outside the local fluid sites");
  CU(cudaSetDevice(d->cfg.device));
  std::vector<int32_t> tmp(std::max<int64_t>(n, 1));
  if (n) CU(cudaMemcpy(tmp.data(), d->inputOfLocal + first, sizeof(int32_t) * n, cudaMemcpyDeviceToHost));
  for (int64_t i = 0; i < n; ++i) idx[i] = tmp[i];
  return 0;
}

int hlb_dom_get_boundary_tables(hlb_dom_t d, uint32_t* wall, uint32_t* iolet, int32_t* iolet_id, double* dist,
                                double* normal) {
  if (!d || !wall || !iolet || !iolet_id || !dist || !normal) return fail("null argument");
  if (!d->built) return fail("domain not built");
  CU(cudaSetDevice(d->cfg.device));
  const int64_t NB = d->NB;
  const int Q = d->cfg.lattice;
  if (!NB) return 0;
  CU(cudaMemcpy(wall, d->bWall, sizeof(uint32_t) * NB, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(iolet, d->bIolet, sizeof(uint32_t) * NB, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(iolet_id, d->bIoletId, sizeof(int32_t) * NB, cudaMemcpyDeviceToHost));
  std::vector<float> t((size_t)NB * (Q - 1));
  CU(cudaMemcpy(t.data(), d->bDist, sizeof(float) * t.size(), cudaMemcpyDeviceToHost));
  for (int64_t b = 0; b < NB; ++b)
    for (int k = 0; k < Q - 1; ++k) dist[b * (Q - 1) + k] = (double)t[(size_t)k * NB + b];
  t.resize((size_t)NB * 3);
  CU(cudaMemcpy(t.data(), d->bNormal, sizeof(float) * t.size(), cudaMemcpyDeviceToHost));
  for (int64_t b = 0; b < NB; ++b)
    for (int k = 0; k < 3; ++k) normal[b * 3 + k] = (double)t[(size_t)k * NB + b];
  return 0;
}

int hlb_dom_gzs_needs(hlb_dom_t d, int64_t capacity, int64_t* n, int64_t* local_site, int32_t* direction,
                      int32_t* owner_rank, int64_t* coords) {
  if (!d || !n) return fail("null argument");
  if (!d->built) return fail("domain not built");
  CU(cudaSetDevice(d->cfg.device));
  *n = 0;
  const int64_t midTotal = d->mid[0] + d->mid[1] + d->mid[2] + d->mid[3] + d->mid[4] + d->mid[5];
  const int64_t nbMid = midTotal - d->mid[0];
  const int64_t nEdgeB = d->NB - nbMid;
  if (nEdgeB <= 0 || d->cfg.nranks <= 1) return 0;
  unsigned long long* counter = nullptr;
  GzsNeed* out = nullptr;
  CU(cudaMalloc(&counter, sizeof(unsigned long long)));
  CU(cudaMemset(counter, 0, sizeof(unsigned long long)));
  if (capacity > 0) CU(cudaMalloc(&out, sizeof(GzsNeed) * capacity));
  gzs_needs_kernel<<<blocks_for(nEdgeB), 256>>>(d->W, d->L, d->grid, d->coordsLocal, d->N, d->NB, nbMid, midTotal, d->edge[0],
                                                d->cfg.rank, d->bWall, d->bIolet, counter, (unsigned long long)capacity, out);
  CU(cudaGetLastError());
  unsigned long long total = 0;
  CU(cudaMemcpy(&total, counter, sizeof(total), cudaMemcpyDeviceToHost));
  cudaFree(counter);
  *n = (int64_t)total;
  if (capacity > 0 && (int64_t)total <= capacity && total) {
    if (!local_site || !direction || !owner_rank || !coords) { cudaFree(out); return fail("null argument"); }
    std::vector<GzsNeed> h(total);
    CU(cudaMemcpy(h.data(), out, sizeof(GzsNeed) * total, cudaMemcpyDeviceToHost));
    // an order both sides of a pair can rely on: owner rank, then requesting site, then direction
    std::sort(h.begin(), h.end(), [](const GzsNeed& a, const GzsNeed& b) {
      if (a.rank != b.rank) return a.rank < b.rank;
      if (a.site != b.site) return a.site < b.site;
      return a.dir < b.dir;
    });
    for (size_t k = 0; k < h.size(); ++k) {
      local_site[k] = h[k].site;
      direction[k] = h[k].dir;
      owner_rank[k] = h[k].rank;
      coords[3 * k] = h[k].x;
      coords[3 * k + 1] = h[k].y;
      coords[3 * k + 2] = h[k].z;
    }
  }
  cudaFree(out);
  return 0;
}

int hlb_dom_lookup_sites(hlb_dom_t d, int64_t n, const int64_t* coords, int64_t* local_site) {
  if (!d || (n && (!coords || !local_site))) return fail("null argument");
  if (!d->built) return fail("domain not built");
  if (n <= 0) return 0;
  CU(cudaSetDevice(d->cfg.device));
  int64_t *dc = nullptr, *ds = nullptr;
  CU(cudaMalloc(&dc, sizeof(int64_t) * 3 * n));
  CU(cudaMalloc(&ds, sizeof(int64_t) * n));
  CU(cudaMemcpy(dc, coords, sizeof(int64_t) * 3 * n, cudaMemcpyHostToDevice));
  lookup_sites_kernel<<<blocks_for(n), 256>>>(d->W, d->grid, n, dc, ds);
  CU(cudaGetLastError());
  CU(cudaMemcpy(local_site, ds, sizeof(int64_t) * n, cudaMemcpyDeviceToHost));
  cudaFree(dc);
  cudaFree(ds);
  return 0;
}

int hlb_dom_get_geometry_sizes(hlb_dom_t d, int64_t* n_sites, int64_t* n_records) {
  if (!d || !n_sites || !n_records) return fail("null argument");
  if (!d->built) return fail("domain not built");
  if (d->source != 1) return fail("geometry download is for the analytic shape source");
  CU(cudaSetDevice(d->cfg.device));
  if (d->nRecords < 0) {
    const int64_t N = d->N;
    // upper bound on records: every boundary-typed site plus bulk-typed sites with only
    // off-lattice cuts; count first with a zero-capacity pass
    unsigned long long* counter = nullptr;
    if (dmalloc(counter, 1)) return 1;
    int64_t cap = 0;
    for (int pass = 0; pass < 2; ++pass) {
      CU(cudaMemset(counter, 0, sizeof(unsigned long long)));
      if (pass == 1) {
        if (dmalloc(d->gRecSite, cap) || dmalloc(d->gType, cap * 26) || dmalloc(d->gIolet, cap * 26) ||
            dmalloc(d->gDist, cap * 26) || dmalloc(d->gNavail, cap) || dmalloc(d->gNormal, cap * 3))
          return 1;
      }
      if (N)
        emit_records_kernel<<<blocks_for(N, 128), 128>>>(d->shape(), d->E, d->W, d->grid, d->coordsLocal, d->travOfLocal, N,
                                                         counter, (unsigned long long)cap, d->gRecSite, d->gType, d->gIolet,
                                                         d->gDist, d->gNavail, d->gNormal);
      CU(cudaGetLastError());
      unsigned long long got = 0;
      CU(cudaMemcpy(&got, counter, sizeof(got), cudaMemcpyDeviceToHost));
      cap = (int64_t)got;
    }
    cudaFree(counter);
    d->nRecords = cap;
  }
  *n_sites = d->N;
  *n_records = d->nRecords;
  return 0;
}

int hlb_dom_get_geometry(hlb_dom_t d, int32_t* coords, int64_t* record_site, uint8_t* type, int32_t* iolet, float* dist,
                         uint8_t* navail, float* normal) {
  if (!d || !coords) return fail("null argument");
  int64_t n = 0, nr = 0;
  if (hlb_dom_get_geometry_sizes(d, &n, &nr)) return 1;
  // sites in traversal order (Morton blocks, z-fastest inside): coords[trav] from the local order
  std::vector<int32_t> c(std::max<int64_t>(3 * n, 1));
  std::vector<uint32_t> trav(std::max<int64_t>(n, 1));
  if (n) {
    CU(cudaMemcpy(c.data(), d->coordsLocal, sizeof(int32_t) * 3 * n, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(trav.data(), d->travOfLocal, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost));
  }
  for (int64_t s = 0; s < n; ++s)
    for (int k = 0; k < 3; ++k) coords[3 * (int64_t)trav[s] + k] = c[(size_t)k * n + s];
  if (nr) {
    CU(cudaMemcpy(record_site, d->gRecSite, sizeof(int64_t) * nr, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(type, d->gType, sizeof(uint8_t) * 26 * nr, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(iolet, d->gIolet, sizeof(int32_t) * 26 * nr, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(dist, d->gDist, sizeof(float) * 26 * nr, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(navail, d->gNavail, sizeof(uint8_t) * nr, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(normal, d->gNormal, sizeof(float) * 3 * nr, cudaMemcpyDeviceToHost));
  }
  return 0;
}

int hlb_dom_build_seconds(hlb_dom_t d, double* s) {
  if (!d || !s) return fail("null argument");
  *s = d->buildSeconds;
  return 0;
}

int hlb_gpu_create_from_domain(hlb_dom_t d, const hlb_gpu_config* policy, hlb_gpu_t* out) {
  if (!d || !policy || !out) return fail("null argument");
  if (!d->built) return fail("domain not built");
  hlb_gpu_config cfg = *policy;
  const int Q = d->cfg.lattice;
  cfg.lattice = Q;
  cfg.device = d->cfg.device;
  cfg.rank = d->cfg.rank;
  cfg.nranks = d->cfg.nranks;
  cfg.n_sites = d->N;
  for (int t = 0; t < 6; ++t) { cfg.mid_count[t] = d->mid[t]; cfg.edge_count[t] = d->edge[t]; }
  cfg.total_shared_fs = d->S;
  cfg.n_neighbours = (int)d->procRank.size();
  hlb_gpu_t h = nullptr;
  if (hlb_gpu_create(&cfg, &h)) return 1;
  auto bail = [&]() { hlb_gpu_destroy(h); return 1; };
  hlb_gpu_raw raw;
  if (hlb_gpu_internal_raw(h, &raw)) return bail();
  const int64_t N = d->N, NB = d->NB;
  if (N) {
    nbr_engine_kernel<<<blocks_for(N * (Q - 1)), 256>>>(d->W, d->L, d->grid, d->coordsLocal, N, d->sendKey, d->sendSlot, d->S,
                                                        raw.stride, raw.nbr);
    if (cudaGetLastError() != cudaSuccess) { fail("nbr_engine_kernel launch failed"); return bail(); }
    if (raw.coordsAll) coords_planes_kernel<<<blocks_for(3 * N), 256>>>(d->coordsLocal, N, raw.stride, raw.coordsAll);
    if (cudaDeviceSynchronize() != cudaSuccess) { fail("device-side table install failed"); return bail(); }
  }
  if (NB != raw.NB) { fail("internal error: boundary site count mismatch"); return bail(); }
  if (NB) {
    std::vector<float> t((size_t)NB * std::max(Q - 1, 3));
    std::vector<int32_t> c(3 * (size_t)NB);
    if (cudaMemcpy(raw.hWall, d->bWall, sizeof(uint32_t) * NB, cudaMemcpyDeviceToHost) != cudaSuccess ||
        cudaMemcpy(raw.hIolet, d->bIolet, sizeof(uint32_t) * NB, cudaMemcpyDeviceToHost) != cudaSuccess ||
        cudaMemcpy(raw.hIoletId, d->bIoletId, sizeof(int32_t) * NB, cudaMemcpyDeviceToHost) != cudaSuccess) {
      fail("boundary table download failed");
      return bail();
    }
    cudaMemcpy(t.data(), d->bDist, sizeof(float) * NB * (Q - 1), cudaMemcpyDeviceToHost);
    for (int k = 0; k < Q - 1; ++k)
      for (int64_t b = 0; b < NB; ++b) raw.hCut[(size_t)k * raw.bStride + b] = t[(size_t)k * NB + b];
    cudaMemcpy(t.data(), d->bNormal, sizeof(float) * NB * 3, cudaMemcpyDeviceToHost);
    for (int k = 0; k < 3; ++k)
      for (int64_t b = 0; b < NB; ++b) raw.hNormal[(size_t)k * raw.bStride + b] = (double)t[(size_t)k * NB + b];
    // coordinates of the boundary-typed sites: two contiguous local ranges
    int64_t midTotal = 0;
    for (int t6 = 0; t6 < 6; ++t6) midTotal += d->mid[t6];
    const int64_t nbMid = midTotal - d->mid[0];
    for (int k = 0; k < 3; ++k) {
      if (nbMid)
        cudaMemcpy(c.data(), d->coordsLocal + (int64_t)k * N + d->mid[0], sizeof(int32_t) * nbMid, cudaMemcpyDeviceToHost);
      if (NB - nbMid)
        cudaMemcpy(c.data() + nbMid, d->coordsLocal + (int64_t)k * N + midTotal + d->edge[0], sizeof(int32_t) * (NB - nbMid),
                   cudaMemcpyDeviceToHost);
      for (int64_t b = 0; b < NB; ++b) raw.hCoords[(size_t)k * raw.bStride + b] = c[b];
    }
    if (cudaGetLastError() != cudaSuccess) { fail("boundary table download failed"); return bail(); }
  }
  if (hlb_gpu_internal_mark_installed(h)) return bail();
  if (cfg.n_neighbours) {
    if (hlb_gpu_set_neighbours(h, d->procRank.data(), d->procCount.data(), d->procFirst.data())) return bail();
    if (hlb_gpu_set_streaming_indices(h, d->streaming.data())) return bail();
  }
  *out = h;
  return 0;
}

}  // extern "C"
