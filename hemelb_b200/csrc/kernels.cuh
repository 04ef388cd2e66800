// Fused collide-and-stream kernels over HemeLB's sparse fluid-site arrays (sm_100a).
//
// One kernel instantiation per (lattice, collision kernel, wall link, inlet link, outlet link)
// policy bundle -- the device-side counterpart of the reference's
//   lb::BulkStreamer<C>                       Code/lb/streamers/BulkStreamer.h:57-99
//   lb::StreamerTypeFactory<WallLink,IoletLink> Code/lb/streamers/StreamerTypeFactory.h:24-109
// with C = lb::Normal<LBGK|MRT|TRT>.  One thread owns one site: it loads the Q pre-collision
// populations from the structure-of-arrays f_old (plane d at f + d*stride: every warp load is one
// fully coalesced 256 B request), collides in registers, and pushes each post-collision
// population through the 32-bit neighbour table into f_new.  Macroscopic-moment extraction
// (UpdateCachePostCollision, Common.h:21-130) is fused behind a run-time mask.
//
// Where the sites of a group of 32 allow it, the Q-1 push targets are not loaded per site but made
// from at most two runs per direction (StepArgs::nbrRuns): 8 B per direction and group.
//
// Site order on the device: inside the mid-domain part and inside the domain-edge part the sites
// of ALL six collision types are sorted together by lattice position (x, y, z) -- a wall site sits
// between the mid-fluid sites of its lattice row, so the 32 pushes of a warp land on consecutive
// addresses whatever the types of its sites, and the 32 B sectors of f_new leave L2 whole.  The
// reference's type-ordered ranges are an API-level view (perm / iperm); which link policy a site
// runs is decided per site from a bitmap of the boundary-typed sites (bInfo: one 8 B word per 32
// sites) and that site's wall / iolet masks, in the order of StreamerTypeFactory.h:65-79.
//
// Algorithmic HBM traffic per site update: Q*8 B read + Q*8 B written + (Q-1)*4 B of indices
// (380 B for D3Q19); with the targets as runs the site kernel moves 313 B.
#pragma once
#include <cuda.h>  // CUtensorMap
#include <cstdint>
#include <cfloat>
#include "lattice.cuh"

// D3Q27: ~168 registers -> 384 threads per SM (6 x 64: 11 381 MLUPS on the 1e8-site sac, 3 x 128: 11 275)
#ifndef HLB_Q27_THREADS
#define HLB_Q27_THREADS 64
#endif
#ifndef HLB_Q27_MIN_CTAS
#define HLB_Q27_MIN_CTAS 6
#endif

// CTA shape of the direct site kernel for Q <= 19: the register file holds 512 threads at 128 registers;
// measured on the 1.03e8-site tree / the cylinder (MLUPS, L2 prefetch 40 960 sites ahead):
// 2 x 256 threads 14 118 / 17 391, 4 x 128: 15 242 / 17 280, 8 x 64: 15 491 / 17 052 -- small CTAs
// retire and are replaced warp by warp, which evens out the mix of loading and computing warps
#ifndef HLB_MIN_CTAS
#define HLB_MIN_CTAS 8
#endif
#ifndef HLB_SITE_THREADS
#define HLB_SITE_THREADS 64
#endif

namespace hlb {

enum KernelKind { K_LBGK = 0, K_MRT = 1, K_TRT = 2 };
enum WallKind { W_SBB = 0, W_BFL = 1, W_GZS = 2, W_NONE = 3 };
enum IoletKind { I_NASH = 0, I_LADD = 1, I_NONE = 2 };
enum CacheBit {
  C_DENSITY = 1, C_VELOCITY = 2, C_WSS = 4, C_VONMISES = 8, C_SHEARRATE = 16, C_STRESS = 32,
  C_TRACTION = 64, C_TANGTRACTION = 128, C_MONITOR = 256
};
constexpr int kMonitorSlots = 4096;  // spread of the monitor atomics (4 x u64 per slot)

struct IoletDev {  // lb::iolets::InOutLet{Cosine,ParabolicVelocity} as the kernels see them
  double normal[3];    // normalised in double (InOutLet.h:160-163)
  double position[3];
  double radius, maxSpeed, warmUpLength;
  int kind;            // 0 pressure (cosine), 1 parabolic velocity
  int pad;
};

struct StepArgs {
  // distributions (SoA): population d of site s at f[d*stride + s]; f[Q*stride] is the rubbish
  // slot, f[Q*stride+1 ...] the halo slots (FieldData.cc:14-25 re-laid-out)
  const double* __restrict__ fOld;
  double* __restrict__ fNew;
  const uint32_t* __restrict__ nbr;  // (Q-1) planes: target of direction d at nbr[(d-1)*stride + s]
  int64_t stride;
  // which internal sites are boundary-typed (collision types 1..5), 32 sites per word:
  // bInfo[s >> 5] = {bitmap, boundary ordinal of the word's first boundary-typed site}
  const uint2* __restrict__ bInfo;
  // boundary-site tables, indexed by boundary ordinal b (see bidx()); ordinals ascend with the
  // internal site id
  const uint32_t* __restrict__ wallMask;
  const uint32_t* __restrict__ ioletMask;
  const int32_t* __restrict__ ioletId;   // SiteData::GetIoletId() | kOutletTypedBit for outlet / outlet-wall typed sites
  const float* __restrict__ cutDist;     // (Q-1) planes of bStride
  const double* __restrict__ wallNormal; // 3 planes of bStride
  const int32_t* __restrict__ coords;    // 3 planes of bStride
  int64_t bStride;
  // the same masks, iolet id and cut distances once more, one record per boundary-typed site
  // (brec_words<Q>() 32-bit words: wallMask, ioletMask, ioletId, 0, then the Q-1 float cut distances,
  // padded to 16 B): what the site kernel stages through shared memory with cp.async
  const uint4* __restrict__ bRec;
  // GZS: whole-site f_old rows of remote neighbours (NeighbouringDataManager); not used unless
  // WALL == W_GZS
  const int32_t* __restrict__ gzsNeighbour;  // (Q-1) planes of bStride: local site id, or -(g+1)
  const double* __restrict__ gzsGhost;       // ghost row g = Q consecutive doubles at gzsGhost[g*Q]
  // the two BoundaryValues objects (lb.hpp:87-113: inletValues for the inlet / inlet-wall
  // streamers, outletValues for the outlet / outlet-wall ones) + their per-step densities
  const IoletDev* __restrict__ iolets[2];
  const double* __restrict__ ioletDensity[2];  // GetBoundaryDensity(id) for this step
  uint64_t timeStep;                           // SimulationState::GetTimeStep() (1-indexed)
  // which link policies this launch applies (warp-uniform).  A whole-part launch runs every site
  // with the streamer of its own collision type (ioletSel = kIoletByType, wallOn = 1); a launch for
  // one streamer slot (StreamerTests-style sub-range calls, plain order) runs that streamer on
  // whatever sites it is given: wallOn = the slot has a wall link; ioletSel = kIoletNone, or 0 / 1 =
  // every iolet link goes to the inlet / outlet object with the inlet / outlet link policy
  int wallOn, ioletSel;
  // LbmParameters
  double tau, omega, stressParameter, omegaMinus;
  // caches (MacroscopicPropertyCache), site-major like the reference
  uint32_t cacheMask;
  double* __restrict__ cDensity;
  double* __restrict__ cVelocity;
  double* __restrict__ cWss;
  double* __restrict__ cVonMises;
  double* __restrict__ cShearRate;
  double* __restrict__ cStress;
  double* __restrict__ cTraction;
  double* __restrict__ cTangTraction;
  const uint32_t* __restrict__ refSiteOf;  // internal site -> reference site id (cache rows), or null
  // optional explicit site list (launches that are not a contiguous run of internal sites)
  const uint32_t* __restrict__ siteList;
  // contiguous launches of the direct site kernel: every CTA asks L2 for what the CTA that starts
  // `prefetchSites` sites further down the grid will load (0: off)
  int prefetchSites;
  // whole-part launches: the push targets of 32 consecutive sites as at most two runs per direction.
  // runFlags bit w: every direction of sites [32w, 32w+32) is {base, split, delta} in
  // nbrRuns[w*(Q-1) + d-1] = {base, delta << 5 | split}: target(lane) = base + lane + (lane >= split ? delta : 0)
  // for every link the launch pushes (cut links are not pushed); else the sites read the nbr planes.
  // Null: off (sub-range launches apply their streamer's policies to whatever sites they get, and push
  // the cut links of the others to the rubbish slot through nbr)
  const uint2* __restrict__ nbrRuns;
  const uint32_t* __restrict__ runFlags;
  // GuoZhengShi per-link kernel: at most this many CTAs, each walking several tiles (0: one CTA per tile)
  int gzsGridLimit;
  // fused monitors (C_MONITOR): {min f, min rho, max rho, max u^2} as order-preserving u64 keys
  unsigned long long* __restrict__ monitorSlots;
};
constexpr int kIoletNone = -1, kIoletByType = 2;
constexpr int32_t kOutletTypedBit = 1 << 30;

template <int Q> __host__ __device__ constexpr int brec_words() { return ((4 + Q - 1) + 3) / 4 * 4; }
template <int Q> __host__ __device__ constexpr int site_threads() { return Q > 19 ? HLB_Q27_THREADS : HLB_SITE_THREADS; }
template <int Q> __host__ __device__ constexpr int site_min_ctas() { return Q > 19 ? HLB_Q27_MIN_CTAS : HLB_MIN_CTAS; }

template <int Q> struct MrtArgs {
  double SMn[mrt_k<Q>() > 0 ? mrt_k<Q>() : 1][Q];  // collisionMatrixDiagonals[k] * normalisedReducedMomentBasis[k][d]
};

// boundary ordinal of an internal site, or -1 for a bulk-typed one
__device__ __forceinline__ int64_t bidx(const StepArgs& A, int64_t site) {
  const uint2 bi = __ldg(A.bInfo + (site >> 5));
  const unsigned lane = (unsigned)site & 31u;
  return ((bi.x >> lane) & 1u) ? (int64_t)bi.y + __popc(bi.x & ((1u << lane) - 1u)) : -1;
}

// order-preserving map double -> u64 (so atomicMin/atomicMax on integers order doubles)
__device__ __forceinline__ unsigned long long mon_enc(double x) {
  const unsigned long long u = (unsigned long long)__double_as_longlong(x);
  return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double mon_dec(unsigned long long u) {
  u = (u & 0x8000000000000000ull) ? (u & 0x7fffffffffffffffull) : ~u;
  return __longlong_as_double((long long)u);
}

// StabilityTester / IncompressibilityChecker inputs gathered while the populations are in
// registers (Code/lb/StabilityTester.h:97-113: any f <= 0; IncompressibilityChecker: density
// extrema, max speed): warp-reduce, then one relaxed atomic per value per warp, spread over slots.
template <int Q>
__device__ __forceinline__ double min_population(const double (&f)[Q]) {
  double smallest = f[0];
#pragma unroll
  for (int d = 1; d < Q; ++d) smallest = fmin(smallest, f[d]);  // (one DMNMX each)
  return smallest;
}
// smallest / largest order-preserving key of a full warp: two 32-bit warp reductions (redux.sync) per
// value -- the high words first, then the low words of the lanes that hold the winning high word
__device__ __forceinline__ unsigned long long warp_min_key(unsigned long long k) {
  const unsigned hi = (unsigned)(k >> 32);
  const unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
  const unsigned mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? (unsigned)k : 0xffffffffu);
  return ((unsigned long long)mhi << 32) | mlo;
}
__device__ __forceinline__ unsigned long long warp_max_key(unsigned long long k) {
  const unsigned hi = (unsigned)(k >> 32);
  const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
  const unsigned mlo = __reduce_max_sync(0xffffffffu, hi == mhi ? (unsigned)k : 0u);
  return ((unsigned long long)mhi << 32) | mlo;
}
__device__ __forceinline__ void fused_monitor(const StepArgs& A, int64_t tid, double fmin, double rho, double u2) {
  unsigned long long kf = mon_enc(fmin), krmin = mon_enc(rho), krmax = krmin, ku = mon_enc(u2);
  const unsigned mask = __activemask();
  unsigned long long* slot = A.monitorSlots + 4 * ((tid >> 5) & (kMonitorSlots - 1));
  if (mask == 0xffffffffu) {
    kf = warp_min_key(kf);
    krmin = warp_min_key(krmin);
    krmax = warp_max_key(krmax);
    ku = warp_max_key(ku);
    if ((threadIdx.x & 31) != 0) return;
  }
  atomicMin(slot + 0, kf);
  atomicMin(slot + 1, krmin);
  atomicMax(slot + 2, krmax);
  atomicMax(slot + 3, ku);
}

// ---------------------------------------------------------------------------------- collisions
template <int Q, int KERNEL> struct HydroVars {
  double rho, m[3], u[3];
  double f[Q], fneq[Q], fpost[Q];
};

// MRT.h:123-134
template <int Q>
__device__ __forceinline__ void mrt_project(const double (&v)[Q], double (&mom)[mrt_k<Q>() > 0 ? mrt_k<Q>() : 1]) {
#pragma unroll
  for (int k = 0; k < mrt_k<Q>(); ++k) {
    double acc = 0.;
#pragma unroll
    for (int d = 0; d < Q; ++d)
      if (mrt_m<Q>(k, d) != 0) acc += double(mrt_m<Q>(k, d)) * v[d];
    mom[k] = acc;
  }
}

// LBGK.h:56-63 / MRT.h:88-105 / TRT.h:94-121 on (f, fneq) -> fpost
template <int Q, int KERNEL>
__device__ __forceinline__ void collide(const StepArgs& A, const MrtArgs<Q>& M, const double (&f)[Q],
                                        const double (&fneq)[Q], double (&fpost)[Q]) {
  if constexpr (KERNEL == K_LBGK) {
#pragma unroll
    for (int d = 0; d < Q; ++d) fpost[d] = f[d] + fneq[d] * A.omega;
  } else if constexpr (KERNEL == K_MRT) {
    double mneq[mrt_k<Q>() > 0 ? mrt_k<Q>() : 1];
    mrt_project<Q>(fneq, mneq);
#pragma unroll
    for (int d = 0; d < Q; ++d) {
      double collision = 0.;
#pragma unroll
      for (int k = 0; k < mrt_k<Q>(); ++k)
        if (mrt_m<Q>(k, d) != 0) collision += M.SMn[k][d] * mneq[k];
      fpost[d] = f[d] - collision;
    }
  } else {
    fpost[0] = f[0] + A.omega * fneq[0];
#pragma unroll
    for (int i = 1; i < Q; i += 2) {
      const int ib = i + 1;
      const double sym = 0.5 * A.omega * (fneq[i] + fneq[ib]);
      const double asym = 0.5 * A.omegaMinus * (fneq[i] - fneq[ib]);
      fpost[i] = f[i] + sym + asym;
      fpost[ib] = f[ib] + sym - asym;
    }
  }
}

// ---------------------------------------------------------------------------------- stresses
// Lattice.h:716-746 CalculatePiTensor (lower triangle, then mirrored)
template <int Q>
__device__ __forceinline__ void pi_tensor(const double (&f)[Q], double (&pi)[3][3]) {
#pragma unroll
  for (int ii = 0; ii < 3; ++ii)
#pragma unroll
    for (int jj = 0; jj <= ii; ++jj) {
      double acc = 0.0;
#pragma unroll
      for (int l = 0; l < Q; ++l) {
        const int ci = ii == 0 ? Lat<Q>::cx(l) : (ii == 1 ? Lat<Q>::cy(l) : Lat<Q>::cz(l));
        const int cj = jj == 0 ? Lat<Q>::cx(l) : (jj == 1 ? Lat<Q>::cy(l) : Lat<Q>::cz(l));
        if (ci * cj != 0) acc += cmul(ci * cj, f[l]);  // (f*ci)*cj with ci,cj in {-1,0,1}
      }
      pi[ii][jj] = acc;
    }
  pi[0][1] = pi[1][0];
  pi[0][2] = pi[2][0];
  pi[1][2] = pi[2][1];
}

template <int Q>
__device__ __forceinline__ void stress_tensor(double rho, double tau, const double (&fneq)[Q], double (&s)[3][3]) {
  pi_tensor<Q>(fneq, s);  // Lattice.h:622-637
  const double fac = 1 - 1 / (2 * tau);
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) s[r][c] *= fac;
  const double pressure = (rho - 1) * kCs2;
#pragma unroll
  for (int r = 0; r < 3; ++r) s[r][r] += pressure;
}

template <int Q, int KERNEL>
__device__ __forceinline__ void update_caches(const StepArgs& A, int64_t site, int64_t b, double rho,
                                           const double (&u)[3], const double (&fneq)[Q]) {
  // UpdateCachePostCollision, Common.h:21-130
  const int64_t row = A.refSiteOf ? A.refSiteOf[site] : site;
  const uint32_t mask = A.cacheMask;
  bool isWall = false;
  double nor[3] = {0, 0, 0};
  if (b >= 0) {
    isWall = A.wallMask[b] != 0;
    nor[0] = A.wallNormal[b];
    nor[1] = A.wallNormal[A.bStride + b];
    nor[2] = A.wallNormal[2 * A.bStride + b];
  }
  if (mask & C_DENSITY) A.cDensity[row] = rho;
  if (mask & C_VELOCITY) {
    A.cVelocity[3 * row] = u[0];
    A.cVelocity[3 * row + 1] = u[1];
    A.cVelocity[3 * row + 2] = u[2];
  }
  if (mask & C_WSS) {
    double stress = DBL_MAX;  // NO_VALUE
    if (isWall) {  // Lattice.h:650-688
      double sv[3] = {0.0, 0.0, 0.0};
      double sq = 0.0, ns = 0.0;
      const double temp = A.stressParameter * (-sqrt(2.0));
      double pi[3][3];
      pi_tensor<Q>(fneq, pi);
#pragma unroll
      for (int i = 0; i < 3; i++) {
#pragma unroll
        for (int j = 0; j < 3; j++) sv[i] += pi[i][j] * nor[j] * temp;
        sq += sv[i] * sv[i];
        ns += sv[i] * nor[i];
      }
      stress = sqrt(sq - ns * ns);
    }
    A.cWss[row] = stress;
  }
  if (mask & C_VONMISES) {  // Lattice.h:510-551
    double xx_yy = 0.0, yy_zz = 0.0, xx_zz = 0.0, xy = 0.0, xz = 0.0, yz = 0.0;
#pragma unroll
    for (int d = 0; d < Q; ++d) {
      const int cx = Lat<Q>::cx(d), cy = Lat<Q>::cy(d), cz = Lat<Q>::cz(d);
      if (cx * cx - cy * cy != 0) xx_yy += cmul(cx * cx - cy * cy, fneq[d]);
      if (cy * cy - cz * cz != 0) yy_zz += cmul(cy * cy - cz * cz, fneq[d]);
      if (cx * cx - cz * cz != 0) xx_zz += cmul(cx * cx - cz * cz, fneq[d]);
      if (cx * cy != 0) xy += cmul(cx * cy, fneq[d]);
      if (cx * cz != 0) xz += cmul(cx * cz, fneq[d]);
      if (cy * cz != 0) yz += cmul(cy * cz, fneq[d]);
    }
    const double a = xx_yy * xx_yy + yy_zz * yy_zz + xx_zz * xx_zz;
    const double b = xy * xy + xz * xz + yz * yz;
    A.cVonMises[row] = A.stressParameter * sqrt(a + 6.0 * b);
  }
  if (mask & C_SHEARRATE) {  // Lattice.h:748-770, 890-908
    double shear = 0.0;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = r; c < 3; ++c) {
        double s = 0.0;
#pragma unroll
        for (int v = 0; v < Q; ++v) {
          const int cr = r == 0 ? Lat<Q>::cx(v) : (r == 1 ? Lat<Q>::cy(v) : Lat<Q>::cz(v));
          const int cc = c == 0 ? Lat<Q>::cx(v) : (c == 1 ? Lat<Q>::cy(v) : Lat<Q>::cz(v));
          if (cr * cc != 0) s += cmul(cr * cc, fneq[v]);
        }
        s *= -1.0 / (2.0 * A.tau * rho * kCs2);
        shear += (c == r) ? s * s : 2 * s * s;
      }
    A.cShearRate[row] = sqrt(2 * shear);
  }
  if (mask & (C_STRESS | C_TRACTION | C_TANGTRACTION)) {
    double s[3][3];
    stress_tensor<Q>(rho, A.tau, fneq, s);
    if (mask & C_STRESS)
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) A.cStress[9 * row + 3 * a + b] = s[a][b];
    double t[3] = {0, 0, 0}, tt[3] = {0, 0, 0};
    if (isWall) {  // Lattice.h:566-608
      double mag = 0.0;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        double acc = 0.0;
#pragma unroll
        for (int b = 0; b < 3; ++b) acc += s[a][b] * nor[b];
        t[a] = acc;
      }
#pragma unroll
      for (int a = 0; a < 3; ++a) mag += t[a] * nor[a];
#pragma unroll
      for (int a = 0; a < 3; ++a) tt[a] = t[a] - nor[a] * mag;
    }
    if (mask & C_TRACTION)
      for (int a = 0; a < 3; ++a) A.cTraction[3 * row + a] = t[a];
    if (mask & C_TANGTRACTION)
      for (int a = 0; a < 3; ++a) A.cTangTraction[3 * row + a] = tt[a];
  }
}

// ---------------------------------------------------------------------------------- iolet links
// InOutLetParabolicVelocity.cc:23-43
__device__ __forceinline__ void parabolic_velocity(const IoletDev& io, const double (&x)[3], uint64_t t, double (&v)[3]) {
  const double d0 = x[0] - io.position[0], d1 = x[1] - io.position[1], d2 = x[2] - io.position[2];
  double z = 0.0;
  z += d0 * io.normal[0];
  z += d1 * io.normal[1];
  z += d2 * io.normal[2];
  double mag2 = 0.0;
  mag2 += d0 * d0;
  mag2 += d1 * d1;
  mag2 += d2 * d2;
  const double rSq = (mag2 - z * z) / (io.radius * io.radius);
  double mx = io.maxSpeed;
  if ((double)t < io.warmUpLength) mx *= t / io.warmUpLength;
  const double s = mx * (1. - rSq);
  v[0] = io.normal[0] * s;
  v[1] = io.normal[1] * s;
  v[2] = io.normal[2] * s;
}

// ---------------------------------------------------------------------------------- GZS wall link
// GuoZhengShi.h:123-284.  Every wall link re-runs a whole collision on an extrapolated wall node,
// of which only ONE post-collision population is kept.  So the links do not ride in the per-site
// thread (that serialises up to Q-1 collisions behind one another): gzs_links_kernel gives every
// (site, wall direction) pair its own thread, recomputes the site's hydrodynamic variables from
// f_old, and evaluates just the component it streams -- the same operations, in the same order,
// as the reference performs for that component.

// What one wall link needs of the wall node's f_neq (GuoZhengShi.h:228-262): the site's own f_neq,
// blended with the fluid neighbour's when the link extrapolates.  Evaluated per component, on
// demand: LBGK needs one component, TRT a pair, MRT all of them.
template <int Q> struct GzsNode {
  const double (*sf)[128];  // the tile's staged f_old rows, [direction][site in tile]
  int t;
  double rho, density_1, mm, m[3];
  bool blend;
  double q;
  const double* nrow;  // neighbour's f_old row
  int64_t nstep;
  double nrho, ndensity_1, nmm, nm[3];
  template <int J> __device__ __forceinline__ double fneq() const {
    double v = sf[J][t] - feq_i<Q>(J, rho, density_1, mm, m);
    if (blend) v = q * v + (1. - q) * (nrow[J * nstep] - feq_i<Q>(J, nrho, ndensity_1, nmm, nm));
    return v;
  }
};

// component D of kernel.Collide on the wall node (or, for the SBB fall-back, on the site itself):
// base = f[D] of that node
template <int Q, int KERNEL, int D>
__device__ __forceinline__ double gzs_component(const StepArgs& A, const MrtArgs<Q>& M, const GzsNode<Q>& N, bool sbb,
                                                double wdensity_1, double wmm, const double (&mw)[3],
                                                const double (&fneqW)[Q],
                                                const double (&mneq)[mrt_k<Q>() > 0 ? mrt_k<Q>() : 1]) {
  // fneqW / mneq are filled for MRT only (once per link, before the per-direction dispatch, so
  // that divergent warps do not repeat the projection); the other kernels evaluate the
  // components they touch
  const double fn = KERNEL == K_MRT ? fneqW[D] : N.template fneq<D>();
  const double base = sbb ? N.sf[D][N.t] : feq_i<Q>(D, N.rho, wdensity_1, wmm, mw) + fn;
  if constexpr (KERNEL == K_LBGK) {
    return base + fn * A.omega;
  } else if constexpr (KERNEL == K_MRT) {
    double collision = 0.;  // MRT.h:88-105
#pragma unroll
    for (int k = 0; k < mrt_k<Q>(); ++k)
      if (mrt_m<Q>(k, D) != 0) collision += M.SMn[k][D] * mneq[k];
    return base - collision;
  } else {
    constexpr int a = (D & 1) ? D : D - 1, b = a + 1;
    const double fa = D == a ? fn : N.template fneq<a>(), fb = D == b ? fn : N.template fneq<b>();
    const double sym = 0.5 * A.omega * (fa + fb);
    const double asym = 0.5 * A.omegaMinus * (fa - fb);
    return D == a ? base + sym + asym : base + sym - asym;
  }
}

// compile-time walk over the directions: the one that equals `o` evaluates its component
template <int Q, int KERNEL, int D>
__device__ __forceinline__ void gzs_pick(const StepArgs& A, const MrtArgs<Q>& M, const GzsNode<Q>& N, int o, bool sbb,
                                         double wdensity_1, double wmm, const double (&mw)[3],
                                         const double (&fneqW)[Q],
                                         const double (&mneq)[mrt_k<Q>() > 0 ? mrt_k<Q>() : 1], double& out) {
  if constexpr (D < Q) {
    if (D == o) out = gzs_component<Q, KERNEL, D>(A, M, N, sbb, wdensity_1, wmm, mw, fneqW, mneq);
    gzs_pick<Q, KERNEL, D + 1>(A, M, N, o, sbb, wdensity_1, wmm, mw, fneqW, mneq, out);
  }
}

// all components at once (MRT)
template <int Q, int J = 0>
__device__ __forceinline__ void gzs_fill(const GzsNode<Q>& N, double (&fneqW)[Q]) {
  if constexpr (J < Q) {
    fneqW[J] = N.template fneq<J>();
    gzs_fill<Q, J + 1>(N, fneqW);
  }
}

// One CTA owns a tile of kGzsTile consecutive sites of the launch: their f_old rows are staged
// through shared memory once, the tile's wall links are compacted into a list (so warps are full
// whatever the wall orientation), and the threads walk the list.
constexpr int kGzsTile = 128;  // (GzsNode::sf is typed on it)
constexpr int kGzsThreads = 256;  // LBGK on Q <= 19 fits 85 registers (three CTAs per SM), the rest needs two

// The launch covers `count` boundary-typed sites: siteList[0 .. count) (a whole part: the part's
// slice of the boundary-site list), or the consecutive internal sites from `first`.
template <int Q, int KERNEL>
__global__ void __launch_bounds__(kGzsThreads, (KERNEL == K_LBGK && Q <= 19) ? 3 : 2) gzs_links_kernel(const StepArgs A, const MrtArgs<Q> M, int64_t first,
                                                               int64_t count) {
  constexpr int T = kGzsTile;
  __shared__ double sf[Q][T];
  __shared__ double srho[T], smom[3][T];
  __shared__ int64_t ssite[T];
  __shared__ int32_t sb[T];
  __shared__ uint32_t swall[T], siolet[T];
  __shared__ uint16_t slink[T * (Q - 1)];
  __shared__ int soff[(Q - 1) * (T / 32) + 1];
  const int tx = threadIdx.x;
  // (one tile per CTA; or, with a grid smaller than the tiles -- StepArgs::gzsGridLimit: the launch that
  // shares the SMs with the site kernel -- every CTA walks its tiles)
  const int nTiles = (int)((count + T - 1) / T);
  for (int tile = blockIdx.x; tile < nTiles; tile += gridDim.x) {
  const int64_t tile0 = (int64_t)tile * T;
  const int nT = (int)((count - tile0) < T ? (count - tile0) : T);
  if (tx < T) {
    uint32_t wall = 0, iol = 0;
    if (tx < nT) {
      const int64_t site = A.siteList ? (int64_t)A.siteList[tile0 + tx] : first + tile0 + tx;
      const int64_t b = bidx(A, site);
      ssite[tx] = site;
      sb[tx] = (int32_t)b;
      if (b >= 0) {
        wall = A.wallMask[b];
        if (A.ioletSel != kIoletNone) iol = A.ioletMask[b];
      }
    }
    swall[tx] = wall;
    siolet[tx] = iol;
  }
  __syncthreads();
  for (int e = tx; e < Q * T; e += kGzsThreads) {
    const int d = e / T, t = e % T;
    if (t < nT) sf[d][t] = A.fOld[(int64_t)d * A.stride + ssite[t]];
  }
  __syncthreads();
  // per site: density and momentum; its GZS links = wall links that are not iolet links
  // (the iolet link takes precedence, StreamerTypeFactory.h:65-79)
  uint32_t links = 0;
  if (tx < nT) {
    double f[Q];
#pragma unroll
    for (int d = 0; d < Q; ++d) f[d] = sf[d][tx];
    double rho, m[3];
    density_momentum<Q>(f, rho, m);
    srho[tx] = rho;
    smom[0][tx] = m[0];
    smom[1][tx] = m[1];
    smom[2][tx] = m[2];
    links = swall[tx] & ~siolet[tx];
  }
  // the tile's link list, direction-major (so a warp works on one direction and the per-direction
  // specialisations below do not diverge): position = links of earlier directions + earlier sites
  constexpr int SW = T / 32;  // warps that hold sites
  const int warp = tx >> 5, lane = tx & 31;
  if (warp < SW) {
#pragma unroll
    for (int d = 1; d < Q; ++d) {
      const unsigned ballot = __ballot_sync(0xffffffffu, (links >> (d - 1)) & 1u);
      if (lane == 0) soff[(d - 1) * SW + warp] = __popc(ballot);
    }
  }
  __syncthreads();
  if (tx == 0) {
    int acc = 0;
    for (int e = 0; e < (Q - 1) * SW; ++e) {
      const int c = soff[e];
      soff[e] = acc;
      acc += c;
    }
    soff[(Q - 1) * SW] = acc;
  }
  __syncthreads();
  const int total = soff[(Q - 1) * SW];
  if (warp < SW) {
#pragma unroll
    for (int d = 1; d < Q; ++d) {
      const bool bit = (links >> (d - 1)) & 1u;
      const unsigned ballot = __ballot_sync(0xffffffffu, bit);
      if (bit) slink[soff[(d - 1) * SW + warp] + __popc(ballot & ((1u << lane) - 1u))] = (uint16_t)(tx | (d << 8));
    }
  }
  __syncthreads();

  for (int k = tx; k < total; k += kGzsThreads) {
    const int t = slink[k] & 255;
    const int iPrime = slink[k] >> 8;  // the direction that hits the wall
    const int i = inv_dir(iPrime);
    const int64_t site = ssite[t];
    const int64_t b = sb[t];
    const uint32_t wallMask = swall[t], ioletMask = siolet[t];
    GzsNode<Q> N;
    N.sf = sf;
    N.t = t;
    N.rho = srho[t];
    N.m[0] = smom[0][t];
    N.m[1] = smom[1][t];
    N.m[2] = smom[2][t];
    N.density_1 = 1. / N.rho;
    N.mm = N.m[0] * N.m[0] + N.m[1] * N.m[1] + N.m[2] * N.m[2];
    N.blend = false;
    N.nrow = A.fOld;
    N.nstep = 0;
    N.nrho = N.ndensity_1 = 1.;
    N.nmm = 0.;
    N.nm[0] = N.nm[1] = N.nm[2] = 0.;
    const double rho = N.rho;
    const double q = (double)A.cutDist[(int64_t)(iPrime - 1) * A.bStride + b];
    N.q = q;
    double mw[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) mw[c] = N.m[c] * (1. - 1. / q);
    bool sbb = false;
    if (q < 0.75) {
      const bool hasIoletI = (ioletMask >> (i - 1)) & 1u;
      const bool hasWallI = (wallMask >> (i - 1)) & 1u;
      if (hasIoletI) {
        const int32_t enc = A.ioletId[b];
        const int sel = A.ioletSel == kIoletByType ? ((enc & kOutletTypedBit) ? 1 : 0) : A.ioletSel;
        const IoletDev& io = A.iolets[sel][enc & (kOutletTypedBit - 1)];
        if (io.kind != 1) {
          sbb = true;
        } else {
          double np[3], nv[3];
          np[0] = (double)A.coords[b] + (double)Lat<Q>::cx(i);
          np[1] = (double)A.coords[A.bStride + b] + (double)Lat<Q>::cy(i);
          np[2] = (double)A.coords[2 * A.bStride + b] + (double)Lat<Q>::cz(i);
          parabolic_velocity(io, np, A.timeStep, nv);
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const double second = nv[c] * (q - 1) / (q + 1);
            mw[c] = q * mw[c] + (1. - q) * rho * second;
          }
        }
      } else if (hasWallI) {
        sbb = true;
      } else {
        // neighbour site's f_old: local plane read, or the phase-0 ghost row of a remote site
        // (read again, from L1, by the components that blend it in)
        const int32_t n = A.gzsNeighbour[(int64_t)(i - 1) * A.bStride + b];
        N.nrow = n >= 0 ? A.fOld + n : A.gzsGhost + (-(int64_t)n - 1) * Q;
        N.nstep = n >= 0 ? A.stride : 1;
        double nrho = 0.0, nm[3] = {0.0, 0.0, 0.0}, nu[3];
#pragma unroll
        for (int j = 0; j < Q; ++j) {  // Lattice.h:181-191, as density_momentum()
          const double v = N.nrow[j * N.nstep];
          nrho += v;
          if (Lat<Q>::cx(j) != 0) nm[0] += cmul(Lat<Q>::cx(j), v);
          if (Lat<Q>::cy(j) != 0) nm[1] += cmul(Lat<Q>::cy(j), v);
          if (Lat<Q>::cz(j) != 0) nm[2] += cmul(Lat<Q>::cz(j), v);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) nu[c] = nm[c] / nrho;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const double second = nu[c] * (q - 1) / (q + 1);
          mw[c] = q * mw[c] + (1. - q) * rho * second;
        }
        N.blend = true;
        N.nrho = nrho;
        N.ndensity_1 = 1. / nrho;
        N.nmm = nm[0] * nm[0] + nm[1] * nm[1] + nm[2] * nm[2];
        N.nm[0] = nm[0];
        N.nm[1] = nm[1];
        N.nm[2] = nm[2];
      }
    }
    // (the reference leaves m_neq of the wall node's HydroVars unset for MRT; we project its f_neq)
    double fneqW[Q];
    double mneq[mrt_k<Q>() > 0 ? mrt_k<Q>() : 1];
    fneqW[0] = mneq[0] = 0.;
    if constexpr (KERNEL == K_MRT) {
      gzs_fill<Q>(N, fneqW);
      mrt_project<Q>(fneqW, mneq);
    }
    // SBB: component iPrime of the site's own collision; else component i of the wall node's
    const int o = sbb ? iPrime : i;
    const double wdensity_1 = 1. / rho;
    const double wmm = mw[0] * mw[0] + mw[1] * mw[1] + mw[2] * mw[2];
    double out = 0.;
    gzs_pick<Q, KERNEL, 1>(A, M, N, o, sbb, wdensity_1, wmm, mw, fneqW, mneq, out);
    A.fNew[(int64_t)i * A.stride + site] = out;
  }
  __syncthreads();  // (the staging area is reused by the CTA's next tile)
  }
}

// ---------------------------------------------------------------------------------- the site kernel
// Q <= 19: 64-thread CTAs, eight resident per SM (<= 128 registers).  D3Q27 needs ~190 registers:
// 64-thread CTAs, six resident (<= 168 registers; 12 warps per SM instead of 16).

// per-site quantities of an iolet link policy (NashZerothOrderPressure.h:27-60 / LaddIolet.h:29-66)
struct IoletSite {
  const IoletDev* io;
  double ghostRho, ghostD1, ghostMM, ghostM[3];  // Nash: the ghost site's density and momentum
  double sx, sy, sz;                             // Ladd: the site's lattice position
};

template <int Q, int IOLET>
__device__ __forceinline__ void iolet_site(const StepArgs& A, int sel, int id, int64_t b, double rho,
                                           const double (&m)[3], IoletSite& S) {
  S.io = A.iolets[sel] + id;
  if constexpr (IOLET == I_NASH) {
    S.ghostRho = A.ioletDensity[sel][id];
    const float nf0 = (float)S.io->normal[0], nf1 = (float)S.io->normal[1], nf2 = (float)S.io->normal[2];
    double dot = 0.0;
    dot += m[0] * (double)nf0;
    dot += m[1] * (double)nf1;
    dot += m[2] * (double)nf2;
    const double component = dot / rho;
    S.ghostM[0] = ((double)nf0 * component) * S.ghostRho;
    S.ghostM[1] = ((double)nf1 * component) * S.ghostRho;
    S.ghostM[2] = ((double)nf2 * component) * S.ghostRho;
    S.ghostD1 = 1. / S.ghostRho;
    S.ghostMM = S.ghostM[0] * S.ghostM[0] + S.ghostM[1] * S.ghostM[1] + S.ghostM[2] * S.ghostM[2];
  } else {
    S.sx = (double)A.coords[b];
    S.sy = (double)A.coords[A.bStride + b];
    S.sz = (double)A.coords[2 * A.bStride + b];
  }
}

// what an iolet link in direction D leaves in f_new[site, inv(D)]
template <int Q, int IOLET, int D>
__device__ __forceinline__ double iolet_link(const StepArgs& A, const IoletSite& S, double rho, double fpostD) {
  constexpr int id = inv_dir(D);
  if constexpr (IOLET == I_NASH) {
    return feq_i<Q>(id, S.ghostRho, S.ghostD1, S.ghostMM, S.ghostM);
  } else {
    double x[3] = {S.sx + 0.5 * Lat<Q>::cx(D), S.sy + 0.5 * Lat<Q>::cy(D), S.sz + 0.5 * Lat<Q>::cz(D)};
    double wallMom[3];
    parabolic_velocity(*S.io, x, A.timeStep, wallMom);
    wallMom[0] *= rho;
    wallMom[1] *= rho;
    wallMom[2] *= rho;
    double dot = 0.0;
    dot += wallMom[0] * (double)Lat<Q>::cx(D);
    dot += wallMom[1] * (double)Lat<Q>::cy(D);
    dot += wallMom[2] * (double)Lat<Q>::cz(D);
    const double correction = 2. * Lat<Q>::W(D) * dot / kCs2;
    return fpostD - correction;
  }
}

// the cut links of one boundary-typed site, after its uncut links were pushed: iolet link first,
// then wall link (StreamerTypeFactory.h:65-79)
template <int Q, int WALL, int INLET, int OUTLET, int RSTRIDE, int D>
__device__ __forceinline__ void cut_links(const StepArgs& A, int64_t site, const uint32_t* __restrict__ rec,
                                          uint32_t wallMask, uint32_t ioletMask, int sel, const IoletSite& S,
                                          double rho, const double (&fpost)[Q]) {
  if constexpr (D < Q) {
    constexpr int id = inv_dir(D);
    if ((ioletMask >> (D - 1)) & 1u) {
      double v;
      if constexpr (INLET == OUTLET) v = iolet_link<Q, INLET, D>(A, S, rho, fpost[D]);
      else v = sel ? iolet_link<Q, OUTLET, D>(A, S, rho, fpost[D]) : iolet_link<Q, INLET, D>(A, S, rho, fpost[D]);
      A.fNew[(int64_t)id * A.stride + site] = v;
    } else if ((wallMask >> (D - 1)) & 1u) {
      if constexpr (WALL == W_SBB) {  // SimpleBounceBack.h:23-42
        A.fNew[(int64_t)id * A.stride + site] = fpost[D];
      } else if constexpr (WALL == W_BFL) {  // BouzidiFirdaousLallemand.h:41-70
        // word 4 + (D - 1) of the staged record: chunk (3 + D) / 4 of the thread's column
        const double q = (double)__uint_as_float(rec[((3 + D) / 4) * (4 * RSTRIDE) + ((3 + D) & 3)]);
        const bool invWall = (wallMask >> (id - 1)) & 1u;
        double v;
        if (invWall || q < 0.5) v = fpost[D];
        else v = (fpost[D] + (2.0 * q - 1) * fpost[id]) / (2.0 * q);
        A.fNew[(int64_t)id * A.stride + site] = v;
      }
      // W_GZS: gzs_links_kernel owns this population
    }
    cut_links<Q, WALL, INLET, OUTLET, RSTRIDE, D + 1>(A, site, rec, wallMask, ioletMask, sel, S, rho, fpost);
  }
}

// The thread's index in the launch, read again from the special registers.  The rare blocks at
// the end of the site kernel (moment extraction, monitors) use it instead of keeping the values
// derived from the first read alive across the whole kernel (two registers that otherwise spill).
__device__ __forceinline__ int64_t launch_tid_again() {
  unsigned t, c, n;
  asm volatile("mov.u32 %0, %%tid.x;" : "=r"(t));
  asm volatile("mov.u32 %0, %%ctaid.x;" : "=r"(c));
  asm volatile("mov.u32 %0, %%ntid.x;" : "=r"(n));
  return (int64_t)c * n + t;
}

// The rare tail of the site kernel (moment extraction on output steps) as a real call,
// so that the hot path reserves no registers for it.  The moments are extracted from f_old read
// again (it is not written during a step): the same loads through the same arithmetic as the
// site's collision, hence the same bits, but neither f nor f_neq has to outlive the collision.
template <int Q, int KERNEL>
__device__ __noinline__ void site_tail(const StepArgs& A, int64_t site, int b) {
  {
    double f[Q];
#pragma unroll
    for (int d = 0; d < Q; ++d) f[d] = A.fOld[(int64_t)d * A.stride + site];
    double rho, m[3], u[3], fneq[Q];
    density_momentum<Q>(f, rho, m);
#pragma unroll
    for (int k = 0; k < 3; ++k) u[k] = m[k] / rho;
    const double density_1 = 1. / rho;
    const double mm = m[0] * m[0] + m[1] * m[1] + m[2] * m[2];
#pragma unroll
    for (int d = 0; d < Q; ++d) fneq[d] = f[d] - feq_i<Q>(d, rho, density_1, mm, m);
    update_caches<Q, KERNEL>(A, site, b, rho, u, fneq);
  }
}

// One site, once its populations and push targets are in registers: collide, stream, (rarely)
// extract moments.  `rec` is the thread's column of the shared-memory staging area for boundary
// records (chunk c of the record at rec[4 * RSTRIDE * c ...]); with DIRECT the record is still on its
// way (cp.async) and is waited for after the collision, and the tail works out the site again from
// the launch index instead of keeping it alive.
//
// The push targets: in registers (DIRECT: `target`), or in the thread's column of the staged tile
// (`tcol`: direction d at tcol[(d - 1) * TSTRIDE]), read as each push is made.
template <int Q, int KERNEL, int WALL, int INLET, int OUTLET, int RSTRIDE, int TSTRIDE, bool DIRECT>
__device__ __forceinline__ void site_finish(const StepArgs& A, const MrtArgs<Q>& M, const int64_t site, const int64_t first,
                                            const int b, const double (&f)[Q], const uint32_t (&target)[Q],
                                            const uint32_t* __restrict__ tcol, const uint32_t* __restrict__ rec) {
  // CalculatePreCollision (Normal.h:29-33 -> kernel.CalculateDensityMomentumFeq)
  double rho, m[3];
  density_momentum<Q>(f, rho, m);
  double fneq[Q], fpost[Q];
  // (the monitor's smallest population and squared speed are taken now: f dies with the collision, and
  // 1 / rho is at hand)
  double fmin = 0.0, u2 = 0.0;
  {
    const double density_1 = 1. / rho;
    const double mm = m[0] * m[0] + m[1] * m[1] + m[2] * m[2];
#pragma unroll
    for (int d = 0; d < Q; ++d) fneq[d] = f[d] - feq_i<Q>(d, rho, density_1, mm, m);
    if (A.cacheMask & C_MONITOR) {
      fmin = min_population<Q>(f);
      u2 = mm * density_1 * density_1;
    }
  }
  collide<Q, KERNEL>(A, M, f, fneq, fpost);

  // the record has had the whole collision to arrive
  if constexpr (DIRECT) asm volatile("cp.async.wait_all;" ::: "memory");
  uint32_t cut = 0;
  if (b >= 0) {
    if (A.wallOn) cut = rec[0];
    if (A.ioletSel != kIoletNone) cut |= rec[1];
  }
  // stream the uncut links (BulkStreamer.h:31-39); a mid-fluid site has no others
  A.fNew[site] = fpost[0];  // direction 0 streams to the site itself (Domain.cc:449)
#pragma unroll
  for (int d = 1; d < Q; ++d)
    if (!((cut >> (d - 1)) & 1u)) A.fNew[DIRECT ? target[d] : tcol[(d - 1) * TSTRIDE]] = fpost[d];
  if (cut) {
    const uint32_t wallMask = A.wallOn ? rec[0] : 0u;
    const uint32_t ioletMask = A.ioletSel != kIoletNone ? rec[1] : 0u;
    IoletSite S;
    int sel = 0;
    if (ioletMask) {
      const int32_t enc = (int32_t)rec[2];
      sel = A.ioletSel == kIoletByType ? ((enc & kOutletTypedBit) ? 1 : 0) : A.ioletSel;
      const int id = enc & (kOutletTypedBit - 1);
      if constexpr (INLET == OUTLET) iolet_site<Q, INLET>(A, sel, id, b, rho, m, S);
      else if (sel) iolet_site<Q, OUTLET>(A, sel, id, b, rho, m, S);
      else iolet_site<Q, INLET>(A, sel, id, b, rho, m, S);
    }
    cut_links<Q, WALL, INLET, OUTLET, RSTRIDE, 1>(A, site, rec, wallMask, ioletMask, sel, S, rho, fpost);
  }

  if (A.cacheMask) {
    // the monitors need eight values that are live anyway; the moment extraction goes through the call
    if (A.cacheMask & C_MONITOR) fused_monitor(A, DIRECT ? launch_tid_again() : site, fmin, rho, u2);
    if (A.cacheMask & 255u) {
      if constexpr (DIRECT) {
        const int64_t tid = launch_tid_again();
        site_tail<Q, KERNEL>(A, A.siteList ? (int64_t)A.siteList[tid] : first + tid, b);
      } else {
        site_tail<Q, KERNEL>(A, site, b);
      }
    }
  }
}

// The direct form of the site kernel: `count` sites, siteList[0 .. count) or the consecutive device
// sites from `first`, one thread each, every load issued by the thread itself.  What a boundary-typed
// site needs beyond its populations -- masks, iolet id, cut distances: its bRec record -- is fetched
// asynchronously (cp.async, 16 B chunks) into the thread's column of `srec` as soon as the bitmap
// word says the site is boundary-typed: no load depends on another one except through that word
// (small enough to stay in L2).  The product path: whole parts (with the targets as runs where
// StepArgs::runFlags says so), sub-ranges and site lists.
template <int Q, int KERNEL, int WALL, int INLET, int OUTLET>
__global__ void __launch_bounds__(site_threads<Q>(), site_min_ctas<Q>()) collide_stream_kernel(const __grid_constant__ StepArgs A, const __grid_constant__ MrtArgs<Q> M, int64_t first, int64_t count) {
  constexpr int T = site_threads<Q>(), RC = brec_words<Q>() / 4;
  __shared__ uint4 srec[RC * T];
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  // The planes of f_old and of the push targets that a CTA further down the grid will read: one
  // 128 B line per thread and round, into L2.  The loads of this kernel are then mostly L2 hits, and
  // DRAM is kept busy by requests that do not wait for a warp to come round to its load phase.
  constexpr int fLines = Q * (T * 8 / 128), nLines = (Q - 1) * (T * 4 / 128);
  constexpr int rLines = (T / 32 * (Q - 1) * 8 + 127) / 128 + 1;  // the CTA's run words, wherever their lines start
  const int64_t site0 = first + (int64_t)blockIdx.x * T + A.prefetchSites;
  const bool ahead = A.prefetchSites && site0 + T <= first + count;
  uint32_t aheadFlags = 0;
  if (ahead) {
    // (which of the CTA's T / 32 <= 2 groups over there are in runs: asked for now, looked at once this
    // thread's own loads are on their way)
    if (A.runFlags) aheadFlags = __ldg(A.runFlags + (site0 >> 10));
#pragma unroll
    for (int l = threadIdx.x; l < fLines; l += T) {
      const char* p = (const char*)(A.fOld + (int64_t)(l / (T * 8 / 128)) * A.stride + site0) + (l % (T * 8 / 128)) * 128;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
    }
  }
  if (tid >= count) return;
  const int64_t site = A.siteList ? (int64_t)A.siteList[tid] : first + tid;
  // is this a boundary-typed site?  One 8 B word per 32 sites, asked for first.
  const uint2 bi = __ldg(A.bInfo + (site >> 5));
  uint32_t flags = 0;
  if (A.runFlags) flags = __ldg(A.runFlags + (site >> 10));
  double f[Q];
#pragma unroll
  for (int d = 0; d < Q; ++d) f[d] = __ldcs(A.fOld + (int64_t)d * A.stride + site);
  uint32_t target[Q];
  target[0] = (uint32_t)site;
  const unsigned lane = (unsigned)site & 31u;
  if (A.runFlags) {
    // the group's run words are asked for whether or not they stand for its targets (no load waits for the
    // flag); a group that is not in runs (rare) goes to the index planes once the flag says so
    const uint2* __restrict__ run = A.nbrRuns + (site >> 5) * (Q - 1);
    uint2 r[Q];
#pragma unroll
    for (int d = 1; d < Q; ++d) r[d] = __ldg(run + (d - 1));
    if (ahead) {
      const uint32_t runBits = aheadFlags >> ((site0 >> 5) & 31);  // (site0 is a multiple of 64)
#pragma unroll
      for (int l = threadIdx.x; l < nLines + rLines; l += T) {
        const char* p;
        if (l < nLines) {
          // (a 128 B line of a target plane = the 32 sites of one group: not asked for where the runs replace it)
          if ((runBits >> (l % (T * 4 / 128))) & 1u) continue;
          p = (const char*)(A.nbr + (int64_t)(l / (T * 4 / 128)) * A.stride + site0) + (l % (T * 4 / 128)) * 128;
        } else {
          p = (const char*)(A.nbrRuns + (site0 >> 5) * (Q - 1)) + (l - nLines) * 128;
        }
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
      }
    }
    if ((flags >> ((site >> 5) & 31)) & 1u) {
#pragma unroll
      for (int d = 1; d < Q; ++d)
        target[d] = r[d].x + lane + (lane >= (r[d].y & 31u) ? (uint32_t)((int32_t)r[d].y >> 5) : 0u);
    } else {
#pragma unroll
      for (int d = 1; d < Q; ++d) target[d] = __ldcs(A.nbr + (int64_t)(d - 1) * A.stride + site);
    }
  } else {
#pragma unroll
    for (int d = 1; d < Q; ++d) target[d] = __ldcs(A.nbr + (int64_t)(d - 1) * A.stride + site);
    if (ahead) {
#pragma unroll
      for (int l = threadIdx.x; l < nLines; l += T) {
        const char* p = (const char*)(A.nbr + (int64_t)(l / (T * 4 / 128)) * A.stride + site0) + (l % (T * 4 / 128)) * 128;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
      }
    }
  }
  int b = -1;
  {
    if ((bi.x >> lane) & 1u) {
      b = (int)(bi.y + __popc(bi.x & ((1u << lane) - 1u)));
      const uint4* src = A.bRec + (int64_t)b * RC;
      const unsigned dst = (unsigned)__cvta_generic_to_shared(srec + threadIdx.x);
#pragma unroll
      for (int c = 0; c < RC; ++c)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (unsigned)(c * T * sizeof(uint4))), "l"(src + c)
                     : "memory");
    }
  }
  site_finish<Q, KERNEL, WALL, INLET, OUTLET, T, 0, true>(A, M, site, first, b, f, target, nullptr,
                                                        reinterpret_cast<const uint32_t*>(srec + threadIdx.x));
}

// ---------------------------------------------------------------------------------- TMA-staged form
// OPT-IN (HLB_TMA=1), NOT the product path: bit-identical, measured slower than the direct form with its
// L2 prefetch (8 860 - 12 110 against 15 237 MLUPS on the tree, profiles/r02_tma_experiments.md).  What
// follows is the reasoning it was built on.
// The site kernel for a whole part whose first site is tile-aligned (the mid-domain part): a
// persistent kernel, one CTA per SM, whose warps work independently of one another.  A warp walks
// its share of the part's tiles of 32 consecutive sites with two shared-memory stages of its own.
// Everything a tile needs is brought into a stage by the TMA unit -- the Q x 32 box of f_old and the
// (Q-1) x 32 box of push targets (one bulk tensor copy each: the planes are the rows of a 2-D tensor
// map), and the contiguous run of boundary records of the tile's boundary-typed sites (a 1-D bulk
// copy) -- with completion counted on the stage's mbarrier.  While the warp works on the tile in one
// stage (populations to registers; targets, masks and cut distances read from the stage where they
// are needed) its next tile is landing in the other; when it is through, lane 0 refills the stage
// with the tile after next.  No load is waited for by an instruction stream that could be issuing
// other loads: how many bytes are on their way from HBM does not depend on how many warps fit the
// register file or on where in its stream a warp is (the direct form holds 2 x 8 warps x 7.2 KB only
// while those warps are in their load phase; here one 7 KB tile per warp is in flight nearly all
// the time, 12 warps per SM, against the ~43 KB per SM that 6.4 TB/s x 1 us needs), a warp delayed by
// the cut links of its boundary-typed sites delays nobody else, and the register file is left to
// 12 warps of up to 168 registers.
constexpr int kTile = 32;

template <int Q> struct TmaCfg {
  static constexpr int warps = Q > 19 ? 8 : 12;
  static constexpr int threads = warps * 32;
  // registers are per SM sub-partition (16384 each): its share of the CTA's warps has to fit
  static constexpr int perPartition = (warps + 3) / 4;
  static constexpr int maxRegs = (16384 / perPartition / 32) / 8 * 8 > 255 ? 255 : (16384 / perPartition / 32) / 8 * 8;
  static constexpr int fBytes = Q * kTile * 8, nBytes = (Q - 1) * kTile * 4;
  static constexpr int recBytes = brec_words<Q>() * 4;
  // boundary records staged per tile; a tile with more boundary-typed sites (a lattice row running
  // along a wall) reads the rest of its records from global memory
  static constexpr int recCap = Q > 19 ? 16 : 24;
  static constexpr int stageBytes = fBytes + nBytes + (recCap * recBytes + 127) / 128 * 128;
  static constexpr int smemBytes = warps * 2 * stageBytes + warps * 2 * 8 + 128;
};

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  unsigned done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (unsigned)__cvta_generic_to_shared(dst)),
               "l"(src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
               : "memory");
}
// one box of a 2-D tensor map (inner coordinate c0, outer c1) into shared memory
__device__ __forceinline__ void tma_box_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          (unsigned)__cvta_generic_to_shared(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"((unsigned)__cvta_generic_to_shared(bar))
      : "memory");
}

// tile `tile` into stage `st`; b0 / b1: the boundary ordinals at which the tile's run of boundary
// records starts and ends (bInfo[tile].y, bInfo[tile + 1].y)
template <int Q>
__device__ __forceinline__ void tma_fill_stage(const StepArgs& A, const CUtensorMap* mapF, const CUtensorMap* mapN,
                                               unsigned char* st, uint64_t* bar, int64_t tile, unsigned b0, unsigned b1) {
  using C = TmaCfg<Q>;
  const unsigned nrec = b1 - b0 < (unsigned)C::recCap ? b1 - b0 : (unsigned)C::recCap;
  const unsigned recBytes = nrec * C::recBytes;
  mbar_expect_tx(bar, C::fBytes + C::nBytes + recBytes);
  tma_box_2d(st, mapF, (int)(tile * kTile), 0, bar);
  tma_box_2d(st + C::fBytes, mapN, (int)(tile * kTile), 0, bar);
  if (recBytes) bulk_g2s(st + C::fBytes + C::nBytes, A.bRec + (int64_t)b0 * (C::recBytes / 16), recBytes, bar);
}

// sites [0, count) of the device order; nTiles = ceil(count / kTile).  The arrays are padded so that
// whole tiles can be copied (stride is a multiple of 256); bInfo -- one word per tile -- has a
// sentinel word behind the last.
template <int Q, int KERNEL, int WALL, int INLET, int OUTLET>
__global__ void __launch_bounds__(TmaCfg<Q>::threads) __maxnreg__(TmaCfg<Q>::maxRegs) site_tma_kernel(const __grid_constant__ StepArgs A, const __grid_constant__ MrtArgs<Q> M, const __grid_constant__ CUtensorMap mapF, const __grid_constant__ CUtensorMap mapN, int64_t count, int64_t nTiles) {
  using C = TmaCfg<Q>;
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* stage0 = smem + (size_t)warp * 2 * C::stageBytes;
  uint64_t* bar0 = reinterpret_cast<uint64_t*>(smem + (size_t)C::warps * 2 * C::stageBytes) + 2 * warp;
  if (lane == 0) {
    mbar_init(bar0, 1);  // the filler's arrive; the copies complete the byte count
    mbar_init(bar0 + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  // the warp's tiles: first, first + step, ...: concurrent warps work on neighbouring tiles
  const int64_t first = (int64_t)blockIdx.x * C::warps + warp, step = (int64_t)gridDim.x * C::warps;
  // lane 0 carries {bitmap, first boundary ordinal, ordinal behind the last} of the tiles in stage 0 / 1
  uint2 info[2] = {make_uint2(0u, 0u), make_uint2(0u, 0u)};
  unsigned end[2] = {0u, 0u};
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int64_t tile = first + s * step;
      if (tile < nTiles) {
        info[s] = __ldg(A.bInfo + tile);
        end[s] = __ldg(&A.bInfo[tile + 1].y);
        tma_fill_stage<Q>(A, &mapF, &mapN, stage0 + s * C::stageBytes, bar0 + s, tile, info[s].y, end[s]);
      }
    }
  }
  unsigned parity[2] = {0u, 0u};
#pragma unroll 1
  for (int64_t tile = first, k = 0; tile < nTiles; tile += step, ++k) {
    const int s = (int)(k & 1);
    unsigned char* st = stage0 + s * C::stageBytes;
    // what the tile after next needs to be filled with: asked for now, used when this tile is done
    const int64_t refill = tile + 2 * step;
    uint2 nInfo = make_uint2(0u, 0u);
    unsigned nEnd = 0u;
    if (lane == 0 && refill < nTiles) {
      nInfo = __ldg(A.bInfo + refill);
      nEnd = __ldg(&A.bInfo[refill + 1].y);
    }
    const unsigned bits = __shfl_sync(0xffffffffu, s ? info[1].x : info[0].x, 0);
    const unsigned b0 = __shfl_sync(0xffffffffu, s ? info[1].y : info[0].y, 0);
    mbar_wait(bar0 + s, s ? parity[1] : parity[0]);
    if (s) parity[1] ^= 1u; else parity[0] ^= 1u;
    const int64_t site = tile * kTile + lane;
    if (site < count) {
      const double* sf = reinterpret_cast<const double*>(st);
      double f[Q];
#pragma unroll
      for (int d = 0; d < Q; ++d) f[d] = sf[d * kTile + lane];
      uint32_t target[Q];  // (unused: the targets are read from the stage)
      int b = -1;
      const uint32_t* rec = nullptr;
      if ((bits >> lane) & 1u) {
        const unsigned r = __popc(bits & ((1u << lane) - 1u));
        b = (int)(b0 + r);
        rec = r < (unsigned)C::recCap ? reinterpret_cast<const uint32_t*>(st + C::fBytes + C::nBytes) + r * brec_words<Q>()
                                      : reinterpret_cast<const uint32_t*>(A.bRec) + (int64_t)b * brec_words<Q>();
      }
      site_finish<Q, KERNEL, WALL, INLET, OUTLET, 1, kTile, false>(
          A, M, site, 0, b, f, target, reinterpret_cast<const uint32_t*>(st + C::fBytes) + lane, rec);
    }
    // every lane is through with the stage: it takes the tile after next
    __syncwarp();
    if (lane == 0) {
      if (s) { info[1] = nInfo; end[1] = nEnd; } else { info[0] = nInfo; end[0] = nEnd; }
      if (refill < nTiles) tma_fill_stage<Q>(A, &mapF, &mapN, st, bar0 + s, refill, nInfo.y, nEnd);
    }
  }
}

// PostStep of an arbitrary set of boundary-typed sites (sub-range calls), one thread per site; whole
// steps go through the link list (bfl_post_links_kernel, abi.cu)
template <int Q>
__global__ void __launch_bounds__(256) bfl_post_step_kernel(const StepArgs A, int64_t first, int64_t count) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= count) return;
  const int64_t site = A.siteList ? (int64_t)A.siteList[tid] : first + tid;
  const int64_t b = bidx(A, site);
  if (b < 0) return;
  const uint32_t wallMask = A.wallMask[b];
  if (!wallMask) return;
#pragma unroll
  for (int d = 1; d < Q; ++d) {
    if (!((wallMask >> (d - 1)) & 1u)) continue;
    const int id = inv_dir(d);
    if ((wallMask >> (id - 1)) & 1u) continue;
    const double q = (double)A.cutDist[(int64_t)(d - 1) * A.bStride + b];
    if (q < 0.5) {
      double* fi = A.fNew + (int64_t)id * A.stride + site;
      const double fd = A.fNew[(int64_t)d * A.stride + site];
      *fi = 2.0 * q * (*fi) + (1.0 - 2.0 * q) * fd;
    }
  }
}

// Host-side launch entry, one per (Q, KERNEL) translation unit: the site kernel of the
// (wall, inlet, outlet) bundle over `count` sites, followed, for GuoZhengShi walls, by the per-link
// kernel over the boundary-typed sites among them (gzsList / gzsCount, or the same sites when null)
typedef void (*LaunchFn)(int wall, int inlet, int outlet, const StepArgs& A, const void* mrt, int64_t first,
                         int64_t count, const uint32_t* gzsList, int64_t gzsFirst, int64_t gzsCount, void* stream);
template <int Q, int KERNEL>
void launch_collide_stream(int wall, int inlet, int outlet, const StepArgs& A, const void* mrt, int64_t first,
                           int64_t count, const uint32_t* gzsList, int64_t gzsFirst, int64_t gzsCount, void* stream);
// mapF: the current f_old as a 2-D tensor {stride, Q} of doubles with boxes {kTile, Q}; mapN: the push
// targets {stride, Q-1} of uint32 with boxes {kTile, Q-1}
typedef void (*TmaLaunchFn)(int wall, int inlet, int outlet, const StepArgs& A, const void* mrt, const CUtensorMap* mapF,
                            const CUtensorMap* mapN, int64_t count, int nSm, const uint32_t* gzsList, int64_t gzsCount,
                            void* stream);
template <int Q, int KERNEL>
void launch_site_tma(int wall, int inlet, int outlet, const StepArgs& A, const void* mrt, const CUtensorMap* mapF,
                     const CUtensorMap* mapN, int64_t count, int nSm, const uint32_t* gzsList, int64_t gzsCount, void* stream);

}  // namespace hlb
