// Lattice constants and per-site arithmetic for the collide-and-stream kernels (sm_100a).
//
// Mirrors the behaviour of the reference's Lattice<Q,V,W,COMPRESSIBLE> static class
// (Code/lb/lattices/Lattice.h:74-950; velocity sets D3Q15.h:17-35, D3Q19.h:18-36, D3Q27.h:18-44).
// Every expression is written in the operation order of the reference's scalar path and the
// translation unit is compiled with -fmad=false, so results are bit-identical to that path.
#pragma once
#include <cstdint>

namespace hlb {

// Velocity sets as constexpr lookup functions (local constexpr tables fold away under unrolling).
#define HLB_CX27 {0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1}
#define HLB_CY27 {0, 0, 0, 1, -1, 0, 0, 1, -1, -1, 1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, -1, 1, -1, 1}
#define HLB_CZ27 {0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, -1, -1, 1, 1, -1, -1, 1, 1, -1, -1, 1, 1, -1, -1, 1}
#define HLB_CX15 {0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1}
#define HLB_CY15 {0, 0, 0, 1, -1, 0, 0, 1, -1, 1, -1, -1, 1, -1, 1}
#define HLB_CZ15 {0, 0, 0, 0, 0, 1, -1, 1, -1, -1, 1, 1, -1, -1, 1}

template <int Q> struct Lat {
  static_assert(Q == 15 || Q == 19 || Q == 27, "D3Q15, D3Q19 or D3Q27");
  __host__ __device__ static constexpr int cx(int i) {
    if constexpr (Q == 15) { constexpr int t[15] = HLB_CX15; return t[i]; }
    else { constexpr int t[27] = HLB_CX27; return t[i]; }  // D3Q19 = first 19 vectors of D3Q27
  }
  __host__ __device__ static constexpr int cy(int i) {
    if constexpr (Q == 15) { constexpr int t[15] = HLB_CY15; return t[i]; }
    else { constexpr int t[27] = HLB_CY27; return t[i]; }
  }
  __host__ __device__ static constexpr int cz(int i) {
    if constexpr (Q == 15) { constexpr int t[15] = HLB_CZ15; return t[i]; }
    else { constexpr int t[27] = HLB_CZ27; return t[i]; }
  }
  __host__ __device__ static constexpr double W(int i) {
    if constexpr (Q == 15) return i == 0 ? 2.0 / 9.0 : (i < 7 ? 1.0 / 9.0 : 1.0 / 72.0);
    else if constexpr (Q == 19) return i == 0 ? 1.0 / 3.0 : (i < 7 ? 1.0 / 18.0 : 1.0 / 36.0);
    else return i == 0 ? 8.0 / 27.0 : (i < 7 ? 2.0 / 27.0 : (i < 19 ? 1.0 / 54.0 : 1.0 / 216.0));
  }
};

// directions come in +/- pairs: INVERSEDIRECTIONS = {0,2,1,4,3,...} (Lattice.h:49-69)
__host__ __device__ constexpr int inv_dir(int d) { return d == 0 ? 0 : (d & 1 ? d + 1 : d - 1); }

constexpr double kCs2 = 1.0 / 3.0;  // constants.h:41

// c * x for c in {-1,0,1}: exact, matches int->double promotion followed by a multiply
__device__ __forceinline__ double cmul(int c, double x) { return c == 0 ? 0.0 : (c > 0 ? x : -x); }

// CX*a + CY*b + CZ*c evaluated left to right, skipping exact zero terms
template <int Q>
__device__ __forceinline__ double dot_c(int i, double a, double b, double c) {
  const int cx = Lat<Q>::cx(i), cy = Lat<Q>::cy(i), cz = Lat<Q>::cz(i);
  double r = 0.0;
  bool started = false;
  if (cx != 0) { r = cmul(cx, a); started = true; }
  if (cy != 0) { r = started ? r + cmul(cy, b) : cmul(cy, b); started = true; }
  if (cz != 0) { r = started ? r + cmul(cz, c) : cmul(cz, c); }
  return r;
}

// Lattice.h:181-191 (scalar CalculateDensityAndMomentum)
template <int Q>
__device__ __forceinline__ void density_momentum(const double (&f)[Q], double& rho, double (&m)[3]) {
  rho = 0.0;
  m[0] = m[1] = m[2] = 0.0;
#pragma unroll
  for (int i = 0; i < Q; ++i) {
    rho += f[i];
    if (Lat<Q>::cx(i) != 0) m[0] += cmul(Lat<Q>::cx(i), f[i]);
    if (Lat<Q>::cy(i) != 0) m[1] += cmul(Lat<Q>::cy(i), f[i]);
    if (Lat<Q>::cz(i) != 0) m[2] += cmul(Lat<Q>::cz(i), f[i]);
  }
}

// one component of Lattice.h:332-355 (scalar CalculateFeq, compressible)
template <int Q>
__device__ __forceinline__ double feq_i(int i, double rho, double density_1, double mm, const double (&m)[3]) {
  const double mde = dot_c<Q>(i, m[0], m[1], m[2]);
  return Lat<Q>::W(i) * (rho - (3. / 2.) * mm * density_1 + (9. / 2.) * density_1 * mde * mde + 3. * mde);
}

template <int Q>
__device__ __forceinline__ void feq_all(double rho, const double (&m)[3], double (&feq)[Q]) {
  const double density_1 = 1. / rho;
  const double mm = m[0] * m[0] + m[1] * m[1] + m[2] * m[2];
#pragma unroll
  for (int i = 0; i < Q; ++i) feq[i] = feq_i<Q>(i, rho, density_1, mm, m);
}

// ------------------------------------------------------------------ MRT bases
// DHumieresD3Q15MRTBasis.h:44-59, DHumieresD3Q19MRTBasis.h:44-62
#define HLB_M15 {{-2, -1, -1, -1, -1, -1, -1, 1, 1, 1, 1, 1, 1, 1, 1}, \
                 {16, -4, -4, -4, -4, -4, -4, 1, 1, 1, 1, 1, 1, 1, 1}, \
                 {0, -4, 4, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1}, \
                 {0, 0, 0, -4, 4, 0, 0, 1, -1, 1, -1, -1, 1, -1, 1}, \
                 {0, 0, 0, 0, 0, -4, 4, 1, -1, -1, 1, 1, -1, -1, 1}, \
                 {0, 2, 2, -1, -1, -1, -1, 0, 0, 0, 0, 0, 0, 0, 0}, \
                 {0, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 0, 0, 0, 0}, \
                 {0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, -1, -1, -1, -1}, \
                 {0, 0, 0, 0, 0, 0, 0, 1, 1, -1, -1, -1, -1, 1, 1}, \
                 {0, 0, 0, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1}, \
                 {0, 0, 0, 0, 0, 0, 0, 1, -1, -1, 1, -1, 1, 1, -1}}
#define HLB_M19 {{-30, -11, -11, -11, -11, -11, -11, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8}, \
                 {12, -4, -4, -4, -4, -4, -4, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1}, \
                 {0, -4, 4, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0}, \
                 {0, 0, 0, -4, 4, 0, 0, 1, -1, -1, 1, 0, 0, 0, 0, 1, -1, 1, -1}, \
                 {0, 0, 0, 0, 0, -4, 4, 0, 0, 0, 0, 1, -1, -1, 1, 1, -1, -1, 1}, \
                 {0, 2, 2, -1, -1, -1, -1, 1, 1, 1, 1, 1, 1, 1, 1, -2, -2, -2, -2}, \
                 {0, -4, -4, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, -2, -2, -2, -2}, \
                 {0, 0, 0, 1, 1, -1, -1, 1, 1, 1, 1, -1, -1, -1, -1, 0, 0, 0, 0}, \
                 {0, 0, 0, -2, -2, 2, 2, 1, 1, 1, 1, -1, -1, -1, -1, 0, 0, 0, 0}, \
                 {0, 0, 0, 0, 0, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 0, 0, 0, 0}, \
                 {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, -1, -1}, \
                 {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0}, \
                 {0, 0, 0, 0, 0, 0, 0, 1, -1, 1, -1, -1, 1, -1, 1, 0, 0, 0, 0}, \
                 {0, 0, 0, 0, 0, 0, 0, -1, 1, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1}, \
                 {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, -1, -1, 1, -1, 1, 1, -1}}

// number of kinetic moments: NUMMOMENTS + 4 == NUMVECTORS (concepts.h:55-58); no D3Q27 basis exists
template <int Q> __host__ __device__ constexpr int mrt_k() { return Q == 15 ? 11 : (Q == 19 ? 15 : 0); }

template <int Q>
__host__ __device__ constexpr int mrt_m(int k, int d) {
  if constexpr (Q == 15) { constexpr int t[11][15] = HLB_M15; return t[k][d]; }
  else if constexpr (Q == 19) { constexpr int t[15][19] = HLB_M19; return t[k][d]; }
  else return 0;
}

// BASIS_TIMES_BASIS_TRANSPOSED[k] and normalisedReducedMomentBasis[k][d] (MRT.h:139-152)
template <int Q>
__host__ __device__ constexpr double mrt_norm(int k) {
  double n = 0.0;
  for (int d = 0; d < Q; ++d) n += double(mrt_m<Q>(k, d)) * double(mrt_m<Q>(k, d));
  return n;
}
template <int Q>
__host__ __device__ constexpr double mrt_mn(int k, int d) { return double(mrt_m<Q>(k, d)) / mrt_norm<Q>(k); }

}  // namespace hlb
