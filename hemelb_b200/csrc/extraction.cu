// Property extraction (.xtr records) and checkpoint loading on the device (sm_100a).
//
// Device-side counterpart of the reference's extraction path:
//   extraction::LocalPropertyOutput::Write          Code/extraction/LocalPropertyOutput.cc:262-367
//   extraction::LbDataSourceIterator                Code/extraction/LbDataSourceIterator.cc:36-87
//   util::UnitConverter                             Code/util/UnitConverter.{h,cc}
//   extraction::*GeometrySelector::IsWithinGeometry Code/extraction/{Plane,StraightLine,SurfacePoint,...}.cc
//   extraction::LocalDistributionInput              Code/extraction/LocalDistributionInput.cc:107-165
// The reference walks its sites one by one through virtual calls and an XDR stream; here the
// selector is evaluated once for every site on the device (the included-site list and the
// big-endian coordinate triples are kept), and each write is one kernel that gathers the
// property caches / distributions of the included sites, converts them to physical units with
// the reference's operand types and order (-fmad=false), casts to the field's file type, and
// lays the big-endian records out in shared-memory tiles that leave as coalesced 128 B stores.
// Byte-identical to the reference's files (tests/test_gpu_extraction.py).
//
// HBM traffic per included site and write: the record (site_length bytes out), 16 B of
// (site id, coordinates), and the 8 B .. 8Q B of cache / distribution values the fields read.
#include <cuda_runtime.h>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "engine_internal.h"

namespace {

int fail(const std::string& m) { return hlb_internal_fail(m.c_str()); }
#define CU(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess) return fail(std::string(#call) + ": " + cudaGetErrorString(e_));   \
  } while (0)

constexpr int kMaxFields = 16;
constexpr int kTileSites = 128;
constexpr double kMmHgToPascal = 133.3223874;  // Code/constants.h:20
constexpr double kCs2 = 1.0 / 3.0;             // Code/constants.h:41

struct FieldDev {
  int source, typecode, len, byteOffset;
  double offset0;
};

struct Converter {  // util::UnitConverter members (UnitConverter.cc:14-24)
  double latticeDistance, latticeTime, latticeMass, latticeSpeed, latticePressure, referencePressure;
  double origin[3];
};

struct EncodeArgs {
  int64_t first, n;       // slice of the included-site list
  int siteLen, nFields, Q, rank;
  FieldDev fields[kMaxFields];
  const uint32_t* sites;     // included sites (reference ids), ascending
  const uint32_t* coordsBE;  // 3 big-endian words per included site
  const double* f;           // current f_old (SoA)
  int64_t stride, bStride;
  const uint2* bInfo;
  const uint32_t* perm;
  const double* wallNormal;
  const double* cache[8];
  Converter conv;
};

__device__ __forceinline__ uint32_t bswap32(uint32_t x) { return __byte_perm(x, 0, 0x0123); }

// boundary ordinal of a device site (-1: bulk-typed), as the engine's kernels find it
__device__ __forceinline__ int64_t bidx_of(int64_t site, const uint2* __restrict__ bInfo) {
  const uint2 bi = bInfo[site >> 5];
  const unsigned lane = (unsigned)site & 31u;
  return ((bi.x >> lane) & 1u) ? (int64_t)bi.y + __popc(bi.x & ((1u << lane) - 1u)) : -1;
}

// x86-64 conversions of a double to the integer file types (cvttsd2si semantics)
__device__ __forceinline__ int64_t x86_f64_to_i64(double v) {
  if (!(v >= -9223372036854775808.0 && v < 9223372036854775808.0)) return INT64_MIN;
  return (int64_t)v;
}
__device__ __forceinline__ int32_t x86_f64_to_i32(double v) {
  if (!(v > -2147483649.0 && v < 2147483648.0)) return INT32_MIN;
  return (int32_t)v;
}
__device__ __forceinline__ uint64_t x86_f64_to_u64(double v) {
  if (v >= 9223372036854775808.0) return (uint64_t)x86_f64_to_i64(v - 9223372036854775808.0) ^ 0x8000000000000000ull;
  return (uint64_t)x86_f64_to_i64(v);
}

// write FileT(v) big-endian at word w of the record; `isInt`: the C++ value is an int (MpiRank)
__device__ __forceinline__ int put_value(uint32_t* rec, int w, int typecode, double v, bool isInt, int ival) {
  switch (typecode) {
    case 0: {  // float
      float x = isInt ? (float)ival : (float)v;
      uint32_t b = __float_as_uint(x);
      if (x != x) b = 0xFFC00000u;  // x86 "real indefinite": NaNs here are born from invalid operations
      rec[w] = bswap32(b);
      return w + 1;
    }
    case 1: {  // double
      double x = isInt ? (double)ival : v;
      unsigned long long b = (unsigned long long)__double_as_longlong(x);
      if (x != x) b = 0xFFF8000000000000ull;
      rec[w] = bswap32((uint32_t)(b >> 32));
      rec[w + 1] = bswap32((uint32_t)b);
      return w + 2;
    }
    case 2: {
      int32_t x = isInt ? ival : x86_f64_to_i32(v);
      rec[w] = bswap32((uint32_t)x);
      return w + 1;
    }
    case 3: {
      uint32_t x = isInt ? (uint32_t)ival : (uint32_t)x86_f64_to_i64(v);
      rec[w] = bswap32(x);
      return w + 1;
    }
    case 4: {
      int64_t x = isInt ? (int64_t)ival : x86_f64_to_i64(v);
      rec[w] = bswap32((uint32_t)((uint64_t)x >> 32));
      rec[w + 1] = bswap32((uint32_t)(uint64_t)x);
      return w + 2;
    }
    default: {
      uint64_t x = isInt ? (uint64_t)(int64_t)ival : x86_f64_to_u64(v);
      rec[w] = bswap32((uint32_t)(x >> 32));
      rec[w + 1] = bswap32((uint32_t)x);
      return w + 2;
    }
  }
}

__global__ void __launch_bounds__(kTileSites) xtr_encode_kernel(EncodeArgs A, uint32_t* __restrict__ out) {
  extern __shared__ uint32_t tile[];
  const int words = A.siteLen >> 2;
  const int64_t tile0 = (int64_t)blockIdx.x * blockDim.x;
  const int64_t k = tile0 + threadIdx.x;  // ordinal inside this slice
  if (k < A.n) {
    uint32_t* rec = tile + (size_t)threadIdx.x * words;
    const int64_t inc = A.first + k;
    const int64_t s = A.sites[inc];
    rec[0] = A.coordsBE[3 * inc];
    rec[1] = A.coordsBE[3 * inc + 1];
    rec[2] = A.coordsBE[3 * inc + 2];
    int w = 3;
    const Converter& C = A.conv;
    for (int fi = 0; fi < A.nFields; ++fi) {
      const FieldDev F = A.fields[fi];
      switch (F.source) {
        case 0: {  // Pressure: float(ConvertPressureToPhysicalUnits(rho * Cs2)) - double offset
          const double p = A.cache[0][s] * kCs2;
          const double phys = C.referencePressure + ((p - kCs2) * C.latticePressure / kMmHgToPascal);
          const float pf = (float)phys;
          w = put_value(rec, w, F.typecode, (double)pf - F.offset0, false, 0);
          break;
        }
        case 1: {  // Velocity: velocity.as<float>() * float(latticeSpeed)
          const float ls = (float)C.latticeSpeed;
          for (int c = 0; c < 3; ++c) {
            const float vf = (float)A.cache[1][3 * s + c] * ls;
            w = put_value(rec, w, F.typecode, (double)vf, false, 0);
          }
          break;
        }
        case 2:  // ShearStress / VonMisesStress: float(cache * latticePressure)
        case 3: {
          const double v = A.cache[F.source == 2 ? 2 : 3][s] * C.latticePressure;
          w = put_value(rec, w, F.typecode, (double)(float)v, false, 0);
          break;
        }
        case 4: {  // ShearRate: float(cache / latticeTime)
          const double v = A.cache[4][s] / C.latticeTime;
          w = put_value(rec, w, F.typecode, (double)(float)v, false, 0);
          break;
        }
        case 5: {  // StressTensor: latticePressure * M, + ref * mmHg on the diagonal; upper triangle, row-wise
          const double diag = C.referencePressure * kMmHgToPascal;
          const double* m = A.cache[5] + 9 * s;
          const int idx[6] = {0, 1, 2, 4, 5, 8};
          for (int c = 0; c < 6; ++c) {
            double v = C.latticePressure * m[idx[c]];
            if (idx[c] == 0 || idx[c] == 4 || idx[c] == 8) v += diag;
            w = put_value(rec, w, F.typecode, v, false, 0);
          }
          break;
        }
        case 6: {  // Traction: t * latticePressure + (normal * ref) * mmHg; the normal of a site the
                   // geometry file gave none is Vector3D<float>(NO_VALUE) = +inf (Domain.cc:209-211)
          const int64_t i = A.perm ? (int64_t)A.perm[s] : s;
          const int64_t b = bidx_of(i, A.bInfo);
          for (int c = 0; c < 3; ++c) {
            const double nrm = b >= 0 ? A.wallNormal[(int64_t)c * A.bStride + b] : (double)INFINITY;
            double v = A.cache[6][3 * s + c] * C.latticePressure;
            v += nrm * C.referencePressure * kMmHgToPascal;
            w = put_value(rec, w, F.typecode, v, false, 0);
          }
          break;
        }
        case 7: {
          for (int c = 0; c < 3; ++c)
            w = put_value(rec, w, F.typecode, A.cache[7][3 * s + c] * C.latticePressure, false, 0);
          break;
        }
        case 8: {  // Distributions: f_old of the site
          const int64_t i = A.perm ? (int64_t)A.perm[s] : s;
          // nine independent loads in flight per thread, then their conversions (Q = 15 / 19 / 27)
          for (int d0 = 0; d0 < A.Q; d0 += 9) {
            double v[9];
#pragma unroll
            for (int j = 0; j < 9; ++j)
              v[j] = d0 + j < A.Q ? __ldcs(A.f + (int64_t)(d0 + j) * A.stride + i) : 0.0;
#pragma unroll
            for (int j = 0; j < 9; ++j)
              if (d0 + j < A.Q) w = put_value(rec, w, F.typecode, v[j], false, 0);
          }
          break;
        }
        default:
          w = put_value(rec, w, F.typecode, 0.0, true, A.rank);
      }
    }
  }
  __syncthreads();
  const int64_t nHere = min((int64_t)blockDim.x, A.n - tile0);
  const int64_t tileWords = nHere * words;
  uint32_t* dst = out + tile0 * words;
  for (int64_t i = threadIdx.x; i < tileWords; i += blockDim.x) dst[i] = tile[i];
}

// ---------------------------------------------------------------- selectors (float arithmetic as
// the reference's util::Vector3D<float> expressions; std::inner_product order)
struct SelectorDev {
  int kind;
  float p[7];        // as given
  float normal[3];   // plane: normalised
  float line[3], lineLength;
  float voxelF;      // float(GetVoxelSize())
  float originF[3];  // GetOrigin().as<float>()
  double voxel;
};

__device__ __forceinline__ float dot3(const float* a, const float* b) {
  float acc = 0.f;
  acc = acc + a[0] * b[0];
  acc = acc + a[1] * b[1];
  acc = acc + a[2] * b[2];
  return acc;
}

__global__ void xtr_select_kernel(SelectorDev S, const int32_t* __restrict__ coords, int64_t coordStride, int64_t first,
                                  int64_t n, const uint32_t* __restrict__ perm, const uint32_t* __restrict__ wallMask,
                                  const uint2* __restrict__ bInfo, int32_t* __restrict__ flags) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int64_t s = first + k;
  bool inc = true;
  if (S.kind != 0) {
    bool isWall = false;
    if (S.kind == 1 || S.kind == 4) {
      const int64_t i = perm ? (int64_t)perm[s] : s;
      const int64_t b = bidx_of(i, bInfo);
      isWall = b >= 0 && wallMask[b] != 0;  // SiteData::IsWall, SiteDataBare.cc:92-95
    }
    float x[3];
    for (int c = 0; c < 3; ++c) x[c] = (float)coords[(int64_t)c * coordStride + k] * S.voxelF + S.originF[c];
    if (S.kind == 1) {
      inc = isWall;
    } else if (S.kind == 2) {  // PlaneGeometrySelector.cc:52-74
      float d[3] = {x[0] - S.p[0], x[1] - S.p[1], x[2] - S.p[2]};
      const float perp = dot3(d, S.normal);
      inc = !((double)fabsf(perp) > 0.5 * S.voxel);
      if (inc && S.p[6] > 0.f) {
        float r[3];
        for (int c = 0; c < 3; ++c) r[c] = (x[c] - S.normal[c] * perp) - S.p[c];
        inc = dot3(r, r) <= S.p[6] * S.p[6];
      }
    } else if (S.kind == 3) {  // StraightLineGeometrySelector.cc:34-56
      float d[3] = {x[0] - S.p[0], x[1] - S.p[1], x[2] - S.p[2]};
      const float along = dot3(S.line, d) / S.lineLength;
      inc = !((double)along < 0. || along > S.lineLength);
      if (inc) {
        float q[3];
        for (int c = 0; c < 3; ++c) q[c] = (S.p[c] + S.line[c] * along / S.lineLength) - x[c];
        inc = (double)dot3(q, q) <= (2.0 * 0.5 * 0.5 * S.voxel * S.voxel);
      }
    } else {  // SurfacePointSelector.cc:28-44
      inc = false;
      if (isWall) {
        float d[3] = {x[0] - S.p[0], x[1] - S.p[1], x[2] - S.p[2]};
        const double dist = (double)sqrtf(dot3(d, d)) / S.voxel;
        inc = dist <= (double)sqrtf(3.0f);
      }
    }
  }
  flags[k] = inc ? 1 : 0;
}

__global__ void xtr_compact_kernel(const int32_t* __restrict__ flags, const int32_t* __restrict__ pos,
                                   const int32_t* __restrict__ coords, int64_t coordStride, int64_t first, int64_t n,
                                   int64_t base, uint32_t* __restrict__ sites, uint32_t* __restrict__ coordsBE) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n || !flags[k]) return;
  const int64_t o = base + pos[k];
  sites[o] = (uint32_t)(first + k);
  for (int c = 0; c < 3; ++c) coordsBE[3 * o + c] = bswap32((uint32_t)coords[(int64_t)c * coordStride + k]);
}

__global__ void xtr_coords_planes_kernel(const int64_t* __restrict__ aos, int32_t* __restrict__ planes, int64_t n) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  for (int c = 0; c < 3; ++c) planes[(int64_t)c * n + k] = (int32_t)aos[3 * k + c];
}

// ---------------------------------------------------------------- checkpoint records -> f_old, f_new
// LocalDistributionInput.cc:120-160: each record = 3 x uint32 grid position + Q doubles; the
// position must be the one of local site iSite; f_new = f_old = value.
__global__ void __launch_bounds__(kTileSites) xtr_load_kernel(const uint32_t* __restrict__ chunk, int64_t first, int64_t n, int Q,
                                                               const int32_t* __restrict__ coords, int64_t coordStride,
                                                               const uint32_t* __restrict__ perm, double* __restrict__ f0,
                                                               double* __restrict__ f1, int64_t stride,
                                                               unsigned long long* __restrict__ firstBad) {
  extern __shared__ uint32_t tile[];
  const int words = 3 + 2 * Q;
  const int64_t tile0 = (int64_t)blockIdx.x * blockDim.x;
  const int64_t nHere = min((int64_t)blockDim.x, n - tile0);
  const uint32_t* src = chunk + tile0 * words;
  for (int64_t i = threadIdx.x; i < nHere * words; i += blockDim.x) tile[i] = src[i];
  __syncthreads();
  const int64_t k = tile0 + threadIdx.x;
  if (k >= n) return;
  const uint32_t* rec = tile + (size_t)threadIdx.x * words;
  const int64_t s = first + k;
  bool ok = true;
  for (int c = 0; c < 3; ++c) ok = ok && (bswap32(rec[c]) == (uint32_t)coords[(int64_t)c * coordStride + k]);
  if (!ok) {
    atomicMin(firstBad, (unsigned long long)s);
    return;
  }
  const int64_t i = perm ? (int64_t)perm[s] : s;
  for (int d = 0; d < Q; ++d) {
    const unsigned long long b = ((unsigned long long)bswap32(rec[3 + 2 * d]) << 32) | bswap32(rec[4 + 2 * d]);
    const double v = __longlong_as_double((long long)b);
    f0[(int64_t)d * stride + i] = v;
    f1[(int64_t)d * stride + i] = v;
  }
}

inline unsigned blocks_for(int64_t n, int per = 256) { return (unsigned)((n + per - 1) / per); }

const int64_t kChunk = 1 << 21;  // sites per pass over the coordinates / records

int field_length(int source, int Q) {  // LocalPropertyOutput::GetFieldLength, LocalPropertyOutput.cc:398-431
  switch (source) {
    case 0: return 1;
    case 1: return 3;
    case 2: case 3: case 4: return 1;
    case 5: return 6;
    case 6: case 7: return 3;
    case 8: return Q;
    case 9: return 1;
  }
  return -1;
}
int type_size(int tc) { return (tc == 0 || tc == 2 || tc == 3) ? 4 : 8; }
uint32_t cache_bit_of(int source) {  // PropertyActor::SetRequiredProperties, PropertyActor.cc:22-75
  static const uint32_t bits[10] = {1, 2, 4, 8, 16, 32, 64, 128, 0, 0};
  return bits[source];
}

struct FieldHost {
  std::string name;
  int source, typecode;
  std::vector<double> offsets;
};

void put_u32(std::vector<unsigned char>& b, uint32_t v) {
  for (int k = 3; k >= 0; --k) b.push_back((unsigned char)(v >> (8 * k)));
}
void put_u64(std::vector<unsigned char>& b, uint64_t v) {
  for (int k = 7; k >= 0; --k) b.push_back((unsigned char)(v >> (8 * k)));
}
void put_f64(std::vector<unsigned char>& b, double v) {
  uint64_t u;
  std::memcpy(&u, &v, 8);
  put_u64(b, u);
}
void put_f32(std::vector<unsigned char>& b, float v) {
  uint32_t u;
  std::memcpy(&u, &v, 4);
  put_u32(b, u);
}

// io::formats::extraction::GetStoredLengthOfString / GetFieldHeaderLength (io/formats/extraction.h:34-53)
uint64_t stored_string_length(const std::string& s) {
  uint64_t len = s.size();
  if (len % 4) len += 4 - len % 4;
  return len + 4;
}
uint64_t field_header_length(const FieldHost& f) {
  return stored_string_length(f.name) + 12 + (uint64_t)type_size(f.typecode) * f.offsets.size();
}

int parse_fields(const hlb_xtr_spec* spec, int Q, std::vector<FieldHost>& out, uint64_t* siteLen) {
  if (!spec) return fail("null argument");
  if (spec->n_fields < 0 || spec->n_fields > kMaxFields) return fail("between 0 and 16 fields per extraction file");
  if (spec->n_fields && !spec->fields) return fail("null argument");
  uint64_t len = 12;  // 3 x uint32 grid position, LocalPropertyOutput.cc:147-148
  for (int i = 0; i < spec->n_fields; ++i) {
    const hlb_xtr_field& f = spec->fields[i];
    if (f.source < 0 || f.source > 9) return fail("unknown field source");
    if (f.typecode < 0 || f.typecode > 5) return fail("Invalid type");
    const int n = field_length(f.source, Q);
    if (!(f.n_offsets == 0 || f.n_offsets == 1 || f.n_offsets == (uint32_t)n))
      return fail("Invalid length of offsets array " + std::to_string(f.n_offsets));  // LocalPropertyOutput.cc:151-160
    if (f.n_offsets && !f.offsets) return fail("null argument");
    FieldHost h;
    h.name = f.name ? f.name : "";
    h.source = f.source;
    h.typecode = f.typecode;
    h.offsets.assign(f.offsets, f.offsets + f.n_offsets);
    out.push_back(h);
    len += (uint64_t)n * type_size(f.typecode);
  }
  *siteLen = len;
  return 0;
}

Converter make_converter(const hlb_xtr_spec* s) {  // util/UnitConverter.cc:14-24, same operand order
  Converter c;
  const double voxelSize = s->voxel_size, timeStep = s->time_step;
  c.latticeDistance = voxelSize;
  c.latticeTime = timeStep;
  c.latticeMass = s->fluid_density * voxelSize * voxelSize * voxelSize;
  c.latticeSpeed = voxelSize / c.latticeTime;
  c.latticePressure = c.latticeMass / (c.latticeDistance * c.latticeTime * c.latticeTime);
  c.referencePressure = s->reference_pressure;
  for (int k = 0; k < 3; ++k) c.origin[k] = s->origin[k];
  return c;
}

}  // namespace

struct hlb_xtr_handle {
  hlb_gpu_t h = nullptr;
  hlb_gpu_view V;
  std::vector<FieldHost> fields;
  Converter conv;
  uint64_t siteLen = 0;
  int64_t nIncluded = 0;
  uint32_t* sites = nullptr;
  uint32_t* coordsBE = nullptr;
  void* dev = nullptr;
  size_t devBytes = 0;
  void* pinned = nullptr;
  size_t pinnedBytes = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  float lastKernelMs = 0.f;
};

namespace {

int build_site_list(hlb_xtr_handle* x, const hlb_xtr_spec* spec, const int32_t* devPlanes, const int64_t* hostCoords) {
  const hlb_gpu_view& V = x->V;
  const int64_t N = V.N;
  SelectorDev S;
  std::memset(&S, 0, sizeof(S));
  S.kind = spec->selector;
  const bool normalised = S.kind == HLB_XTR_PLANE_NORMALISED;
  if (normalised) S.kind = HLB_XTR_PLANE;
  if (S.kind < 0 || S.kind > 4) return fail("unknown geometry selector");
  for (int k = 0; k < 7; ++k) S.p[k] = spec->selector_params[k];
  S.voxel = spec->voxel_size;
  S.voxelF = (float)spec->voxel_size;
  for (int k = 0; k < 3; ++k) S.originF[k] = (float)spec->origin[k];
  if (S.kind == 2) {  // normal.GetNormalised(): each component / sqrt(dot)
    const float* nrm = S.p + 3;
    float acc = 0.f;
    for (int k = 0; k < 3; ++k) acc = acc + nrm[k] * nrm[k];
    const float mag = std::sqrt(acc);
    for (int k = 0; k < 3; ++k) S.normal[k] = normalised ? nrm[k] : nrm[k] / mag;
  }
  if (S.kind == 3) {  // lineVector = endpoint2 - endpoint1, lineLength = |lineVector|
    float acc = 0.f;
    for (int k = 0; k < 3; ++k) {
      S.line[k] = S.p[3 + k] - S.p[k];
      acc = acc + S.line[k] * S.line[k];
    }
    S.lineLength = std::sqrt(acc);
  }
  const int64_t C = std::min<int64_t>(kChunk, std::max<int64_t>(N, 1));
  int32_t *flags = nullptr, *pos = nullptr, *planes = nullptr;
  int64_t* aos = nullptr;
  void* scanTmp = nullptr;
  size_t scanBytes = 0;
  CU(cudaMalloc(&flags, sizeof(int32_t) * C));
  CU(cudaMalloc(&pos, sizeof(int32_t) * C));
  cub::DeviceScan::ExclusiveSum(nullptr, scanBytes, flags, pos, (int)C);
  CU(cudaMalloc(&scanTmp, scanBytes));
  if (!devPlanes) {
    CU(cudaMalloc(&planes, sizeof(int32_t) * 3 * C));
    CU(cudaMalloc(&aos, sizeof(int64_t) * 3 * C));
  }
  std::vector<int64_t> chunkCount;
  int rc = 0;
  for (int pass = 0; pass < 2 && !rc; ++pass) {
    int64_t base = 0;
    for (int64_t s0 = 0, ci = 0; s0 < N && !rc; s0 += C, ++ci) {
      const int64_t m = std::min(C, N - s0);
      const int32_t* cp;
      int64_t cstride;
      if (devPlanes) {
        cp = devPlanes + s0;
        cstride = N;
      } else {
        CU(cudaMemcpy(aos, hostCoords + 3 * s0, sizeof(int64_t) * 3 * m, cudaMemcpyHostToDevice));
        xtr_coords_planes_kernel<<<blocks_for(m), 256>>>(aos, planes, m);
        cp = planes;
        cstride = m;
      }
      xtr_select_kernel<<<blocks_for(m), 256>>>(S, cp, cstride, s0, m, V.perm, V.wallMask, V.bInfo, flags);
      cub::DeviceScan::ExclusiveSum(scanTmp, scanBytes, flags, pos, (int)m);
      if (pass == 0) {
        int32_t lastPos = 0, lastFlag = 0;
        CU(cudaMemcpy(&lastPos, pos + m - 1, 4, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(&lastFlag, flags + m - 1, 4, cudaMemcpyDeviceToHost));
        chunkCount.push_back((int64_t)lastPos + lastFlag);
      } else {
        xtr_compact_kernel<<<blocks_for(m), 256>>>(flags, pos, cp, cstride, s0, m, base, x->sites, x->coordsBE);
        base += chunkCount[ci];
      }
      CU(cudaGetLastError());
    }
    if (pass == 0) {
      x->nIncluded = 0;
      for (int64_t c : chunkCount) x->nIncluded += c;
      const int64_t cap = std::max<int64_t>(x->nIncluded, 1);
      CU(cudaMalloc(&x->sites, sizeof(uint32_t) * cap));
      CU(cudaMalloc(&x->coordsBE, sizeof(uint32_t) * 3 * cap));
    }
  }
  CU(cudaDeviceSynchronize());
  cudaFree(flags);
  cudaFree(pos);
  cudaFree(scanTmp);
  if (planes) cudaFree(planes);
  if (aos) cudaFree(aos);
  hlb_gpu_internal_count_launch(x->h, 2 * (int64_t)chunkCount.size() * 3);
  return rc;
}

int create_common(hlb_gpu_t h, const hlb_xtr_spec* spec, const int32_t* devPlanes, const int64_t* hostCoords,
                  hlb_xtr_t* out) {
  if (!h || !spec || !out) return fail("null argument");
  auto x = new hlb_xtr_handle();
  x->h = h;
  if (hlb_gpu_internal_view(h, &x->V)) {
    delete x;
    return 1;
  }
  if (parse_fields(spec, x->V.Q, x->fields, &x->siteLen)) {
    delete x;
    return 1;
  }
  if ((x->siteLen / 4) * kTileSites * 4 > 200 * 1024) {
    delete x;
    return fail("site record too long for one shared-memory tile");
  }
  x->conv = make_converter(spec);
  cudaSetDevice(x->V.device);
  if (build_site_list(x, spec, devPlanes, hostCoords)) {
    hlb_xtr_destroy(x);
    return 1;
  }
  cudaEventCreate(&x->ev0);
  cudaEventCreate(&x->ev1);
  const size_t smem = (size_t)kTileSites * x->siteLen;
  if (smem > 48 * 1024)
    CU(cudaFuncSetAttribute(xtr_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  *out = x;
  return 0;
}

}  // namespace

extern "C" {

int hlb_xtr_create(hlb_gpu_t h, const hlb_xtr_spec* spec, const int64_t* site_coords, hlb_xtr_t* out) {
  if (!site_coords) return fail("null argument");
  return create_common(h, spec, nullptr, site_coords, out);
}

int hlb_xtr_create_from_domain(hlb_gpu_t h, hlb_dom_t d, const hlb_xtr_spec* spec, hlb_xtr_t* out) {
  const int32_t* planes = nullptr;
  int64_t n = 0;
  int dev = 0;
  if (hlb_dom_internal_coords(d, &planes, &n, &dev)) return 1;
  hlb_gpu_view V;
  if (!h) return fail("null argument");
  if (hlb_gpu_internal_view(h, &V)) return 1;
  if (n != V.N || dev != V.device) return fail("domain and engine handle do not describe the same rank");
  return create_common(h, spec, planes, nullptr, out);
}

int hlb_xtr_destroy(hlb_xtr_t x) {
  if (!x) return 0;
  cudaSetDevice(x->V.device);
  if (x->sites) cudaFree(x->sites);
  if (x->coordsBE) cudaFree(x->coordsBE);
  if (x->dev) cudaFree(x->dev);
  if (x->pinned) cudaFreeHost(x->pinned);
  if (x->ev0) cudaEventDestroy(x->ev0);
  if (x->ev1) cudaEventDestroy(x->ev1);
  delete x;
  return 0;
}

int hlb_xtr_sizes(hlb_xtr_t x, uint64_t* local_site_count, uint64_t* site_length, uint64_t* header_length) {
  if (!x) return fail("null argument");
  if (local_site_count) *local_site_count = (uint64_t)x->nIncluded;
  if (site_length) *site_length = x->siteLen;
  if (header_length) {
    uint64_t len = 60;  // io::formats::extraction::MainHeaderLength
    for (auto& f : x->fields) len += field_header_length(f);
    *header_length = len;
  }
  return 0;
}

int hlb_xtr_required_caches(hlb_xtr_t x, uint32_t* cache_mask) {
  if (!x || !cache_mask) return fail("null argument");
  uint32_t m = 0;
  for (auto& f : x->fields) m |= cache_bit_of(f.source);
  *cache_mask = m;
  return 0;
}

int hlb_xtr_header(hlb_xtr_t x, uint64_t global_site_count, void* buf, uint64_t capacity) {
  if (!x || !buf) return fail("null argument");
  std::vector<unsigned char> b;
  uint32_t fieldHeaderLen = 0;
  for (auto& f : x->fields) fieldHeaderLen += (uint32_t)field_header_length(f);
  put_u32(b, 0x686c6221u);  // io::formats::HemeLbMagicNumber
  put_u32(b, 0x78747204u);  // io::formats::extraction::MagicNumber
  put_u32(b, 5u);           // VersionNumber
  put_f64(b, x->conv.latticeDistance);
  for (int k = 0; k < 3; ++k) put_f64(b, x->conv.origin[k]);
  put_u64(b, global_site_count);
  put_u32(b, (uint32_t)x->fields.size());
  put_u32(b, fieldHeaderLen);
  for (auto& f : x->fields) {
    put_u32(b, (uint32_t)f.name.size());  // XDR string: length, bytes, zero padding to 4
    for (char c : f.name) b.push_back((unsigned char)c);
    while (b.size() % 4) b.push_back(0);
    put_u32(b, (uint32_t)field_length(f.source, x->V.Q));
    put_u32(b, (uint32_t)f.typecode);
    put_u32(b, (uint32_t)f.offsets.size());
    for (double o : f.offsets) {
      switch (f.typecode) {
        case 0: put_f32(b, (float)o); break;
        case 1: put_f64(b, o); break;
        case 2: put_u32(b, (uint32_t)(int32_t)o); break;
        case 3: put_u32(b, (uint32_t)o); break;
        case 4: put_u64(b, (uint64_t)(int64_t)o); break;
        default: put_u64(b, (uint64_t)o); break;
      }
    }
  }
  if (b.size() != 60 + (size_t)fieldHeaderLen) return fail("internal: header length mismatch");
  if (capacity < b.size()) return fail("header buffer too small");
  std::memcpy(buf, b.data(), b.size());
  return 0;
}

int hlb_xtr_pinned_buffer(hlb_xtr_t x, uint64_t bytes, void** ptr) {
  if (!x || !ptr) return fail("null argument");
  if (x->pinnedBytes < bytes) {
    if (x->pinned) cudaFreeHost(x->pinned);
    x->pinned = nullptr;
    x->pinnedBytes = 0;
    CU(cudaHostAlloc(&x->pinned, std::max<uint64_t>(bytes, 1), cudaHostAllocDefault));
    x->pinnedBytes = bytes;
  }
  *ptr = x->pinned;
  return 0;
}

int hlb_xtr_encode(hlb_xtr_t x, uint64_t first_site, uint64_t n_sites, void* host_buf, uint64_t capacity) {
  if (!x) return fail("null argument");
  if (first_site + n_sites > (uint64_t)x->nIncluded) return fail("site slice outside the included sites");
  const uint64_t bytes = n_sites * x->siteLen;
  if (n_sites && !host_buf) return fail("null argument");
  if (capacity < bytes) return fail("record buffer too small");
  if (!n_sites) return 0;
  CU(cudaSetDevice(x->V.device));
  if (hlb_gpu_internal_view(x->h, &x->V)) return 1;  // f_old / f_new swap every step; caches appear on demand
  EncodeArgs A;
  std::memset(&A, 0, sizeof(A));
  A.first = (int64_t)first_site;
  A.n = (int64_t)n_sites;
  A.siteLen = (int)x->siteLen;
  A.nFields = (int)x->fields.size();
  A.Q = x->V.Q;
  A.rank = x->V.rank;
  for (size_t i = 0; i < x->fields.size(); ++i) {
    const FieldHost& f = x->fields[i];
    A.fields[i].source = f.source;
    A.fields[i].typecode = f.typecode;
    A.fields[i].len = field_length(f.source, A.Q);
    A.fields[i].offset0 = f.offsets.empty() ? 0.0 : f.offsets[0];
    const uint32_t bit = cache_bit_of(f.source);
    if (bit) {
      int ci = 0;
      while ((1u << ci) != bit) ++ci;
      if (!x->V.cache[ci]) return fail("extraction field needs a property cache that no step has filled yet");
    }
  }
  A.sites = x->sites;
  A.coordsBE = x->coordsBE;
  A.f = x->V.f[0];
  A.stride = x->V.stride;
  A.bInfo = x->V.bInfo;
  A.bStride = x->V.bStride;
  A.perm = x->V.perm;
  A.wallNormal = x->V.wallNormal;
  for (int i = 0; i < 8; ++i) A.cache[i] = x->V.cache[i];
  A.conv = x->conv;
  if (x->devBytes < bytes) {
    if (x->dev) cudaFree(x->dev);
    x->dev = nullptr;
    x->devBytes = 0;
    CU(cudaMalloc(&x->dev, bytes));
    x->devBytes = bytes;
  }
  cudaStream_t st = (cudaStream_t)x->V.computeStream;
  CU(cudaEventRecord(x->ev0, st));
  xtr_encode_kernel<<<blocks_for((int64_t)n_sites, kTileSites), kTileSites, (size_t)kTileSites * x->siteLen, st>>>(
      A, (uint32_t*)x->dev);
  CU(cudaGetLastError());
  CU(cudaEventRecord(x->ev1, st));
  CU(cudaMemcpyAsync(host_buf, x->dev, bytes, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  CU(cudaEventElapsedTime(&x->lastKernelMs, x->ev0, x->ev1));
  hlb_gpu_internal_count_launch(x->h, 1);
  return 0;
}

int hlb_xtr_last_encode_ms(hlb_xtr_t x, float* ms) {
  if (!x || !ms) return fail("null argument");
  *ms = x->lastKernelMs;
  return 0;
}

// extraction::LocalDistributionInput::LoadDistribution, the per-rank part (LocalDistributionInput.cc:107-165):
// `records` = this rank's slice of one time step (without the IO rank's 8-byte time stamp)
static int load_common(hlb_gpu_t h, const void* records, uint64_t n_bytes, const int32_t* devPlanes,
                       const int64_t* hostCoords) {
  if (!h || (!records && n_bytes)) return fail("null argument");
  hlb_gpu_view V;
  if (hlb_gpu_internal_view(h, &V)) return 1;
  const uint64_t siteLen = 12 + 8 * (uint64_t)V.Q;
  // the reference walks the slice record by record and then compares the count (":162-164")
  if (n_bytes % siteLen || n_bytes / siteLen != (uint64_t)V.N)
    return fail("Read " + std::to_string(n_bytes / siteLen) + " sites but expected " + std::to_string(V.N));
  CU(cudaSetDevice(V.device));
  cudaStream_t st = (cudaStream_t)V.computeStream;
  CU(cudaStreamSynchronize(st));
  const int64_t C = std::min<int64_t>(kChunk, std::max<int64_t>(V.N, 1));
  void* dev = nullptr;
  int32_t* planes = nullptr;
  int64_t* aos = nullptr;
  unsigned long long* bad = nullptr;
  CU(cudaMalloc(&dev, C * siteLen));
  CU(cudaMalloc(&bad, 8));
  CU(cudaMemset(bad, 0xff, 8));
  if (!devPlanes) {
    CU(cudaMalloc(&planes, sizeof(int32_t) * 3 * C));
    CU(cudaMalloc(&aos, sizeof(int64_t) * 3 * C));
  }
  const size_t smem = (size_t)kTileSites * siteLen;
  if (smem > 48 * 1024) CU(cudaFuncSetAttribute(xtr_load_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t launches = 0;
  for (int64_t s0 = 0; s0 < V.N; s0 += C) {
    const int64_t m = std::min(C, V.N - s0);
    CU(cudaMemcpy(dev, (const char*)records + (uint64_t)s0 * siteLen, (uint64_t)m * siteLen, cudaMemcpyHostToDevice));
    const int32_t* cp;
    int64_t cstride;
    if (devPlanes) {
      cp = devPlanes + s0;
      cstride = V.N;
    } else {
      CU(cudaMemcpy(aos, hostCoords + 3 * s0, sizeof(int64_t) * 3 * m, cudaMemcpyHostToDevice));
      xtr_coords_planes_kernel<<<blocks_for(m), 256>>>(aos, planes, m);
      cp = planes;
      cstride = m;
      ++launches;
    }
    xtr_load_kernel<<<blocks_for(m, kTileSites), kTileSites, smem>>>((const uint32_t*)dev, s0, m, V.Q, cp, cstride, V.perm,
                                                                    V.f[0], V.f[1], V.stride, bad);
    CU(cudaGetLastError());
    ++launches;
  }
  unsigned long long firstBad = 0;
  CU(cudaMemcpy(&firstBad, bad, 8, cudaMemcpyDeviceToHost));
  cudaFree(dev);
  cudaFree(bad);
  if (planes) cudaFree(planes);
  if (aos) cudaFree(aos);
  hlb_gpu_internal_count_launch(h, launches);
  if (firstBad != ~0ull)
    return fail("Site read at index " + std::to_string(firstBad) +
                " is not the site this rank holds there (grid coordinates differ)");
  return 0;
}

int hlb_gpu_load_distributions(hlb_gpu_t h, const void* records, uint64_t n_bytes, const int64_t* site_coords) {
  if (!site_coords) return fail("null argument");
  return load_common(h, records, n_bytes, nullptr, site_coords);
}

int hlb_gpu_load_distributions_from_domain(hlb_gpu_t h, hlb_dom_t d, const void* records, uint64_t n_bytes) {
  const int32_t* planes = nullptr;
  int64_t n = 0;
  int dev = 0;
  if (hlb_dom_internal_coords(d, &planes, &n, &dev)) return 1;
  hlb_gpu_view V;
  if (!h) return fail("null argument");
  if (hlb_gpu_internal_view(h, &V)) return 1;
  if (n != V.N || dev != V.device) return fail("domain and engine handle do not describe the same rank");
  return load_common(h, records, n_bytes, planes, nullptr);
}

}  // extern "C"
