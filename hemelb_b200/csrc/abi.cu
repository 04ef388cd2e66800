// C ABI of the B200 collide-and-stream engine (see include/hemelb_b200.h for the contract and
// the reference interfaces each entry point replaces).
//
// Device layout per handle (one reference rank = one GPU):
//   f[2]      : Q*stride + 1 + S doubles each.  Population d of site s at d*stride + s (SoA;
//               stride = N rounded up to 64 so every plane is 512 B aligned), the reference's
//               rubbish slot at Q*stride, the per-neighbour halo slices after it in the
//               reference's own order (FieldData.cc:14-25, Domain.cc:404-419), so NCCL sends /
//               receives straight out of / into the arrays with no pack step.
//   nbr       : (Q-1) planes of stride uint32: internal target index of each push.
//   boundary  : masks, iolet ids, float cut distances, normals, coordinates -- only for the
//               wall / inlet / outlet typed sites (two contiguous id ranges), plane-major.
#include <cuda_runtime.h>
#include <cub/cub.cuh>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/hemelb_b200.h"
#include "engine_internal.h"
#include "kernels.cuh"

namespace hlb {
// per-(lattice, kernel) launchers, defined in cs_q*_*.cu
#define HLB_DECL(Q, K)                                                                                       \
  extern template void launch_collide_stream<Q, K>(int, int, int, const StepArgs&, const void*, int64_t, int64_t, \
                                                   const uint32_t*, int64_t, int64_t, void*);            \
  extern template void launch_site_tma<Q, K>(int, int, int, const StepArgs&, const void*, const CUtensorMap*,    \
                                             const CUtensorMap*, int64_t, int, const uint32_t*, int64_t, void*);
HLB_DECL(15, K_LBGK) HLB_DECL(15, K_MRT) HLB_DECL(15, K_TRT)
HLB_DECL(19, K_LBGK) HLB_DECL(19, K_MRT) HLB_DECL(19, K_TRT)
HLB_DECL(27, K_LBGK) HLB_DECL(27, K_TRT)
#undef HLB_DECL
}  // namespace hlb

using namespace hlb;

namespace {

thread_local std::string g_err;
int fail(const std::string& m) {
  g_err = m;
  return 1;
}
#define CU(call)                                                                          \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess)                                                                \
      return fail(std::string(#call) + ": " + cudaGetErrorString(e_));                    \
  } while (0)

// ---------------------------------------------------------------- NCCL through dlopen
typedef struct { char internal[128]; } NcclUniqueId;
typedef void* NcclComm;
struct Nccl {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*Send)(const void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool Load() {
    if (lib) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
    }
    if (!lib) return false;
#define SYM(f, n) f = (decltype(f))dlsym(lib, n)
    SYM(GetUniqueId, "ncclGetUniqueId");
    SYM(CommInitRank, "ncclCommInitRank");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(Send, "ncclSend");
    SYM(Recv, "ncclRecv");
    SYM(GroupStart, "ncclGroupStart");
    SYM(GroupEnd, "ncclGroupEnd");
    SYM(AllReduce, "ncclAllReduce");
    SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    return GetUniqueId && CommInitRank && Send && Recv && GroupStart && GroupEnd;
  }
} g_nccl;
const int kNcclDouble = 8;  // ncclFloat64
const int kNcclMin = 3;     // ncclRedOp_t: ncclSum 0, ncclProd 1, ncclMax 2, ncclMin 3

// ---------------------------------------------------------------- host-side iolet providers
// InOutLetCosine::GetDensity, Code/lb/iolets/InOutLetCosine.cc:26-43
struct IoletHost {
  int kind;
  double densityMean, densityAmp, phase, period, warmUpLength, minimumSimulationDensity;
  double Density(uint64_t time_step) const {
    if (kind == 1) return 1.0;  // InOutLetVelocity::GetDensity
    const double PI = 3.14159265358979323846264338327950288;
    const double w = 2.0 * PI / period;
    const double target = densityMean + densityAmp * std::cos(w * time_step + phase);
    if ((double)time_step >= warmUpLength) return target;
    const double interpolationFactor = ((double)time_step) / ((double)warmUpLength);
    return interpolationFactor * target + (1. - interpolationFactor) * minimumSimulationDensity;
  }
};

struct Neighbour {
  int rank;
  int64_t count, first;  // first = reference index (N*Q + 1 + offset)
};

}  // namespace

struct hlb_gpu_handle {
  hlb_gpu_config cfg;
  int Q = 0;
  int64_t N = 0, stride = 0, S = 0, fLen = 0;
  int64_t mid[6], edge[6], midTotal = 0, midBulk = 0, edgeBulk = 0, NB = 0, bStride = 0;
  cudaStream_t compute = nullptr, comm = nullptr;
  cudaEvent_t evEdge = nullptr, evComm = nullptr, evT0 = nullptr, evT1 = nullptr;
  double* f[2] = {nullptr, nullptr};
  int cur = 0;
  uint32_t* nbr = nullptr;
  uint32_t *wallMask = nullptr, *ioletMask = nullptr;
  int32_t* ioletId = nullptr;
  float* cutDist = nullptr;
  double* wallNormal = nullptr;
  int32_t* coords = nullptr;
  int32_t* gzsNeighbour = nullptr;
  uint32_t* streamIdx = nullptr;
  IoletDev* ioletsDev[2] = {nullptr, nullptr};
  double* ioletDensityDev[2] = {nullptr, nullptr};
  double* ioletDensityPinned[2] = {nullptr, nullptr};
  uint64_t pinnedCursor[2] = {0, 0};
  std::vector<IoletHost> ioletsHost[2];
  std::vector<Neighbour> neighbours;
  double* cache[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  uint32_t cacheMask = 0;
  uint64_t timeStep = 1;  // SimulationState.cc:16
  std::vector<char> mrt;
  LaunchFn launch = nullptr;
  TmaLaunchFn launchTma = nullptr;
  bool useTma = false;  // the mid-domain part through the TMA-staged persistent kernel (HLB_TMA=1; measured slower)
  int prefetchSites = 0;  // direct site kernel over a whole part: L2 prefetch distance in sites (HLB_PREFETCH)
  int nSm = 148;
  CUtensorMap mapF[2], mapN;  // f[0], f[1] and the push targets as 2-D tensors (rows = planes), boxes of one tile
  // ---- the product schedule.  Device order: all sites of the mid-domain part sorted by lattice
  // position whatever their collision type, then all sites of the domain-edge part likewise; one
  // site-kernel launch per part, every site running the streamer of its own type.  Whole-range
  // requests of the phase API (hlb_gpu_stream_and_collide with the range's own streamer) are held
  // until every non-empty range of the part has been asked for and then leave as that one launch;
  // when the caller has announced a whole step (hlb_gpu_request_comms, LBM::RequestComms) the
  // part is launched at its first request and the later requests of the step find it done.
  // Everything else (sub-ranges, a streamer on another type's range, hlb_gpu_set_overlap(0)) runs at
  // once, one launch per request, in the caller's order.
  struct Deferred { uint32_t all = 0, pending = 0, done = 0; };
  Deferred sc[2];   // stream-and-collide: [0] mid-domain ranges, [1] domain-edge ranges (bit = type)
  Deferred post;    // PostStep: bit k of the 12 ranges (BFL only)
  bool schedule = true, scheduleDefault = true, announced = false, inFlush = false;
  // boundary-typed sites in device order
  uint2* bInfo = nullptr;      // per 32 internal sites {bitmap, ordinal of the first}
  uint32_t* bSite = nullptr;   // internal site of boundary ordinal b
  uint4* bRec = nullptr;       // per boundary-typed site: masks, iolet id, cut distances (StepArgs::bRec)
  uint2* nbrRuns = nullptr;    // push targets as runs per 32 sites (StepArgs::nbrRuns); HLB_NBR_RUNS=0: not built
  uint32_t* runFlags = nullptr;
  bool useRuns = true;
  int gzsOverlap = 0;          // GuoZhengShi: per-link kernel on `aux`, this many CTAs per SM, beside the site kernel (HLB_GZS_OVERLAP)
  cudaStream_t aux = nullptr;
  cudaEvent_t evFork = nullptr, evJoin = nullptr;
  int64_t runWords = 0, runWordsTotal = 0;  // 32-site words served by runs / all
  int64_t nbMid = 0;           // boundary-typed sites of the mid-domain part (ordinals [0, nbMid))
  std::vector<int64_t> refOrdToB;  // boundary ordinal in reference order -> device ordinal
  // BFL PostStep links {slot of f_new[site, inv d], slot of f_new[site, d], q}, device site order
  uint32_t *postI = nullptr, *postD = nullptr;
  float* postQ = nullptr;
  int64_t nPost = 0;
  // what hlb_gpu_set_step_scalars was last given
  bool scalarsSet = false;
  uint64_t lastStep = 0;
  uint32_t lastMask = 0;
  std::vector<double> lastDens[2];
  double omegaMinus = 0;
  // host staging of the boundary tables until finalise
  std::vector<uint32_t> hWall, hIolet;
  std::vector<int32_t> hIoletId, hCoords;
  std::vector<float> hCut;
  std::vector<double> hNormal;
  bool haveNbr = false, haveSiteData = false, haveCut = false, haveNormal = false, haveCoords = false,
       haveNeighbours = false, haveStream = false, haveIolets[2] = {false, false}, finalised = false;
  bool edgePending = false, commPosted = false, haloProvided = false;
  NcclComm comm_nccl = nullptr;
  void* staging = nullptr;
  size_t stagingBytes = 0;
  uint32_t* siteListDev = nullptr;
  int64_t siteListCap = 0;
  double* monitorDev = nullptr;
  double* monitorPinned = nullptr;
  double* monitorPinnedAsync = nullptr;  // hlb_gpu_monitor_begin / _end
  cudaEvent_t evMonitor = nullptr;
  int monitorPending = 0;                // 1: copy in flight (evMonitor), 2: values ready
  unsigned long long* monitorSlots = nullptr;
  bool monitorFused = false;  // the collide kernels of the current step(s) feed the slots
  int64_t monitorLaunches = 0; // collide launches that fed the slots since the last fold
  int64_t launches = 0;
  // internal renumbering: sites of each of the 12 ranges sorted into long z-runs
  uint32_t* perm = nullptr;    // reference site -> internal site (null = identity)
  uint32_t* iperm = nullptr;   // internal site -> reference site
  int32_t* coordsAll = nullptr;  // 3 planes of stride, reference order, until finalise
  int64_t coordsCovered = 0;
  int64_t rangeFirst[13] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  // GZS site halo (NeighbouringDataManager): whole f_old rows of remote sites, once per step
  struct GzsPeer { int rank; int64_t first, count; };
  std::vector<int64_t> gzsNeedSite, gzsNeedOwnerSite;   // as given (reference ids)
  std::vector<int32_t> gzsNeedDir, gzsNeedOwner;
  std::vector<int64_t> gzsServeSite;
  std::vector<int32_t> gzsServeRank;
  std::vector<GzsPeer> gzsRecvPeers, gzsSendPeers;
  double* gzsGhost = nullptr;      // nNeed rows of Q
  double* gzsSendBuf = nullptr;    // nServe rows of Q
  uint32_t* gzsServeDev = nullptr; // internal site ids to pack
  int64_t nGzsNeed = 0, nGzsServe = 0;
  cudaEvent_t evGzsPack = nullptr, evGzsDone = nullptr;
  bool gzsGhostProvided = false;
  bool profileBulk = false;
  std::vector<cudaEvent_t> profEv;  // pairs around the mid-domain site-kernel launches
  size_t profUsed = 0;
  int64_t profSites = 0;
};

namespace {

const int64_t kChunkSites = 1 << 20;
const int kPinnedSlots = 256;

int ensure_staging(hlb_gpu_t h, size_t bytes) {
  if (h->stagingBytes >= bytes) return 0;
  if (h->staging) cudaFree(h->staging);
  h->staging = nullptr;
  h->stagingBytes = 0;
  CU(cudaMalloc(&h->staging, bytes));
  h->stagingBytes = bytes;
  return 0;
}

int flush_deferred(hlb_gpu_t h);

// f_old changed (swap, upload, initial condition): nothing of the current step has run yet
void new_step_state(hlb_gpu_t h) {
  h->sc[0].done = h->sc[1].done = h->post.done = 0;
  h->announced = false;
}

// every ordering point (copy, swap, read-back, scalar change): whatever was held back leaves now
int join_aux(hlb_gpu_t h) { return flush_deferred(h); }

// boundary ordinal, in REFERENCE site order, of a reference site id; -1 for bulk-typed sites
inline int64_t host_bidx(const hlb_gpu_handle* h, int64_t site) {
  if (site < h->midBulk) return -1;
  if (site < h->midTotal) return site - h->midBulk;
  if (site < h->midTotal + h->edgeBulk) return -1;
  return site - h->midTotal - h->edgeBulk + (h->midTotal - h->midBulk);
}

__global__ void convert_nbr_kernel(const int64_t* __restrict__ aos, uint32_t* __restrict__ nbr, int64_t first,
                                   int64_t n, int Q, int64_t N, int64_t stride, int64_t S, int* bad) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= n * Q) return;
  const int d = (int)(tid / n);
  const int64_t s = tid % n;
  const int64_t v = aos[s * Q + d];
  const int64_t site = first + s;
  if (d == 0) {
    if (v != site * Q) atomicExch(bad, 1);  // Domain.cc:449: direction 0 streams to itself
    return;
  }
  int64_t internal;
  if (v < 0 || v > N * Q + S) {  // local slots, the rubbish slot N*Q, the S halo slots behind it
    atomicExch(bad, 2);
    return;
  }
  if (v < N * Q) internal = (v % Q) * stride + v / Q;
  else internal = v - N * Q + (int64_t)Q * stride;
  nbr[(int64_t)(d - 1) * stride + site] = (uint32_t)internal;
}

__global__ void nbr_to_ref_kernel(const uint32_t* __restrict__ nbr, int64_t* __restrict__ aos, int64_t first, int64_t n,
                                  int Q, int64_t N, int64_t stride, const uint32_t* __restrict__ perm,
                                  const uint32_t* __restrict__ iperm) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= n * Q) return;
  const int d = (int)(tid / n);
  const int64_t s = tid % n;
  const int64_t site = first + s;  // reference id
  int64_t v;
  if (d == 0) v = site * Q;
  else {
    const int64_t isite = perm ? (int64_t)perm[site] : site;
    const int64_t in = nbr[(int64_t)(d - 1) * stride + isite];
    if (in < (int64_t)Q * stride) {
      const int64_t t = in % stride;
      v = (iperm ? (int64_t)iperm[t] : t) * Q + in / stride;
    } else v = in - (int64_t)Q * stride + N * Q;
  }
  aos[s * Q + d] = v;
}

// ---------------------------------------------------------------- internal renumbering
__global__ void coords_to_planes_kernel(const int64_t* __restrict__ aos, int32_t* __restrict__ planes, int64_t first,
                                        int64_t n, int64_t stride) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= n * 3) return;
  const int k = (int)(tid / n);
  const int64_t s = tid % n;
  planes[(int64_t)k * stride + first + s] = (int32_t)aos[s * 3 + k];
}
// sort key: (part, x, y, z) -- sites stay inside their part (mid-domain / domain-edge) and become
// long z-runs whatever their collision type, so the pushes of a warp land on consecutive addresses
__global__ void sort_keys_kernel(const int32_t* __restrict__ coords, int64_t stride, int64_t N, int64_t midTotal,
                                 int lox, int loy, int loz, int64_t Ly, int64_t Lz, uint64_t* __restrict__ keys,
                                 uint32_t* __restrict__ vals) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= N) return;
  const uint64_t part = s >= midTotal ? 1u : 0u;
  const int64_t x = coords[s] - lox, y = coords[stride + s] - loy, z = coords[2 * stride + s] - loz;
  keys[s] = (part << 58) | (uint64_t)((x * Ly + y) * Lz + z);
  vals[s] = (uint32_t)s;
}
// which device sites are boundary-typed: one bitmap word per 32 sites
__global__ void boundary_bits_kernel(const uint32_t* __restrict__ iperm, int64_t N, int64_t nWords, int64_t midBulk,
                                     int64_t midTotal, int64_t edgeBulk, uint32_t* __restrict__ bits,
                                     uint32_t* __restrict__ counts) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if ((s >> 5) >= nWords) return;
  bool isB = false;
  if (s < N) {
    const int64_t ref = iperm ? (int64_t)iperm[s] : s;
    isB = (ref >= midBulk && ref < midTotal) || ref >= midTotal + edgeBulk;
  }
  const unsigned w = __ballot_sync(0xffffffffu, isB);
  if ((threadIdx.x & 31) == 0) {
    bits[s >> 5] = w;
    counts[s >> 5] = __popc(w);
  }
}
__global__ void boundary_info_kernel(const uint32_t* __restrict__ bits, const uint32_t* __restrict__ base,
                                     int64_t nWords, uint2* __restrict__ info, uint32_t* __restrict__ bSite) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t w = s >> 5;
  if (w >= nWords) return;
  const unsigned lane = threadIdx.x & 31;
  const uint32_t bm = bits[w], b0 = base[w];
  if (lane == 0) info[w] = make_uint2(bm, b0);
  if ((bm >> lane) & 1u) bSite[b0 + __popc(bm & ((1u << lane) - 1u))] = (uint32_t)s;
}
// device ordinal of every boundary-typed site, listed in reference order
__global__ void ref_ordinal_kernel(const uint32_t* __restrict__ perm, const uint2* __restrict__ info, int64_t NB,
                                   int64_t nbMid, int64_t midBulk, int64_t midTotal, int64_t edgeBulk,
                                   int32_t* __restrict__ out) {
  const int64_t rb = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (rb >= NB) return;
  const int64_t ref = rb < nbMid ? rb + midBulk : rb - nbMid + midTotal + edgeBulk;
  const int64_t s = perm ? (int64_t)perm[ref] : ref;
  const uint2 bi = info[s >> 5];
  const unsigned lane = (unsigned)s & 31u;
  out[rb] = ((bi.x >> lane) & 1u) ? (int32_t)(bi.y + __popc(bi.x & ((1u << lane) - 1u))) : -1;
}
// PostStep: only BFL does work (BouzidiFirdaousLallemand.h:72-91), after all streaming and the
// halo unpack.  Which links it touches is fixed by the geometry: they are listed once
// (hlb_gpu_finalise) as {slot of f_new[site, inv d], slot of f_new[site, d], q}, in site order, and
// one thread corrects one link -- two loads and a store, no mask -> distance -> value chain.  The
// links are independent of one another (the corrected slot belongs to a direction without a wall
// link, the slot read to one with).
__global__ void __launch_bounds__(256) bfl_post_links_kernel(double* __restrict__ fNew, const uint32_t* __restrict__ slotI,
                                                             const uint32_t* __restrict__ slotD,
                                                             const float* __restrict__ cut, int64_t n) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const double q = (double)cut[k];
  const uint32_t i = slotI[k];
  const double fd = fNew[slotD[k]];
  fNew[i] = 2.0 * q * fNew[i] + (1.0 - 2.0 * q) * fd;
}

__global__ void iota_kernel(uint32_t* __restrict__ v, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = (uint32_t)i;
}
__global__ void gather_links_kernel(const uint32_t* __restrict__ order, const uint32_t* __restrict__ dIn,
                                    const float* __restrict__ qIn, uint32_t* __restrict__ dOut, float* __restrict__ qOut,
                                    int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  dOut[i] = dIn[order[i]];
  qOut[i] = qIn[order[i]];
}
// the site kernel's per-site record (kernels.cuh, StepArgs::bRec) from the plane-major tables
__global__ void boundary_records_kernel(const uint32_t* __restrict__ wallMask, const uint32_t* __restrict__ ioletMask,
                                        const int32_t* __restrict__ ioletId, const float* __restrict__ cut,
                                        int64_t bStride, int64_t NB, int Q, int words, uint32_t* __restrict__ rec) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= NB * words) return;
  const int64_t b = tid / words;
  const int w = (int)(tid % words);
  uint32_t v = 0;
  if (w == 0) v = wallMask[b];
  else if (w == 1) v = ioletMask[b];
  else if (w == 2) v = (uint32_t)ioletId[b];
  else if (w >= 4 && w < 4 + Q - 1) v = __float_as_uint(cut[(int64_t)(w - 4) * bStride + b]);
  rec[tid] = v;
}
// The push targets of 32 consecutive device sites as runs (StepArgs::nbrRuns): one warp per word of
// sites.  Along a lattice row the sites and their neighbours in a direction are both consecutive, so
// target - lane is constant; a word that covers the end of one row and the start of the next has two
// such values.  Links the whole-part launch does not push (cut by a wall or an iolet: bRec masks) do
// not count.  A word whose every direction fits gets its bit in `flags`.
__global__ void nbr_runs_kernel(const uint32_t* __restrict__ nbr, int64_t stride, int64_t N, int Q,
                                const uint2* __restrict__ bInfo, const uint4* __restrict__ bRec, int recChunks,
                                uint2* __restrict__ runs, uint32_t* __restrict__ flags, int64_t nWords) {
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= nWords) return;
  const unsigned lane = threadIdx.x & 31u;
  const int64_t site = w * 32 + lane;
  const bool valid = site < N;
  uint32_t cutMask = 0;
  if (valid) {
    const uint2 bi = bInfo[w];
    if ((bi.x >> lane) & 1u) {
      const uint4 r0 = bRec[(int64_t)(bi.y + __popc(bi.x & ((1u << lane) - 1u))) * recChunks];
      cutMask = r0.x | r0.y;
    }
  }
  bool ok = true;
  for (int d = 1; d < Q; ++d) {
    const uint32_t t = valid ? nbr[(int64_t)(d - 1) * stride + site] : 0u;
    const bool live = valid && !((cutMask >> (d - 1)) & 1u);
    const uint32_t delta = t - lane;
    const unsigned liveMask = __ballot_sync(0xffffffffu, live);
    uint32_t base = 0, split = 0;
    int64_t step = 0;
    if (liveMask) {
      base = __shfl_sync(0xffffffffu, delta, __ffs(liveMask) - 1);
      const unsigned diff = __ballot_sync(0xffffffffu, live && delta != base);
      if (diff) {
        split = (uint32_t)(__ffs(diff) - 1);
        const uint32_t second = __shfl_sync(0xffffffffu, delta, (int)split);
        step = (int64_t)second - (int64_t)base;
        if (__ballot_sync(0xffffffffu, live && lane >= split && delta != second)) ok = false;
        if (step < -(1 << 26) || step >= (1 << 26)) ok = false;
      }
    }
    if (lane == 0) runs[w * (Q - 1) + (d - 1)] = make_uint2(base, ok ? (((uint32_t)(int32_t)step) << 5) | split : 0u);
  }
  if (lane == 0 && ok && valid) atomicOr(flags + (w >> 5), 1u << (w & 31));
}
// BFL PostStep links (BouzidiFirdaousLallemand.h:72-91) of boundary site b: wall link d whose
// opposite is not a wall link and whose cut distance is below one half
__global__ void post_links_count_kernel(const uint32_t* __restrict__ wallMask, const float* __restrict__ cut,
                                        int64_t bStride, int64_t NB, int Q, uint32_t* __restrict__ counts) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= NB) return;
  const uint32_t wm = wallMask[b];
  uint32_t n = 0;
  for (int d = 1; d < Q; ++d) {
    if (!((wm >> (d - 1)) & 1u)) continue;
    const int id = inv_dir(d);
    if ((wm >> (id - 1)) & 1u) continue;
    if ((double)cut[(int64_t)(d - 1) * bStride + b] < 0.5) ++n;
  }
  counts[b] = n;
}
__global__ void post_links_fill_kernel(const uint32_t* __restrict__ wallMask, const float* __restrict__ cut,
                                       int64_t bStride, int64_t NB, int Q, int64_t stride,
                                       const uint32_t* __restrict__ bSite, const uint32_t* __restrict__ offset,
                                       uint32_t* __restrict__ slotI, uint32_t* __restrict__ slotD,
                                       float* __restrict__ q) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= NB) return;
  const uint32_t wm = wallMask[b];
  const int64_t site = bSite[b];
  uint32_t k = offset[b];
  for (int d = 1; d < Q; ++d) {
    if (!((wm >> (d - 1)) & 1u)) continue;
    const int id = inv_dir(d);
    if ((wm >> (id - 1)) & 1u) continue;
    const float c = cut[(int64_t)(d - 1) * bStride + b];
    if ((double)c < 0.5) {
      slotI[k] = (uint32_t)((int64_t)id * stride + site);
      slotD[k] = (uint32_t)((int64_t)d * stride + site);
      q[k] = c;
      ++k;
    }
  }
}
__global__ void invert_perm_kernel(const uint32_t* __restrict__ iperm, uint32_t* __restrict__ perm, int64_t N) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < N) perm[iperm[j]] = (uint32_t)j;
}
__global__ void permute_nbr_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                   const uint32_t* __restrict__ perm, int Q, int64_t N, int64_t stride) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= N * (Q - 1)) return;
  const int64_t d1 = tid / N, s = tid % N;
  uint32_t v = in[d1 * stride + s];
  if ((int64_t)v < (int64_t)Q * stride) v = (uint32_t)(((int64_t)v / stride) * stride + perm[(int64_t)v % stride]);
  out[d1 * stride + perm[s]] = v;
}
__global__ void remap_stream_kernel(uint32_t* __restrict__ idx, const uint32_t* __restrict__ perm, int64_t S,
                                    int64_t stride) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S) return;
  const int64_t v = idx[i];
  idx[i] = (uint32_t)((v / stride) * stride + perm[v % stride]);
}
// pack whole f_old rows (site-major) of the sites other ranks' GZS links extrapolate from
__global__ void gzs_pack_kernel(const double* __restrict__ f, const uint32_t* __restrict__ sites, int64_t n, int Q,
                                int64_t stride, double* __restrict__ out) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= n * Q) return;
  const int64_t k = tid / Q;
  const int j = (int)(tid % Q);
  out[tid] = f[(int64_t)j * stride + sites[k]];
}
__global__ void aos_to_soa_kernel(const double* __restrict__ aos, double* __restrict__ f, int64_t first, int64_t n,
                                  int Q, int64_t stride, const uint32_t* __restrict__ perm) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= n * Q) return;
  const int d = (int)(tid / n);
  const int64_t s = tid % n;
  const int64_t site = perm ? (int64_t)perm[first + s] : first + s;
  f[(int64_t)d * stride + site] = aos[s * Q + d];
}
__global__ void soa_to_aos_kernel(const double* __restrict__ f, double* __restrict__ aos, int64_t first, int64_t n,
                                  int Q, int64_t stride, const uint32_t* __restrict__ perm) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= n * Q) return;
  const int d = (int)(tid / n);
  const int64_t s = tid % n;
  const int64_t site = perm ? (int64_t)perm[first + s] : first + s;
  aos[s * Q + d] = f[(int64_t)d * stride + site];
}
__global__ void fill_planes_kernel(double* __restrict__ f0, double* __restrict__ f1, int64_t N, int64_t stride, int Q,
                                   const double* __restrict__ feq) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= N * Q) return;
  const int d = (int)(tid / N);
  const int64_t s = tid % N;
  f0[(int64_t)d * stride + s] = feq[d];
  f1[(int64_t)d * stride + s] = feq[d];
}
// FieldData::CopyReceived, FieldData.cc:41-48
__global__ void copy_received_kernel(double* __restrict__ fNew, const double* __restrict__ fOldShared,
                                     const uint32_t* __restrict__ streamIdx, int64_t S) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= S) return;
  fNew[streamIdx[tid]] = fOldShared[tid];
}
__global__ void gzs_neighbour_kernel(const uint32_t* __restrict__ nbr, int32_t* __restrict__ out, int Q, int64_t stride,
                                     int64_t bStride, int64_t NB, const uint32_t* __restrict__ bSite) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= NB * (Q - 1)) return;
  const int d = (int)(tid / NB) + 1;
  const int64_t b = tid % NB;
  const int64_t site = bSite[b];
  const int64_t in = nbr[(int64_t)(d - 1) * stride + site];
  int32_t v = INT32_MIN;  // not a local fluid neighbour (rubbish / remote)
  if (in < (int64_t)Q * stride) v = (int32_t)(in - (int64_t)d * stride);
  out[(int64_t)(d - 1) * bStride + b] = v;
}

// {min f_old, min rho, max rho, max |u|^2}: stand-alone pass (when the fused monitor was not on)
#define enc mon_enc
#define dec mon_dec
template <int Q>
__global__ void __launch_bounds__(256) monitor_kernel(const double* __restrict__ f, int64_t N, int64_t stride,
                                                      unsigned long long* __restrict__ out) {
  double fmin = 1e300, rmin = 1e300, rmax = -1e300, umax = 0.0;
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < N; s += (int64_t)gridDim.x * blockDim.x) {
    double v[Q];
#pragma unroll
    for (int d = 0; d < Q; ++d) v[d] = __ldcs(f + (int64_t)d * stride + s);
    double rho = 0, mx = 0, my = 0, mz = 0;
#pragma unroll
    for (int d = 0; d < Q; ++d) {
      fmin = fmin < v[d] ? fmin : v[d];
      rho += v[d];
      mx += Lat<Q>::cx(d) * v[d];
      my += Lat<Q>::cy(d) * v[d];
      mz += Lat<Q>::cz(d) * v[d];
    }
    rmin = rmin < rho ? rmin : rho;
    rmax = rmax > rho ? rmax : rho;
    const double u2 = (mx * mx + my * my + mz * mz) / (rho * rho);
    umax = umax > u2 ? umax : u2;
  }
  for (int o = 16; o > 0; o >>= 1) {  // shuffles must stay convergent: fetch, then select
    const double a = __shfl_xor_sync(0xffffffffu, fmin, o), b = __shfl_xor_sync(0xffffffffu, rmin, o);
    const double c = __shfl_xor_sync(0xffffffffu, rmax, o), d = __shfl_xor_sync(0xffffffffu, umax, o);
    fmin = fmin < a ? fmin : a;
    rmin = rmin < b ? rmin : b;
    rmax = rmax > c ? rmax : c;
    umax = umax > d ? umax : d;
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(out + 0, enc(fmin));
    atomicMin(out + 1, enc(rmin));
    atomicMax(out + 2, enc(rmax));
    atomicMax(out + 3, enc(umax));
  }
}
// StabilityTester's site loop (Code/lb/StabilityTester.h:97-141) as one reduction: out[0] = how many
// populations of f_new fail "value > 0.0" (negative, zero or NaN), out[1] = the largest |u_new - u_old|
// of any site as an order-preserving key (ComputeRelativeDifference, :156-180, before the division by
// the reference value; momentum and density by the scalar Lattice::CalculateDensityAndMomentum)
template <int Q>
__global__ void __launch_bounds__(256) stability_kernel(const double* __restrict__ fNew, const double* __restrict__ fOld,
                                                        int64_t N, int64_t stride, int withConvergence,
                                                        unsigned long long* __restrict__ out) {
  unsigned long long bad = 0;
  double worst = 0.0;
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < N; s += (int64_t)gridDim.x * blockDim.x) {
    double v[Q];
#pragma unroll
    for (int d = 0; d < Q; ++d) v[d] = __ldcs(fNew + (int64_t)d * stride + s);
#pragma unroll
    for (int d = 0; d < Q; ++d)
      if (!(v[d] > 0.0)) ++bad;
    if (withConvergence) {
      double rn, mn[3], ro, mo[3];
      density_momentum<Q>(v, rn, mn);
#pragma unroll
      for (int d = 0; d < Q; ++d) v[d] = __ldcs(fOld + (int64_t)d * stride + s);
      density_momentum<Q>(v, ro, mo);
      const double dx = mn[0] / rn - mo[0] / ro, dy = mn[1] / rn - mo[1] / ro, dz = mn[2] / rn - mo[2] / ro;
      double m2 = 0.0;  // std::inner_product from 0 (util/Vector3D.h:584-588)
      m2 += dx * dx;
      m2 += dy * dy;
      m2 += dz * dz;
      const double e = sqrt(m2);
      // a NaN difference never compares greater than the tolerance in the reference either
      worst = e > worst ? e : worst;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long b = __shfl_xor_sync(0xffffffffu, bad, o);
    const double w = __shfl_xor_sync(0xffffffffu, worst, o);
    bad += b;
    worst = worst > w ? worst : w;
  }
  if ((threadIdx.x & 31) == 0) {
    if (bad) atomicAdd(out + 0, bad);
    atomicMax(out + 1, mon_enc(worst));
  }
}
__global__ void stability_decode_kernel(unsigned long long* io) {
  if (threadIdx.x == 0) {
    ((double*)io)[2] = (double)io[0];
    ((double*)io)[3] = mon_dec(io[1]);
  }
}
// the fused monitor's read-out in one launch: fold the spread slots, re-arm them, decode to doubles
// (out[0..3] = min f, min rho, max rho, max |u|)
__global__ void __launch_bounds__(256) monitor_fold_decode_kernel(unsigned long long* __restrict__ slots,
                                                                 double* __restrict__ out) {
  __shared__ unsigned long long sh[4];
  if (threadIdx.x == 0) { sh[0] = ~0ull; sh[1] = ~0ull; sh[2] = 0ull; sh[3] = 0ull; }
  __syncthreads();
  unsigned long long a = ~0ull, b = ~0ull, c = 0ull, d = 0ull;
  for (int i = threadIdx.x; i < kMonitorSlots; i += blockDim.x) {
    unsigned long long* s = slots + 4 * i;
    a = a < s[0] ? a : s[0];
    b = b < s[1] ? b : s[1];
    c = c > s[2] ? c : s[2];
    d = d > s[3] ? d : s[3];
    s[0] = ~0ull; s[1] = ~0ull; s[2] = 0ull; s[3] = 0ull;
  }
  atomicMin(sh + 0, a);
  atomicMin(sh + 1, b);
  atomicMax(sh + 2, c);
  atomicMax(sh + 3, d);
  __syncthreads();
  if (threadIdx.x < 4) {
    double v = dec(sh[threadIdx.x]);
    if (threadIdx.x == 3) v = sqrt(v);
    out[threadIdx.x] = v;
  }
}
__global__ void monitor_arm_kernel(unsigned long long* __restrict__ slots) {
  for (int i = threadIdx.x + blockIdx.x * blockDim.x; i < kMonitorSlots; i += blockDim.x * gridDim.x) {
    slots[4 * i] = ~0ull; slots[4 * i + 1] = ~0ull; slots[4 * i + 2] = 0ull; slots[4 * i + 3] = 0ull;
  }
}
__global__ void monitor_decode_kernel(unsigned long long* io) {
  if (threadIdx.x < 4) {
    double v = dec(io[threadIdx.x]);
    if (threadIdx.x == 3) v = sqrt(v);
    ((double*)io)[4 + threadIdx.x] = v;
  }
}

inline unsigned blocks_for(int64_t n) { return (unsigned)((n + 255) / 256); }

StepArgs make_args(hlb_gpu_t h) {
  StepArgs A;
  std::memset(&A, 0, sizeof(A));
  A.fOld = h->f[h->cur];
  A.fNew = h->f[h->cur ^ 1];
  A.nbr = h->nbr;
  A.stride = h->stride;
  A.wallMask = h->wallMask;
  A.ioletMask = h->ioletMask;
  A.ioletId = h->ioletId;
  A.cutDist = h->cutDist;
  A.wallNormal = h->wallNormal;
  A.coords = h->coords;
  A.bStride = h->bStride;
  A.bInfo = h->bInfo;
  A.bRec = h->bRec;
  A.gzsNeighbour = h->gzsNeighbour;
  A.gzsGhost = h->gzsGhost;
  for (int w = 0; w < 2; ++w) {
    A.iolets[w] = h->ioletsDev[w];
    A.ioletDensity[w] = h->ioletDensityDev[w];
  }
  A.wallOn = 1;
  A.ioletSel = kIoletByType;
  A.timeStep = h->timeStep;
  A.tau = h->cfg.tau;
  A.omega = -1.0 / h->cfg.tau;  // LbmParameters.h:36
  A.stressParameter = (1.0 - 1.0 / (2.0 * h->cfg.tau)) / std::sqrt(2.0);
  A.omegaMinus = h->omegaMinus;
  A.cacheMask = h->cacheMask;
  A.cDensity = h->cache[0];
  A.cVelocity = h->cache[1];
  A.cWss = h->cache[2];
  A.cVonMises = h->cache[3];
  A.cShearRate = h->cache[4];
  A.cStress = h->cache[5];
  A.cTraction = h->cache[6];
  A.cTangTraction = h->cache[7];
  A.monitorSlots = h->monitorSlots;
  A.refSiteOf = h->iperm;
  A.siteList = nullptr;
  return A;
}

int ensure_caches(hlb_gpu_t h, uint32_t mask) {
  static const int per[8] = {1, 3, 1, 1, 1, 9, 3, 3};
  for (int i = 0; i < 8; ++i)
    if ((mask >> i) & 1u)
      if (!h->cache[i]) {
        CU(cudaMalloc(&h->cache[i], sizeof(double) * per[i] * std::max<int64_t>(h->N, 1)));
        // on the engine's own stream: a null-stream memset is not ordered against `compute`
        // (non-blocking stream) and could still be clearing the array while the step writes it
        CU(cudaMemsetAsync(h->cache[i], 0, sizeof(double) * per[i] * std::max<int64_t>(h->N, 1), h->compute));
      }
  return 0;
}

template <int Q>
void fill_mrt(hlb_gpu_t h) {
  h->mrt.assign(sizeof(MrtArgs<Q>), 0);
  MrtArgs<Q>& M = *reinterpret_cast<MrtArgs<Q>*>(h->mrt.data());
  if constexpr (mrt_k<Q>() > 0) {
    const double tau = h->cfg.tau;
    double S[15];
    if (Q == 15) {  // DHumieresD3Q15MRTBasis.cc:10-25
      const double s[11] = {1.6, 1.2, 1.6, 1.6, 1.6, 1.0 / tau, 1.0 / tau, 1.0 / tau, 1.0 / tau, 1.0 / tau, 1.2};
      for (int k = 0; k < 11; ++k) S[k] = s[k];
    } else {  // DHumieresD3Q19MRTBasis.cc:11-37
      const double s[15] = {1.19, 1.4, 1.2, 1.2, 1.2, 1.0 / tau, 1.4, 1.0 / tau, 1.4, 1.0 / tau,
                            1.0 / tau, 1.0 / tau, 1.98, 1.98, 1.98};
      for (int k = 0; k < 15; ++k) S[k] = s[k];
    }
    for (int k = 0; k < mrt_k<Q>(); ++k)
      for (int d = 0; d < Q; ++d) M.SMn[k][d] = S[k] * mrt_mn<Q>(k, d);
  }
}

// Renumber the sites inside each of the 12 ranges by (x, y, z): long z-runs.  Everything indexed by
// site moves with it: the neighbour table (positions and values), the streaming indices of the
// received distributions and the staged boundary tables.  Halo slots and range bounds do not move.
int build_permutation(hlb_gpu_t h) {
  const int Q = h->Q;
  const int64_t N = h->N;
  // bounding box of the coordinates
  int lo[3], hi[3];
  {
    void* tmp = nullptr;
    size_t tmpBytes = 0;
    int* dres = nullptr;
    CU(cudaMalloc(&dres, sizeof(int) * 6));
    cub::DeviceReduce::Min(tmp, tmpBytes, h->coordsAll, dres, (int)N);
    CU(cudaMalloc(&tmp, tmpBytes + 16));
    for (int k = 0; k < 3; ++k) {
      cub::DeviceReduce::Min(tmp, tmpBytes, h->coordsAll + (int64_t)k * h->stride, dres + k, (int)N);
      cub::DeviceReduce::Max(tmp, tmpBytes, h->coordsAll + (int64_t)k * h->stride, dres + 3 + k, (int)N);
    }
    int res[6];
    CU(cudaMemcpy(res, dres, sizeof(res), cudaMemcpyDeviceToHost));
    for (int k = 0; k < 3; ++k) { lo[k] = res[k]; hi[k] = res[3 + k]; }
    cudaFree(tmp);
    cudaFree(dres);
  }
  const int64_t Ly = (int64_t)hi[1] - lo[1] + 1, Lz = (int64_t)hi[2] - lo[2] + 1, Lx = (int64_t)hi[0] - lo[0] + 1;
  if ((double)Lx * (double)Ly * (double)Lz >= 2.8e17) return fail("coordinate bounding box too large to renumber");
  uint64_t *keysIn = nullptr, *keysOut = nullptr;
  uint32_t* valsIn = nullptr;
  CU(cudaMalloc(&keysIn, sizeof(uint64_t) * N));
  CU(cudaMalloc(&keysOut, sizeof(uint64_t) * N));
  CU(cudaMalloc(&valsIn, sizeof(uint32_t) * N));
  CU(cudaMalloc(&h->iperm, sizeof(uint32_t) * N));
  CU(cudaMalloc(&h->perm, sizeof(uint32_t) * N));
  sort_keys_kernel<<<blocks_for(N), 256>>>(h->coordsAll, h->stride, N, h->midTotal, lo[0], lo[1], lo[2], Ly, Lz, keysIn,
                                           valsIn);
  CU(cudaGetLastError());
  {
    void* tmp = nullptr;
    size_t tmpBytes = 0;
    cub::DeviceRadixSort::SortPairs(tmp, tmpBytes, keysIn, keysOut, valsIn, h->iperm, (int)N);
    CU(cudaMalloc(&tmp, tmpBytes + 16));
    cub::DeviceRadixSort::SortPairs(tmp, tmpBytes, keysIn, keysOut, valsIn, h->iperm, (int)N);
    CU(cudaGetLastError());
    CU(cudaDeviceSynchronize());
    cudaFree(tmp);
  }
  cudaFree(keysIn);
  cudaFree(keysOut);
  cudaFree(valsIn);
  invert_perm_kernel<<<blocks_for(N), 256>>>(h->iperm, h->perm, N);
  CU(cudaGetLastError());
  // neighbour table
  uint32_t* nbr2 = nullptr;
  CU(cudaMalloc(&nbr2, sizeof(uint32_t) * (Q - 1) * h->stride));
  CU(cudaMemset(nbr2, 0xff, sizeof(uint32_t) * (Q - 1) * h->stride));
  permute_nbr_kernel<<<blocks_for(N * (Q - 1)), 256>>>(h->nbr, nbr2, h->perm, Q, N, h->stride);
  CU(cudaGetLastError());
  CU(cudaDeviceSynchronize());
  cudaFree(h->nbr);
  h->nbr = nbr2;
  if (h->S) {
    remap_stream_kernel<<<blocks_for(h->S), 256>>>(h->streamIdx, h->perm, h->S, h->stride);
    CU(cudaGetLastError());
  }
  return 0;
}

// The boundary-typed sites in device order: bitmap + running ordinal per 32 sites (bInfo), the site of
// every ordinal (bSite), and where each boundary ordinal of the reference order went (refOrdToB);
// then the staged boundary tables, which arrive in reference order, move to device order.
int build_boundary_order(hlb_gpu_t h) {
  const int Q = h->Q;
  // (one word more than the sites need: the sentinel {0, NB} behind the last word tells the TMA-staged
  // kernel where the last tile's run of boundary records ends)
  const int64_t nWords = h->stride / 32 + 1;
  uint32_t *bits = nullptr, *counts = nullptr, *base = nullptr;
  CU(cudaMalloc(&bits, sizeof(uint32_t) * nWords));
  CU(cudaMalloc(&counts, sizeof(uint32_t) * nWords));
  CU(cudaMalloc(&base, sizeof(uint32_t) * nWords));
  CU(cudaMalloc(&h->bInfo, sizeof(uint2) * nWords));
  CU(cudaMalloc(&h->bSite, sizeof(uint32_t) * std::max<int64_t>(h->NB, 1)));
  boundary_bits_kernel<<<blocks_for(nWords * 32), 256>>>(h->iperm, h->N, nWords, h->midBulk, h->midTotal, h->edgeBulk,
                                                         bits, counts);
  CU(cudaGetLastError());
  {
    void* tmp = nullptr;
    size_t tmpBytes = 0;
    cub::DeviceScan::ExclusiveSum(tmp, tmpBytes, counts, base, (int)nWords);
    CU(cudaMalloc(&tmp, tmpBytes + 16));
    cub::DeviceScan::ExclusiveSum(tmp, tmpBytes, counts, base, (int)nWords);
    CU(cudaGetLastError());
    CU(cudaDeviceSynchronize());
    cudaFree(tmp);
  }
  boundary_info_kernel<<<blocks_for(nWords * 32), 256>>>(bits, base, nWords, h->bInfo, h->bSite);
  CU(cudaGetLastError());
  CU(cudaDeviceSynchronize());
  cudaFree(bits);
  cudaFree(counts);
  cudaFree(base);
  h->nbMid = h->midTotal - h->midBulk;
  h->refOrdToB.assign(h->NB, 0);
  if (h->NB) {
    int32_t* dOrd = nullptr;
    CU(cudaMalloc(&dOrd, sizeof(int32_t) * h->NB));
    ref_ordinal_kernel<<<blocks_for(h->NB), 256>>>(h->perm, h->bInfo, h->NB, h->nbMid, h->midBulk, h->midTotal,
                                                   h->edgeBulk, dOrd);
    CU(cudaGetLastError());
    std::vector<int32_t> ord(h->NB);
    CU(cudaMemcpy(ord.data(), dOrd, sizeof(int32_t) * h->NB, cudaMemcpyDeviceToHost));
    cudaFree(dOrd);
    std::vector<char> seen(h->NB, 0);
    for (int64_t rb = 0; rb < h->NB; ++rb) {
      if (ord[rb] < 0 || ord[rb] >= h->NB || seen[ord[rb]])
        return fail("internal error: the boundary-typed sites did not keep their count under renumbering");
      if ((rb < h->nbMid) != (ord[rb] < h->nbMid)) return fail("internal error: renumbering moved a site out of its part");
      seen[ord[rb]] = 1;
      h->refOrdToB[rb] = ord[rb];
    }
    auto& to = h->refOrdToB;
    auto move = [&](auto& v, int planes) {
      auto old = v;
      for (int k = 0; k < planes; ++k)
        for (int64_t b = 0; b < h->NB; ++b) v[(size_t)k * h->bStride + to[b]] = old[(size_t)k * h->bStride + b];
    };
    move(h->hWall, 1);
    move(h->hIolet, 1);
    move(h->hIoletId, 1);
    move(h->hCut, Q - 1);
    move(h->hNormal, 3);
    move(h->hCoords, 3);
  }
  return 0;
}

// f[0], f[1] and the neighbour table as TMA tensor maps: 2-D, one row per plane (row pitch `stride`),
// boxes of kTile sites x all planes -- what one bulk tensor copy of the TMA-staged site kernel moves
int build_tensor_maps(hlb_gpu_t h) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CU(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) return fail("cuTensorMapEncodeTiled not available from the driver");
    encode = (EncodeFn)fn;
  }
  const cuuint32_t ones[2] = {1, 1};
  for (int i = 0; i < 3; ++i) {
    const bool isF = i < 2;
    const cuuint64_t dims[2] = {(cuuint64_t)h->stride, (cuuint64_t)(isF ? h->Q : h->Q - 1)};
    const cuuint64_t pitch[1] = {(cuuint64_t)h->stride * (isF ? 8u : 4u)};
    const cuuint32_t box[2] = {(cuuint32_t)kTile, (cuuint32_t)(isF ? h->Q : h->Q - 1)};
    const CUresult rc = encode(isF ? &h->mapF[i] : &h->mapN, isF ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_UINT32,
                               2, isF ? (void*)h->f[i] : (void*)h->nbr, dims, pitch, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed (" + std::to_string((int)rc) + ")");
  }
  return 0;
}

// the BFL PostStep link list, once the boundary tables are on the device
int build_post_links(hlb_gpu_t h) {
  h->nPost = 0;
  if (h->cfg.wall != HLB_WALL_BFL || h->NB == 0) return 0;
  uint32_t *counts = nullptr, *offset = nullptr;
  CU(cudaMalloc(&counts, sizeof(uint32_t) * (h->NB + 1)));
  CU(cudaMalloc(&offset, sizeof(uint32_t) * (h->NB + 1)));
  CU(cudaMemset(counts, 0, sizeof(uint32_t) * (h->NB + 1)));
  post_links_count_kernel<<<blocks_for(h->NB), 256>>>(h->wallMask, h->cutDist, h->bStride, h->NB, h->Q, counts);
  CU(cudaGetLastError());
  {
    void* tmp = nullptr;
    size_t tmpBytes = 0;
    cub::DeviceScan::ExclusiveSum(tmp, tmpBytes, counts, offset, (int)(h->NB + 1));
    CU(cudaMalloc(&tmp, tmpBytes + 16));
    cub::DeviceScan::ExclusiveSum(tmp, tmpBytes, counts, offset, (int)(h->NB + 1));
    CU(cudaGetLastError());
    CU(cudaDeviceSynchronize());
    cudaFree(tmp);
  }
  uint32_t total = 0;
  CU(cudaMemcpy(&total, offset + h->NB, sizeof(uint32_t), cudaMemcpyDeviceToHost));
  h->nPost = total;
  if (total) {
    CU(cudaMalloc(&h->postI, sizeof(uint32_t) * total));
    CU(cudaMalloc(&h->postD, sizeof(uint32_t) * total));
    CU(cudaMalloc(&h->postQ, sizeof(float) * total));
    post_links_fill_kernel<<<blocks_for(h->NB), 256>>>(h->wallMask, h->cutDist, h->bStride, h->NB, h->Q, h->stride,
                                                       h->bSite, offset, h->postI, h->postD, h->postQ);
    CU(cudaGetLastError());
    // in the order of the slot corrected: plane by plane, sites ascending -- neighbouring wall sites of a
    // lattice row are cut in the same directions, so consecutive links touch consecutive addresses (the
    // links are independent of one another: any order gives the same result)
    uint32_t *keys = nullptr, *order = nullptr, *order2 = nullptr, *d2 = nullptr;
    float* q2 = nullptr;
    CU(cudaMalloc(&keys, sizeof(uint32_t) * total));
    CU(cudaMalloc(&order, sizeof(uint32_t) * total));
    CU(cudaMalloc(&order2, sizeof(uint32_t) * total));
    CU(cudaMalloc(&d2, sizeof(uint32_t) * total));
    CU(cudaMalloc(&q2, sizeof(float) * total));
    iota_kernel<<<blocks_for(total), 256>>>(order, total);
    void* tmp = nullptr;
    size_t tmpBytes = 0;
    cub::DeviceRadixSort::SortPairs(tmp, tmpBytes, h->postI, keys, order, order2, (int)total);
    CU(cudaMalloc(&tmp, tmpBytes + 16));
    cub::DeviceRadixSort::SortPairs(tmp, tmpBytes, h->postI, keys, order, order2, (int)total);
    CU(cudaGetLastError());
    gather_links_kernel<<<blocks_for(total), 256>>>(order2, h->postD, h->postQ, d2, q2, total);
    CU(cudaGetLastError());
    CU(cudaDeviceSynchronize());
    cudaFree(tmp);
    cudaFree(order);
    cudaFree(order2);
    cudaFree(h->postI);
    cudaFree(h->postD);
    cudaFree(h->postQ);
    h->postI = keys;
    h->postD = d2;
    h->postQ = q2;
  }
  cudaFree(counts);
  cudaFree(offset);
  return 0;
}

int prof_begin(hlb_gpu_t h) {
  while (h->profEv.size() < h->profUsed + 2) {
    cudaEvent_t e;
    CU(cudaEventCreate(&e));
    h->profEv.push_back(e);
  }
  CU(cudaEventRecord(h->profEv[h->profUsed], h->compute));
  return 0;
}
int prof_end(hlb_gpu_t h, int64_t sites) {
  CU(cudaEventRecord(h->profEv[h->profUsed + 1], h->compute));
  h->profUsed += 2;
  h->profSites += sites;
  return 0;
}

// One launch of the site kernel over a whole part (0 mid-domain, 1 domain-edge): every site runs
// the streamer of its own collision type -- what LBM::PreReceive / PreSend ask for with their six
// StreamAndCollide calls (lb.hpp:176-251), which read f_old and write disjoint slots of f_new.
int launch_part(hlb_gpu_t h, int part) {
  const int64_t first = part ? h->midTotal : 0;
  const int64_t count = part ? h->N - h->midTotal : h->midTotal;
  if (count <= 0) return 0;
  StepArgs A = make_args(h);
  const bool prof = h->profileBulk && part == 0;
  if (prof && prof_begin(h)) return 1;
  const int64_t gFirst = part ? h->nbMid : 0, gCount = part ? h->NB - h->nbMid : h->nbMid;
  if (part == 0) A.prefetchSites = h->prefetchSites;  // (line-aligned planes: the part starts at site 0)
  A.nbrRuns = h->nbrRuns;  // every site by its own type, every cut link masked: the runs stand for the target planes
  A.runFlags = h->runFlags;
  if (part == 0 && h->useTma)
    h->launchTma(h->cfg.wall, h->cfg.inlet, h->cfg.outlet, A, h->mrt.data(), &h->mapF[h->cur], &h->mapN, count, h->nSm,
                 h->bSite, gCount, h->compute);
  else if (h->gzsOverlap && h->cfg.wall == HLB_WALL_GZS && gCount > 0) {
    // GuoZhengShi: the per-link kernel (FP64-issue-bound) beside the site kernel (memory-bound).  Both read
    // f_old; the link kernel writes the wall-link populations of the boundary-typed sites, which the site
    // kernel leaves alone.  The link kernel goes first, on a high-priority stream, so that its CTAs take
    // their share of every SM and the site kernel fills the rest.
    CU(cudaEventRecord(h->evFork, h->compute));
    CU(cudaStreamWaitEvent(h->aux, h->evFork, 0));
    StepArgs G = A;
    G.gzsGridLimit = h->nSm * h->gzsOverlap;  // resident from the start: the site kernel's CTAs are dispatched beside them
    h->launch(h->cfg.wall, h->cfg.inlet, h->cfg.outlet, G, h->mrt.data(), first, 0, h->bSite + gFirst, 0, gCount, h->aux);
    CU(cudaEventRecord(h->evJoin, h->aux));
    h->launch(h->cfg.wall, h->cfg.inlet, h->cfg.outlet, A, h->mrt.data(), first, count, h->bSite + gFirst, 0, 0, h->compute);
    CU(cudaStreamWaitEvent(h->compute, h->evJoin, 0));
  } else
    h->launch(h->cfg.wall, h->cfg.inlet, h->cfg.outlet, A, h->mrt.data(), first, count, h->bSite + gFirst, 0, gCount,
              h->compute);
  h->launches++;
  if (h->cfg.wall == HLB_WALL_GZS && gCount > 0) h->launches++;  // the per-link kernel behind the per-site one
  if (h->cacheMask & C_MONITOR) h->monitorLaunches++;
  if (prof && prof_end(h, count)) return 1;
  CU(cudaGetLastError());
  return 0;
}

// every BFL PostStep of the step in one launch over the link list
int launch_post_links(hlb_gpu_t h) {
  if (h->nPost == 0) return 0;
  bfl_post_links_kernel<<<blocks_for(h->nPost), 256, 0, h->compute>>>(h->f[h->cur ^ 1], h->postI, h->postD, h->postQ,
                                                                      h->nPost);
  h->launches++;
  CU(cudaGetLastError());
  return 0;
}

// One streamer on one range of reference site ids, at once: `slot` decides the link policies,
// whatever the types of the sites (as the reference's streamer objects behave when
// StreamerTests.cc calls them on a range of its choosing).
int launch_now(hlb_gpu_t h, int slot, int64_t first, int64_t count, bool post) {
  const bool canWall = (slot == 1 || slot == 4 || slot == 5);
  const bool canIolet = slot >= 2;
  const bool isInletSlot = (slot == 2 || slot == 4);
  StepArgs A = make_args(h);
  A.wallOn = canWall ? 1 : 0;
  A.ioletSel = canIolet ? (isInletSlot ? 0 : 1) : kIoletNone;
  if (h->perm) A.siteList = h->perm + first;  // reference id -> device site, contiguous in reference order
  if (post) {
    // StreamerTypeFactory::PostStep: only the BFL wall link does anything
    if (!canWall || h->cfg.wall != HLB_WALL_BFL) return 0;
    switch (h->Q) {
      case 15: bfl_post_step_kernel<15><<<blocks_for(count), 256, 0, h->compute>>>(A, first, count); break;
      case 19: bfl_post_step_kernel<19><<<blocks_for(count), 256, 0, h->compute>>>(A, first, count); break;
      case 27: bfl_post_step_kernel<27><<<blocks_for(count), 256, 0, h->compute>>>(A, first, count); break;
    }
    h->launches++;
    CU(cudaGetLastError());
    return 0;
  }
  const bool prof = h->profileBulk && slot == 0;
  if (prof && prof_begin(h)) return 1;
  h->launch(h->cfg.wall, h->cfg.inlet, h->cfg.outlet, A, h->mrt.data(), first, count, nullptr, first, count, h->compute);
  h->launches++;
  if (canWall && h->cfg.wall == HLB_WALL_GZS) h->launches++;
  if (h->cacheMask & C_MONITOR) h->monitorLaunches++;
  if (prof && prof_end(h, count)) return 1;
  CU(cudaGetLastError());
  return 0;
}

int flush_sc(hlb_gpu_t h, int part) {
  const uint32_t pending = h->sc[part].pending;
  h->sc[part].pending = 0;
  if (!pending) return 0;
  if (pending == h->sc[part].all) return launch_part(h, part);
  // not every range of the part was asked for: one by one
  int rc = 0;
  for (int t = 0; t < 6 && !rc; ++t)
    if (pending & (1u << t)) {
      const int k = part ? 6 + t : t;
      rc = launch_now(h, t, h->rangeFirst[k], h->rangeFirst[k + 1] - h->rangeFirst[k], false);
    }
  return rc;
}

int flush_post(hlb_gpu_t h) {
  const uint32_t pending = h->post.pending;
  h->post.pending = 0;
  if (!pending) return 0;
  if (pending == h->post.all) return launch_post_links(h);
  int rc = 0;
  // (the reference's order: domain-edge ranges, then mid-domain ranges, lb.hpp:257-309)
  for (int k : {6, 7, 8, 9, 10, 11, 0, 1, 2, 3, 4, 5})
    if (!rc && (pending & (1u << k)))
      rc = launch_now(h, k % 6, h->rangeFirst[k], h->rangeFirst[k + 1] - h->rangeFirst[k], true);
  return rc;
}

int flush_deferred(hlb_gpu_t h) {
  if (flush_sc(h, 1) || flush_sc(h, 0)) return 1;
  return flush_post(h);
}

// a request of the phase API: streamer `slot` over reference sites [first, first + count)
int launch_range(hlb_gpu_t h, int slot, int64_t first, int64_t count, bool post) {
  if (!h->finalised) return fail("handle not finalised");
  if (slot < 0 || slot > 5) return fail("streamer slot out of range");
  if (count <= 0) return 0;
  if (first < 0 || first + count > h->N) return fail("site range outside the local fluid sites");
  if (slot != 0) {
    // boundary streamers read per-site tables that only exist for boundary-typed sites
    if (host_bidx(h, first) < 0 || host_bidx(h, first + count - 1) < 0 ||
        (first < h->midTotal && first + count > h->midTotal))
      return fail("boundary streamer called on bulk-typed sites");
  }
  int whole = -1;
  for (int k = 0; k < 12; ++k)
    if (first == h->rangeFirst[k] && first + count == h->rangeFirst[k + 1]) whole = k;
  if (h->schedule && whole >= 0 && slot == whole % 6) {
    const uint32_t bit = 1u << (whole % 6);
    if (post) {
      const uint32_t pbit = 1u << whole;
      if (flush_sc(h, 1) || flush_sc(h, 0)) return 1;  // PostStep follows all streaming
      if (!(h->post.all & pbit)) return 0;             // nothing to do on this range
      if (h->post.done & pbit) return 0;
      if (h->announced) {
        h->post.done = h->post.all;
        return launch_post_links(h);
      }
      if ((h->post.pending & pbit) && flush_post(h)) return 1;
      h->post.pending |= pbit;
      if (h->post.pending == h->post.all) return flush_post(h);
      return 0;
    }
    const int part = whole / 6;
    hlb_gpu_handle::Deferred& D = h->sc[part];
    if (D.done & bit) return 0;  // ran with the rest of its part earlier in this announced step
    if (h->announced) {
      if (flush_sc(h, part)) return 1;
      D.done = D.all;
      return launch_part(h, part);
    }
    if ((D.pending & bit) && flush_sc(h, part)) return 1;
    D.pending |= bit;
    if (D.pending == D.all) return flush_sc(h, part);
    return 0;
  }
  if (flush_deferred(h)) return 1;
  return launch_now(h, slot, first, count, post);
}

int post_comms(hlb_gpu_t h) {
  // FieldData::SendAndReceive (FieldData.cc:27-39): per neighbour, receive into the slice of
  // f_old and send the same slice of f_new, on the comm stream after the edge ranges finished.
  if (flush_sc(h, 1)) return 1;  // the domain-edge sites write what is sent
  if (h->neighbours.empty()) return 0;
  if (!h->comm_nccl) return 0;  // host-staged exchange: the caller moves the halo (get_halo / set_halo)
  CU(cudaEventRecord(h->evEdge, h->compute));
  CU(cudaStreamWaitEvent(h->comm, h->evEdge, 0));
  double* fOld = h->f[h->cur];
  double* fNew = h->f[h->cur ^ 1];
  const int64_t base = (int64_t)h->Q * h->stride - h->N * h->Q;  // reference index -> internal
  int rc = g_nccl.GroupStart();
  for (auto& nb : h->neighbours) {
    if (rc) break;
    rc = g_nccl.Recv(fOld + nb.first + base, (size_t)nb.count, kNcclDouble, nb.rank, h->comm_nccl, h->comm);
    if (rc) break;
    rc = g_nccl.Send(fNew + nb.first + base, (size_t)nb.count, kNcclDouble, nb.rank, h->comm_nccl, h->comm);
  }
  int rc2 = g_nccl.GroupEnd();
  if (rc || rc2) return fail(std::string("NCCL send/recv: ") + g_nccl.GetErrorString(rc ? rc : rc2));
  CU(cudaEventRecord(h->evComm, h->comm));
  h->commPosted = true;
  return 0;
}

int upload_densities(hlb_gpu_t h, int which, const double* d) {
  const int n = which ? h->cfg.n_outlets : h->cfg.n_inlets;
  if (n == 0) return 0;
  if (join_aux(h)) return 1;  // held-back launches belong to the step whose densities are still set
  if (!d) return fail("iolet densities missing");
  // ring of pinned slots: a slot is reused only after the stream drained (every kPinnedSlots uploads;
  // the H2D copies out of the ring are stream-ordered on `compute`)
  const int slot = (int)(h->pinnedCursor[which]++ % kPinnedSlots);
  if (slot == 0 && h->pinnedCursor[which] > 1) CU(cudaStreamSynchronize(h->compute));
  double* src = h->ioletDensityPinned[which] + (size_t)slot * n;
  std::memcpy(src, d, sizeof(double) * n);
  CU(cudaMemcpyAsync(h->ioletDensityDev[which], src, sizeof(double) * n, cudaMemcpyHostToDevice, h->compute));
  return 0;
}

// NeighbouringDataManager::TransferFieldDependentInformation (phase 0 of the step,
// Code/geometry/neighbouring/NeighbouringDataManager.cc:101-142): ship whole f_old rows of the sites
// that other ranks' GZS links extrapolate from.
int exchange_site_halo(hlb_gpu_t h) {
  if (h->nGzsNeed == 0 && h->nGzsServe == 0) return 0;
  if (join_aux(h)) return 1;
  const int Q = h->Q;
  if (h->nGzsServe) {
    gzs_pack_kernel<<<blocks_for(h->nGzsServe * Q), 256, 0, h->compute>>>(h->f[h->cur], h->gzsServeDev, h->nGzsServe, Q,
                                                                          h->stride, h->gzsSendBuf);
    h->launches++;
    CU(cudaGetLastError());
  }
  if (!h->comm_nccl) return 0;  // host-staged: hlb_gpu_get_gzs_send / hlb_gpu_set_gzs_ghost
  CU(cudaEventRecord(h->evGzsPack, h->compute));
  CU(cudaStreamWaitEvent(h->comm, h->evGzsPack, 0));
  int rc = g_nccl.GroupStart();
  for (auto& p : h->gzsRecvPeers)
    if (!rc) rc = g_nccl.Recv(h->gzsGhost + p.first * Q, (size_t)(p.count * Q), kNcclDouble, p.rank, h->comm_nccl, h->comm);
  for (auto& p : h->gzsSendPeers)
    if (!rc) rc = g_nccl.Send(h->gzsSendBuf + p.first * Q, (size_t)(p.count * Q), kNcclDouble, p.rank, h->comm_nccl, h->comm);
  int rc2 = g_nccl.GroupEnd();
  if (rc || rc2) return fail(std::string("NCCL site halo: ") + g_nccl.GetErrorString(rc ? rc : rc2));
  CU(cudaEventRecord(h->evGzsDone, h->comm));
  CU(cudaStreamWaitEvent(h->compute, h->evGzsDone, 0));
  return 0;
}

int one_step(hlb_gpu_t h) {
  if (exchange_site_halo(h)) return 1;
  h->scalarsSet = false;  // the densities below replace whatever hlb_gpu_set_step_scalars uploaded
  h->sc[0].done = h->sc[1].done = h->post.done = 0;
  // BoundaryValues::GetBoundaryDensity -> iolet->GetDensity(Get0IndexedTimeStep())
  for (int w = 0; w < 2; ++w) {
    const int n = w ? h->cfg.n_outlets : h->cfg.n_inlets;
    if (!n) continue;
    std::vector<double> dens(n);
    for (int i = 0; i < n; ++i) dens[i] = h->ioletsHost[w][i].Density(h->timeStep - 1);
    if (upload_densities(h, w, dens.data())) return 1;
  }
  int64_t off = h->midTotal;
  for (int t = 0; t < 6; ++t) {  // LBM::PreSend, lb.hpp:176-212
    if (launch_range(h, t, off, h->edge[t], false)) return 1;
    off += h->edge[t];
  }
  if (post_comms(h)) return 1;  // RequestComms + Net::Receive/Send
  off = 0;
  for (int t = 0; t < 6; ++t) {  // LBM::PreReceive, lb.hpp:215-251
    if (launch_range(h, t, off, h->mid[t], false)) return 1;
    off += h->mid[t];
  }
  if (hlb_gpu_copy_received(h)) return 1;  // LBM::PostReceive, lb.hpp:254-309
  off = h->midTotal;
  for (int t = 0; t < 6; ++t) {
    if (launch_range(h, t, off, h->edge[t], true)) return 1;
    off += h->edge[t];
  }
  off = 0;
  for (int t = 0; t < 6; ++t) {
    if (launch_range(h, t, off, h->mid[t], true)) return 1;
    off += h->mid[t];
  }
  if (flush_deferred(h)) return 1;
  h->cur ^= 1;    // FieldData::SwapOldAndNew
  h->timeStep++;  // SimulationState::Increment
  new_step_state(h);
  return 0;
}

}  // namespace

// ---- internal seam for the device-side Domain builder (engine_internal.h)
int hlb_internal_fail(const char* msg) { return fail(msg); }

int hlb_gpu_internal_raw(hlb_gpu_t h, hlb_gpu_raw* out) {
  if (!h || !out) return fail("null argument");
  if (h->finalised) return fail("handle already finalised");
  CU(cudaSetDevice(h->cfg.device));
  if (h->cfg.reorder && !h->coordsAll) CU(cudaMalloc(&h->coordsAll, sizeof(int32_t) * 3 * h->stride));
  out->nbr = h->nbr;
  out->coordsAll = h->coordsAll;
  out->stride = h->stride;
  out->hWall = h->hWall.data();
  out->hIolet = h->hIolet.data();
  out->hIoletId = h->hIoletId.data();
  out->hCut = h->hCut.data();
  out->hNormal = h->hNormal.data();
  out->hCoords = h->hCoords.data();
  out->bStride = h->bStride;
  out->NB = h->NB;
  return 0;
}

int hlb_gpu_internal_mark_installed(hlb_gpu_t h) {
  if (!h) return fail("null argument");
  h->haveNbr = h->haveSiteData = h->haveCut = h->haveNormal = h->haveCoords = true;
  if (h->cfg.reorder) h->coordsCovered = h->N;
  return 0;
}

int hlb_gpu_internal_view(hlb_gpu_t h, hlb_gpu_view* out) {
  if (!h || !out) return fail("null argument");
  if (!h->finalised) return fail("handle not finalised");
  if (join_aux(h)) return 1;
  new_step_state(h);  // (the checkpoint loader writes f through this view)
  out->Q = h->Q;
  out->device = h->cfg.device;
  out->rank = h->cfg.rank;
  out->nranks = h->cfg.nranks;
  out->N = h->N;
  out->stride = h->stride;
  out->bInfo = h->bInfo;
  out->bStride = h->bStride;
  out->f[0] = h->f[h->cur];
  out->f[1] = h->f[h->cur ^ 1];
  out->perm = h->perm;
  out->wallMask = h->wallMask;
  out->wallNormal = h->wallNormal;
  for (int i = 0; i < 8; ++i) out->cache[i] = h->cache[i];
  out->computeStream = (void*)h->compute;
  return 0;
}

int hlb_gpu_internal_count_launch(hlb_gpu_t h, int64_t n) {
  if (!h) return fail("null argument");
  h->launches += n;
  return 0;
}

extern "C" {

const char* hlb_gpu_last_error(void) { return g_err.c_str(); }

int hlb_gpu_device_count(int* count) {
  CU(cudaGetDeviceCount(count));
  return 0;
}

int hlb_gpu_create(const hlb_gpu_config* cfg, hlb_gpu_t* out) {
  if (!cfg || !out) return fail("null argument");
  const int Q = cfg->lattice;
  if (Q != 15 && Q != 19 && Q != 27) return fail("lattice must be 15, 19 or 27");
  if (cfg->kernel < 0 || cfg->kernel > 2) return fail("unknown kernel");
  if (cfg->kernel == HLB_KERNEL_MRT && Q == 27)
    return fail("No MRT basis for D3Q27 (Code/lb/Kernels.h:32-39)");
  if (cfg->wall < 0 || cfg->wall > 2) return fail("unknown wall boundary");
  if (cfg->inlet < 0 || cfg->inlet > 1 || cfg->outlet < 0 || cfg->outlet > 1) return fail("unknown iolet boundary");
  if (!(cfg->tau > 0.5)) return fail("tau must exceed 0.5");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail("no CUDA device: the collide-and-stream engine has no CPU fallback");
  if (cfg->device < 0 || cfg->device >= ndev) return fail("CUDA device ordinal out of range");
  CU(cudaSetDevice(cfg->device));
  hlb_gpu_handle* h = new hlb_gpu_handle();
  h->cfg = *cfg;
  h->Q = Q;
  h->N = cfg->n_sites;
  h->S = cfg->total_shared_fs;
  int64_t tot = 0;
  h->midTotal = 0;
  for (int t = 0; t < 6; ++t) {
    h->mid[t] = cfg->mid_count[t];
    h->edge[t] = cfg->edge_count[t];
    if (h->mid[t] < 0 || h->edge[t] < 0) { delete h; return fail("negative collision count"); }
    h->midTotal += h->mid[t];
    tot += h->mid[t] + h->edge[t];
  }
  if (tot != h->N) { delete h; return fail("collision counts do not sum to n_sites"); }
  {
    int64_t acc = 0;
    for (int t = 0; t < 6; ++t) { h->rangeFirst[t] = acc; acc += h->mid[t]; }
    for (int t = 0; t < 6; ++t) { h->rangeFirst[6 + t] = acc; acc += h->edge[t]; }
    h->rangeFirst[12] = acc;
  }
  h->midBulk = h->mid[0];
  h->edgeBulk = h->edge[0];
  h->NB = h->N - h->midBulk - h->edgeBulk;
  h->bStride = ((h->NB + 63) / 64) * 64;
  if (h->bStride == 0) h->bStride = 64;
  h->stride = ((h->N + 255) / 256) * 256;  // whole tiles of the TMA-staged kernel can be copied from every plane
  if (h->stride == 0) h->stride = 256;
  h->fLen = (int64_t)Q * h->stride + 1 + h->S;
  if (h->fLen >= ((int64_t)1 << 32)) { delete h; return fail("too many sites for 32-bit streaming indices on one GPU"); }
  CU(cudaStreamCreateWithFlags(&h->compute, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&h->comm, cudaStreamNonBlocking));
  {
    // HLB_SCHEDULE=0: every request its own launch, in the caller's order (as hlb_gpu_set_overlap(0))
    const char* e = getenv("HLB_SCHEDULE");
    h->schedule = !(e && e[0] == '0');
    h->scheduleDefault = h->schedule;
    const char* t = getenv("HLB_TMA");
    h->useTma = t && t[0] == '1';
    const char* go = getenv("HLB_GZS_OVERLAP");
    h->gzsOverlap = go ? atoi(go) : 0;
    {
      int lo = 0, hi = 0;
      CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      CU(cudaStreamCreateWithPriority(&h->aux, cudaStreamNonBlocking, hi));
      CU(cudaEventCreateWithFlags(&h->evFork, cudaEventDisableTiming));
      CU(cudaEventCreateWithFlags(&h->evJoin, cudaEventDisableTiming));
    }
    const char* nr = getenv("HLB_NBR_RUNS");
    h->useRuns = !(nr && nr[0] == '0');
    const char* pf = getenv("HLB_PREFETCH");
    // measured on the 1e8-site tree: 16 140 / 16 445 / 16 600 / 16 656 / 16 661 / 16 580 / 16 635 MLUPS at 10 240 /
    // 20 480 / 30 720 / 40 960 / 51 200 / 61 440 / 81 920 sites ahead (profiles/r02_prefetch_sweep_runs.json);
    // 15 % less without, and beyond ~150 000 the lines leave L2 again before they are used
    h->prefetchSites = pf ? atoi(pf) : 40960;
    h->prefetchSites = h->prefetchSites / 256 * 256;  // whole 128 B lines of every plane
    CU(cudaDeviceGetAttribute(&h->nSm, cudaDevAttrMultiProcessorCount, cfg->device));
  }
  CU(cudaEventCreateWithFlags(&h->evEdge, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&h->evComm, cudaEventDisableTiming));
  CU(cudaEventCreate(&h->evT0));
  CU(cudaEventCreate(&h->evT1));
  for (int i = 0; i < 2; ++i) {
    CU(cudaMalloc(&h->f[i], sizeof(double) * h->fLen));
    CU(cudaMemset(h->f[i], 0, sizeof(double) * h->fLen));
  }
  CU(cudaMalloc(&h->nbr, sizeof(uint32_t) * (Q - 1) * h->stride));
  CU(cudaMemset(h->nbr, 0xff, sizeof(uint32_t) * (Q - 1) * h->stride));
  CU(cudaMalloc(&h->wallMask, sizeof(uint32_t) * h->bStride));
  CU(cudaMalloc(&h->ioletMask, sizeof(uint32_t) * h->bStride));
  CU(cudaMalloc(&h->ioletId, sizeof(int32_t) * h->bStride));
  CU(cudaMalloc(&h->cutDist, sizeof(float) * (Q - 1) * h->bStride));
  CU(cudaMalloc(&h->wallNormal, sizeof(double) * 3 * h->bStride));
  CU(cudaMalloc(&h->coords, sizeof(int32_t) * 3 * h->bStride));
  CU(cudaMalloc(&h->streamIdx, sizeof(uint32_t) * std::max<int64_t>(h->S, 1)));
  CU(cudaMalloc(&h->monitorDev, 64));
  CU(cudaMalloc(&h->monitorSlots, sizeof(unsigned long long) * 4 * kMonitorSlots));
  monitor_arm_kernel<<<16, 256>>>(h->monitorSlots);
  CU(cudaGetLastError());
  for (int w = 0; w < 2; ++w) {
    const int n = std::max(1, w ? cfg->n_outlets : cfg->n_inlets);
    CU(cudaMalloc(&h->ioletsDev[w], sizeof(IoletDev) * n));
    CU(cudaMalloc(&h->ioletDensityDev[w], sizeof(double) * n));
    CU(cudaMallocHost(&h->ioletDensityPinned[w], sizeof(double) * n * 256));
  }
  h->hWall.assign(h->bStride, 0);
  h->hIolet.assign(h->bStride, 0);
  h->hIoletId.assign(h->bStride, -1);
  h->hCut.assign((size_t)(Q - 1) * h->bStride, -1.0f);
  h->hNormal.assign((size_t)3 * h->bStride, 0.0);
  h->hCoords.assign((size_t)3 * h->bStride, 0);
  h->omegaMinus = -1.0 / (0.5 + (3.0 / 16.0) / (cfg->tau - 0.5));  // TRT.h:100-107
#define HLB_PICK(QQ, KK, KE)                                      \
  if (Q == QQ && cfg->kernel == KK) {                             \
    h->launch = &launch_collide_stream<QQ, KE>;                   \
    h->launchTma = &launch_site_tma<QQ, KE>;                      \
    fill_mrt<QQ>(h);                                              \
  }
  HLB_PICK(15, 0, K_LBGK) HLB_PICK(15, 1, K_MRT) HLB_PICK(15, 2, K_TRT)
  HLB_PICK(19, 0, K_LBGK) HLB_PICK(19, 1, K_MRT) HLB_PICK(19, 2, K_TRT)
  HLB_PICK(27, 0, K_LBGK) HLB_PICK(27, 2, K_TRT)
#undef HLB_PICK
  if (cfg->n_neighbours == 0) h->haveNeighbours = true;
  if (h->S == 0) h->haveStream = true;
  if (cfg->n_inlets == 0) h->haveIolets[0] = true;
  if (cfg->n_outlets == 0) h->haveIolets[1] = true;
  *out = h;
  return 0;
}

int hlb_gpu_destroy(hlb_gpu_t h) {
  if (!h) return 0;
  cudaSetDevice(h->cfg.device);
  cudaDeviceSynchronize();
  if (h->comm_nccl && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm_nccl);
  for (int i = 0; i < 2; ++i) {
    cudaFree(h->f[i]);
    cudaFree(h->ioletsDev[i]);
    cudaFree(h->ioletDensityDev[i]);
    cudaFreeHost(h->ioletDensityPinned[i]);
  }
  for (int i = 0; i < 8; ++i) cudaFree(h->cache[i]);
  cudaFree(h->nbr);
  cudaFree(h->wallMask);
  cudaFree(h->ioletMask);
  cudaFree(h->ioletId);
  cudaFree(h->cutDist);
  cudaFree(h->wallNormal);
  cudaFree(h->coords);
  cudaFree(h->gzsNeighbour);
  cudaFree(h->streamIdx);
  cudaFree(h->staging);
  cudaFree(h->siteListDev);
  cudaFree(h->monitorDev);
  if (h->monitorPinned) cudaFreeHost(h->monitorPinned);
  if (h->monitorPinnedAsync) cudaFreeHost(h->monitorPinnedAsync);
  if (h->evMonitor) cudaEventDestroy(h->evMonitor);
  cudaFree(h->monitorSlots);
  cudaFree(h->perm);
  cudaFree(h->gzsGhost);
  cudaFree(h->gzsSendBuf);
  cudaFree(h->gzsServeDev);
  if (h->evGzsPack) cudaEventDestroy(h->evGzsPack);
  if (h->evGzsDone) cudaEventDestroy(h->evGzsDone);
  cudaFree(h->iperm);
  cudaFree(h->coordsAll);
  cudaFree(h->bInfo);
  cudaFree(h->bSite);
  cudaFree(h->bRec);
  cudaFree(h->nbrRuns);
  cudaFree(h->runFlags);
  if (h->aux) cudaStreamDestroy(h->aux);
  if (h->evFork) cudaEventDestroy(h->evFork);
  if (h->evJoin) cudaEventDestroy(h->evJoin);
  cudaFree(h->postI);
  cudaFree(h->postD);
  cudaFree(h->postQ);
  cudaEventDestroy(h->evEdge);
  cudaEventDestroy(h->evComm);
  cudaEventDestroy(h->evT0);
  cudaEventDestroy(h->evT1);
  cudaStreamDestroy(h->compute);
  cudaStreamDestroy(h->comm);
  delete h;
  return 0;
}

int hlb_gpu_set_neighbour_indices(hlb_gpu_t h, int64_t first, int64_t n, const int64_t* idx) {
  if (!h || !idx) return fail("null argument");
  if (first < 0 || n < 0 || first + n > h->N) return fail("site range outside the local fluid sites");
  CU(cudaSetDevice(h->cfg.device));
  const int Q = h->Q;
  if (ensure_staging(h, sizeof(int64_t) * kChunkSites * Q + 16)) return 1;
  int* bad = (int*)((char*)h->staging + sizeof(int64_t) * kChunkSites * Q);
  CU(cudaMemset(bad, 0, sizeof(int)));
  for (int64_t s0 = 0; s0 < n; s0 += kChunkSites) {
    const int64_t m = std::min(kChunkSites, n - s0);
    CU(cudaMemcpy(h->staging, idx + s0 * Q, sizeof(int64_t) * m * Q, cudaMemcpyHostToDevice));
    convert_nbr_kernel<<<blocks_for(m * Q), 256>>>((const int64_t*)h->staging, h->nbr, first + s0, m, Q, h->N,
                                                   h->stride, h->S, bad);
    CU(cudaGetLastError());
    CU(cudaDeviceSynchronize());
  }
  int hb = 0;
  CU(cudaMemcpy(&hb, bad, sizeof(int), cudaMemcpyDeviceToHost));
  if (hb == 1) return fail("neighbourIndices: direction 0 must stream to the site itself");
  if (hb) return fail("neighbourIndices: value outside the distribution array");
  h->haveNbr = true;
  return 0;
}

int hlb_gpu_get_neighbour_indices(hlb_gpu_t h, int64_t* idx) {
  if (!h || !idx) return fail("null argument");
  CU(cudaSetDevice(h->cfg.device));
  const int Q = h->Q;
  if (ensure_staging(h, sizeof(int64_t) * kChunkSites * Q + 16)) return 1;
  for (int64_t s0 = 0; s0 < h->N; s0 += kChunkSites) {
    const int64_t m = std::min(kChunkSites, h->N - s0);
    nbr_to_ref_kernel<<<blocks_for(m * Q), 256>>>(h->nbr, (int64_t*)h->staging, s0, m, Q, h->N, h->stride, h->perm,
                                                  h->iperm);
    CU(cudaGetLastError());
    CU(cudaMemcpy(idx + s0 * Q, h->staging, sizeof(int64_t) * m * Q, cudaMemcpyDeviceToHost));
  }
  return 0;
}

int hlb_gpu_set_site_data(hlb_gpu_t h, int64_t first, int64_t n, const uint32_t* wall, const uint32_t* iolet,
                          const int32_t* ioletId) {
  if (!h || !wall || !iolet || !ioletId) return fail("null argument");
  if (first < 0 || n < 0 || first + n > h->N) return fail("site range outside the local fluid sites");
  for (int64_t i = 0; i < n; ++i) {
    const int64_t b = host_bidx(h, first + i);
    if (b < 0) {
      if (wall[i] || iolet[i]) return fail("site data: a bulk-typed site has cut links");
      continue;
    }
    h->hWall[b] = wall[i];
    h->hIolet[b] = iolet[i];
    h->hIoletId[b] = ioletId[i];
  }
  h->haveSiteData = true;
  return 0;
}

int hlb_gpu_set_wall_distances(hlb_gpu_t h, int64_t first, int64_t n, const double* dist) {
  if (!h || !dist) return fail("null argument");
  if (first < 0 || n < 0 || first + n > h->N) return fail("site range outside the local fluid sites");
  const int Q = h->Q;
  for (int64_t i = 0; i < n; ++i) {
    const int64_t b = host_bidx(h, first + i);
    if (b < 0) continue;
    for (int d = 1; d < Q; ++d) {
      const double v = dist[i * (Q - 1) + d - 1];
      const float fv = (float)v;
      // cut distances originate as float32 in the .gmy (GeometrySiteLink.h:29); keep them so
      if ((double)fv != v) return fail("wall distance is not float32-representable");
      h->hCut[(size_t)(d - 1) * h->bStride + b] = fv;
    }
  }
  h->haveCut = true;
  return 0;
}

int hlb_gpu_set_wall_normals(hlb_gpu_t h, int64_t first, int64_t n, const double* normals) {
  if (!h || !normals) return fail("null argument");
  if (first < 0 || n < 0 || first + n > h->N) return fail("site range outside the local fluid sites");
  for (int64_t i = 0; i < n; ++i) {
    const int64_t b = host_bidx(h, first + i);
    if (b < 0) continue;
    for (int k = 0; k < 3; ++k) h->hNormal[(size_t)k * h->bStride + b] = normals[3 * i + k];
  }
  h->haveNormal = true;
  return 0;
}

int hlb_gpu_set_site_coords(hlb_gpu_t h, int64_t first, int64_t n, const int64_t* coords) {
  if (!h || !coords) return fail("null argument");
  if (first < 0 || n < 0 || first + n > h->N) return fail("site range outside the local fluid sites");
  for (int64_t i = 0; i < n; ++i) {
    const int64_t b = host_bidx(h, first + i);
    if (b < 0) continue;
    for (int k = 0; k < 3; ++k) h->hCoords[(size_t)k * h->bStride + b] = (int32_t)coords[3 * i + k];
  }
  h->haveCoords = true;
  if (h->cfg.reorder) {
    CU(cudaSetDevice(h->cfg.device));
    if (!h->coordsAll) CU(cudaMalloc(&h->coordsAll, sizeof(int32_t) * 3 * h->stride));
    if (ensure_staging(h, sizeof(int64_t) * kChunkSites * h->Q + 16)) return 1;
    for (int64_t s0 = 0; s0 < n; s0 += kChunkSites) {
      const int64_t m = std::min(kChunkSites, n - s0);
      CU(cudaMemcpy(h->staging, coords + 3 * s0, sizeof(int64_t) * 3 * m, cudaMemcpyHostToDevice));
      coords_to_planes_kernel<<<blocks_for(3 * m), 256>>>((const int64_t*)h->staging, h->coordsAll, first + s0, m,
                                                          h->stride);
      CU(cudaGetLastError());
      CU(cudaDeviceSynchronize());
    }
    h->coordsCovered += n;
  }
  return 0;
}

int hlb_gpu_set_neighbours(hlb_gpu_t h, const int* rank, const int64_t* count, const int64_t* first) {
  if (!h) return fail("null argument");
  if (h->cfg.n_neighbours > 0 && (!rank || !count || !first)) return fail("null neighbour arrays");
  h->neighbours.clear();
  int64_t expect = h->N * h->Q + 1, total = 0;
  for (int i = 0; i < h->cfg.n_neighbours; ++i) {
    if (first[i] != expect) return fail("FirstSharedDistribution is not the running prefix (Domain.cc:412-419)");
    if (rank[i] < 0 || rank[i] >= h->cfg.nranks || rank[i] == h->cfg.rank) return fail("bad neighbour rank");
    h->neighbours.push_back({rank[i], count[i], first[i]});
    expect += count[i];
    total += count[i];
  }
  if (total != h->S) return fail("SharedDistributionCounts do not sum to totalSharedFs");
  h->haveNeighbours = true;
  return 0;
}

int hlb_gpu_set_streaming_indices(hlb_gpu_t h, const int64_t* idx) {
  if (!h || (!idx && h->S)) return fail("null argument");
  CU(cudaSetDevice(h->cfg.device));
  std::vector<uint32_t> v(h->S);
  for (int64_t i = 0; i < h->S; ++i) {
    if (idx[i] < 0 || idx[i] >= h->N * h->Q) return fail("streamingIndicesForReceivedDistributions out of range");
    v[i] = (uint32_t)((idx[i] % h->Q) * h->stride + idx[i] / h->Q);
  }
  if (h->S) CU(cudaMemcpy(h->streamIdx, v.data(), sizeof(uint32_t) * h->S, cudaMemcpyHostToDevice));
  h->haveStream = true;
  return 0;
}

int hlb_gpu_set_iolets(hlb_gpu_t h, int which, int n, const double* rec) {
  if (!h || which < 0 || which > 1) return fail("bad argument");
  if (n != (which ? h->cfg.n_outlets : h->cfg.n_inlets)) return fail("iolet count differs from the config");
  CU(cudaSetDevice(h->cfg.device));
  std::vector<IoletDev> dev(n);
  h->ioletsHost[which].resize(n);
  for (int i = 0; i < n; ++i) {
    const double* r = rec + HLB_IOLET_RECORD_DOUBLES * i;
    IoletDev& d = dev[i];
    std::memset(&d, 0, sizeof(d));
    d.kind = (int)r[0];
    const double mag = std::sqrt(r[1] * r[1] + r[2] * r[2] + r[3] * r[3]);  // InOutLet.h:160-163
    if (!(mag > 0)) return fail("iolet normal has zero length");
    for (int k = 0; k < 3; ++k) {
      d.normal[k] = r[1 + k] / mag;
      d.position[k] = r[4 + k];
    }
    d.radius = r[7];
    d.maxSpeed = r[8];
    d.warmUpLength = r[13];
    IoletHost& hh = h->ioletsHost[which][i];
    hh.kind = d.kind;
    hh.densityMean = r[9];
    hh.densityAmp = r[10];
    hh.phase = r[11];
    hh.period = r[12];
    hh.warmUpLength = r[13];
    hh.minimumSimulationDensity = r[14];
  }
  if (n) CU(cudaMemcpy(h->ioletsDev[which], dev.data(), sizeof(IoletDev) * n, cudaMemcpyHostToDevice));
  h->haveIolets[which] = true;
  return 0;
}

int hlb_gpu_set_gzs_remote(hlb_gpu_t h, int64_t n, const int64_t* local_site, const int32_t* direction,
                           const int32_t* owner_rank, const int64_t* owner_site) {
  if (!h) return fail("null argument");
  if (h->finalised) return fail("hlb_gpu_set_gzs_remote after hlb_gpu_finalise");
  for (int64_t k = 0; k < n; ++k) {
    if (local_site[k] < 0 || local_site[k] >= h->N || host_bidx(h, local_site[k]) < 0)
      return fail("GZS remote need: not a boundary-typed local site");
    if (direction[k] < 1 || direction[k] >= h->Q) return fail("GZS remote need: bad direction");
    if (owner_rank[k] < 0 || owner_rank[k] >= h->cfg.nranks || owner_rank[k] == h->cfg.rank)
      return fail("GZS remote need: bad owner rank");
    if (k && owner_rank[k] < owner_rank[k - 1]) return fail("GZS remote needs must be grouped by ascending owner rank");
  }
  h->gzsNeedSite.assign(local_site, local_site + n);
  h->gzsNeedDir.assign(direction, direction + n);
  h->gzsNeedOwner.assign(owner_rank, owner_rank + n);
  h->gzsNeedOwnerSite.assign(owner_site, owner_site + n);
  return 0;
}

int hlb_gpu_set_gzs_serve(hlb_gpu_t h, int64_t n, const int32_t* requester_rank, const int64_t* local_site) {
  if (!h) return fail("null argument");
  if (h->finalised) return fail("hlb_gpu_set_gzs_serve after hlb_gpu_finalise");
  for (int64_t k = 0; k < n; ++k) {
    if (local_site[k] < 0 || local_site[k] >= h->N) return fail("GZS serve list: site outside the local fluid sites");
    if (requester_rank[k] < 0 || requester_rank[k] >= h->cfg.nranks || requester_rank[k] == h->cfg.rank)
      return fail("GZS serve list: bad requester rank");
    if (k && requester_rank[k] < requester_rank[k - 1]) return fail("GZS serve list must be grouped by ascending rank");
  }
  h->gzsServeRank.assign(requester_rank, requester_rank + n);
  h->gzsServeSite.assign(local_site, local_site + n);
  return 0;
}

int hlb_gpu_finalise(hlb_gpu_t h) {
  if (!h) return fail("null argument");
  if (!h->haveNbr && h->N) return fail("neighbourIndices not set");
  if (h->NB && !h->haveSiteData) return fail("site data not set");
  if (h->NB && !h->haveCut && h->cfg.wall != HLB_WALL_SIMPLEBOUNCEBACK) return fail("wall distances not set");
  if (!h->haveNeighbours) return fail("neighbouring processors not set");
  if (!h->haveStream) return fail("streaming indices for received distributions not set");
  if (!h->haveIolets[0] || !h->haveIolets[1]) return fail("iolets not set");
  CU(cudaSetDevice(h->cfg.device));
  const int Q = h->Q;
  {
    // iolet ids index the BoundaryValues object of the site's own streamer: the inlet table for the
    // inlet / inlet-wall ranges, the outlet table for the outlet / outlet-wall ranges
    // (lb.hpp:87-113 hands inletValues / outletValues to the respective streamers)
    const int64_t nbMid = h->midTotal - h->midBulk;
    for (int64_t b = 0; b < h->NB; ++b) {
      if (!h->hIolet[b]) continue;
      const int64_t site = b < nbMid ? b + h->midBulk : b - nbMid + h->midTotal + h->edgeBulk;
      int k = 0;
      while (k < 11 && site >= h->rangeFirst[k + 1]) ++k;
      const int type = k % 6;
      const int limit = (type == 2 || type == 4) ? h->cfg.n_inlets : (type == 3 || type == 5) ? h->cfg.n_outlets : 0;
      if (h->hIoletId[b] < 0) return fail("iolet site without an iolet id");
      if (h->hIoletId[b] >= limit)
        return fail("iolet id of a site is outside the iolet table of its range (ids index the inlet table in "
                    "inlet-typed ranges and the outlet table in outlet-typed ranges)");
      // which of the two BoundaryValues objects the id indexes travels with the id
      if (type == 3 || type == 5) h->hIoletId[b] |= kOutletTypedBit;
    }
  }
  for (int part = 0; part < 2; ++part) {
    h->sc[part] = hlb_gpu_handle::Deferred();
    for (int t = 0; t < 6; ++t)
      if ((part ? h->edge[t] : h->mid[t]) > 0) h->sc[part].all |= 1u << t;
  }
  h->post = hlb_gpu_handle::Deferred();
  if (h->cfg.wall == HLB_WALL_BFL)
    for (int k = 0; k < 12; ++k)
      if ((k % 6 == 1 || k % 6 >= 4) && h->rangeFirst[k + 1] > h->rangeFirst[k]) h->post.all |= 1u << k;
  if (h->cfg.reorder && h->N > 0) {
    if (h->coordsCovered != h->N)
      return fail("reorder requested but hlb_gpu_set_site_coords did not cover every site exactly once");
    if (build_permutation(h)) return 1;
  }
  if (build_boundary_order(h)) return 1;
  CU(cudaMemcpy(h->wallMask, h->hWall.data(), sizeof(uint32_t) * h->bStride, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(h->ioletMask, h->hIolet.data(), sizeof(uint32_t) * h->bStride, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(h->ioletId, h->hIoletId.data(), sizeof(int32_t) * h->bStride, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(h->cutDist, h->hCut.data(), sizeof(float) * (Q - 1) * h->bStride, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(h->wallNormal, h->hNormal.data(), sizeof(double) * 3 * h->bStride, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(h->coords, h->hCoords.data(), sizeof(int32_t) * 3 * h->bStride, cudaMemcpyHostToDevice));
  if (h->cfg.wall == HLB_WALL_GZS && h->NB) {
    CU(cudaMalloc(&h->gzsNeighbour, sizeof(int32_t) * (Q - 1) * h->bStride));
    gzs_neighbour_kernel<<<blocks_for(h->NB * (Q - 1)), 256>>>(h->nbr, h->gzsNeighbour, Q, h->stride, h->bStride,
                                                               h->NB, h->bSite);
    CU(cudaGetLastError());
    CU(cudaDeviceSynchronize());
    // a GZS link that would extrapolate from a site on another rank needs the phase-0 site halo
    std::vector<int32_t> g((size_t)(Q - 1) * h->bStride);
    std::vector<std::pair<int64_t, int>> missing;
    CU(cudaMemcpy(g.data(), h->gzsNeighbour, sizeof(int32_t) * g.size(), cudaMemcpyDeviceToHost));
    for (int64_t b = 0; b < h->NB; ++b)
      for (int d = 1; d < Q; ++d) {
        if (!((h->hWall[b] >> (d - 1)) & 1u)) continue;
        const int i = inv_dir(d);
        if (((h->hWall[b] >> (i - 1)) & 1u) || ((h->hIolet[b] >> (i - 1)) & 1u)) continue;
        if (!(h->hCut[(size_t)(d - 1) * h->bStride + b] < 0.75f)) continue;
        if (g[(size_t)(i - 1) * h->bStride + b] < 0 && g[(size_t)(i - 1) * h->bStride + b] == INT32_MIN)
          missing.push_back({b, i});
      }
    // remote rows: links that extrapolate from the same remote site (same owner rank and owner_site key)
    // share one ghost row -- NeighbouringDataManager::RegisterNeededSite keeps a site once
    // (NeighbouringDataManager.cc:27-39); rows are numbered in order of first appearance, which keeps
    // them grouped by owner rank like the links
    const int64_t nLinks = (int64_t)h->gzsNeedSite.size();
    std::vector<uint32_t> hp;
    if (h->perm && (nLinks || !h->gzsServeSite.empty())) {
      hp.resize(h->N);
      CU(cudaMemcpy(hp.data(), h->perm, sizeof(uint32_t) * h->N, cudaMemcpyDeviceToHost));
    }
    auto internal = [&](int64_t s) { return hp.empty() ? s : (int64_t)hp[s]; };
    {
      std::map<std::pair<int32_t, int64_t>, int64_t> rowOf;
      for (int64_t k = 0; k < nLinks; ++k) {
        const std::pair<int32_t, int64_t> key(h->gzsNeedOwner[k], h->gzsNeedOwnerSite[k]);
        auto it = rowOf.find(key);
        if (it == rowOf.end()) {
          it = rowOf.emplace(key, (int64_t)rowOf.size()).first;
          if (h->gzsRecvPeers.empty() || h->gzsRecvPeers.back().rank != key.first)
            h->gzsRecvPeers.push_back({key.first, it->second, 0});
          h->gzsRecvPeers.back().count++;
        }
        const int64_t b = h->refOrdToB[host_bidx(h, h->gzsNeedSite[k])];
        g[(size_t)(h->gzsNeedDir[k] - 1) * h->bStride + b] = (int32_t)(-(it->second + 1));
      }
      h->nGzsNeed = (int64_t)rowOf.size();
    }
    for (auto& ms : missing)
      if (g[(size_t)(ms.second - 1) * h->bStride + ms.first] == INT32_MIN)
        return fail("GZS wall link extrapolates from a site on another rank that hlb_gpu_set_gzs_remote did not name");
    CU(cudaMemcpy(h->gzsNeighbour, g.data(), sizeof(int32_t) * g.size(), cudaMemcpyHostToDevice));
    if (h->nGzsNeed) {
      CU(cudaMalloc(&h->gzsGhost, sizeof(double) * Q * h->nGzsNeed));
      CU(cudaMemset(h->gzsGhost, 0, sizeof(double) * Q * h->nGzsNeed));
    }
    h->nGzsServe = (int64_t)h->gzsServeSite.size();
    if (h->nGzsServe) {
      std::vector<uint32_t> sv(h->nGzsServe);
      for (int64_t k = 0; k < h->nGzsServe; ++k) {
        sv[k] = (uint32_t)internal(h->gzsServeSite[k]);
        if (h->gzsSendPeers.empty() || h->gzsSendPeers.back().rank != h->gzsServeRank[k])
          h->gzsSendPeers.push_back({h->gzsServeRank[k], k, 0});
        h->gzsSendPeers.back().count++;
      }
      CU(cudaMalloc(&h->gzsServeDev, sizeof(uint32_t) * h->nGzsServe));
      CU(cudaMemcpy(h->gzsServeDev, sv.data(), sizeof(uint32_t) * h->nGzsServe, cudaMemcpyHostToDevice));
      CU(cudaMalloc(&h->gzsSendBuf, sizeof(double) * Q * h->nGzsServe));
    }
    CU(cudaEventCreateWithFlags(&h->evGzsPack, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&h->evGzsDone, cudaEventDisableTiming));
  }
  {
    const int words = Q == 15 ? brec_words<15>() : (Q == 19 ? brec_words<19>() : brec_words<27>());
    CU(cudaMalloc(&h->bRec, sizeof(uint32_t) * words * std::max<int64_t>(h->NB, 1)));
    if (h->NB) {
      boundary_records_kernel<<<blocks_for(h->NB * words), 256>>>(h->wallMask, h->ioletMask, h->ioletId, h->cutDist,
                                                                  h->bStride, h->NB, Q, words, (uint32_t*)h->bRec);
      CU(cudaGetLastError());
      CU(cudaDeviceSynchronize());
    }
  }
  if (build_post_links(h)) return 1;
  if (h->useRuns && h->N > 0) {
    const int64_t nWords = (h->N + 31) / 32, nFlagWords = nWords / 32 + 2;
    const int words = Q == 15 ? brec_words<15>() : (Q == 19 ? brec_words<19>() : brec_words<27>());
    CU(cudaMalloc(&h->nbrRuns, sizeof(uint2) * (Q - 1) * (nWords + 1)));
    CU(cudaMalloc(&h->runFlags, sizeof(uint32_t) * nFlagWords));
    CU(cudaMemset(h->runFlags, 0, sizeof(uint32_t) * nFlagWords));
    nbr_runs_kernel<<<blocks_for(nWords * 32), 256>>>(h->nbr, h->stride, h->N, Q, h->bInfo, h->bRec, words / 4, h->nbrRuns,
                                                      h->runFlags, nWords);
    CU(cudaGetLastError());
    std::vector<uint32_t> fl(nFlagWords);
    CU(cudaMemcpy(fl.data(), h->runFlags, sizeof(uint32_t) * nFlagWords, cudaMemcpyDeviceToHost));
    h->runWordsTotal = nWords;
    h->runWords = 0;
    for (uint32_t v : fl) h->runWords += __builtin_popcount(v);
  }
  if (h->useTma && build_tensor_maps(h)) return 1;
  if (h->coordsAll) {
    cudaFree(h->coordsAll);
    h->coordsAll = nullptr;
  }
  h->finalised = true;
  return 0;
}

int hlb_gpu_comm_unique_id(void* id128) {
  if (!g_nccl.Load()) return fail("libnccl.so.2 could not be loaded");
  NcclUniqueId id;
  int rc = g_nccl.GetUniqueId(&id);
  if (rc) return fail(std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(rc));
  std::memcpy(id128, &id, 128);
  return 0;
}

int hlb_gpu_comm_init(hlb_gpu_t h, const void* id128) {
  if (!h || !id128) return fail("null argument");
  if (!g_nccl.Load()) return fail("libnccl.so.2 could not be loaded");
  CU(cudaSetDevice(h->cfg.device));
  NcclUniqueId id;
  std::memcpy(&id, id128, 128);
  int rc = g_nccl.CommInitRank(&h->comm_nccl, h->cfg.nranks, id, h->cfg.rank);
  if (rc) return fail(std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(rc));
  return 0;
}

int hlb_gpu_set_f(hlb_gpu_t h, int which, const double* f) {
  if (!h || !f) return fail("null argument");
  if (!h->finalised) return fail("hlb_gpu_set_f before hlb_gpu_finalise");
  CU(cudaSetDevice(h->cfg.device));
  if (join_aux(h)) return 1;
  new_step_state(h);
  CU(cudaStreamSynchronize(h->compute));
  double* dst = h->f[which ? h->cur ^ 1 : h->cur];
  const int Q = h->Q;
  if (ensure_staging(h, sizeof(int64_t) * kChunkSites * Q + 16)) return 1;
  for (int64_t s0 = 0; s0 < h->N; s0 += kChunkSites) {
    const int64_t m = std::min(kChunkSites, h->N - s0);
    CU(cudaMemcpy(h->staging, f + s0 * Q, sizeof(double) * m * Q, cudaMemcpyHostToDevice));
    aos_to_soa_kernel<<<blocks_for(m * Q), 256>>>((const double*)h->staging, dst, s0, m, Q, h->stride, h->perm);
    CU(cudaGetLastError());
    CU(cudaDeviceSynchronize());
  }
  CU(cudaMemcpy(dst + (int64_t)Q * h->stride, f + h->N * Q, sizeof(double) * (1 + h->S), cudaMemcpyHostToDevice));
  return 0;
}

int hlb_gpu_get_f(hlb_gpu_t h, int which, double* f) {
  if (!h || !f) return fail("null argument");
  CU(cudaSetDevice(h->cfg.device));
  if (join_aux(h)) return 1;
  CU(cudaStreamSynchronize(h->compute));
  CU(cudaStreamSynchronize(h->comm));
  const double* src = h->f[which ? h->cur ^ 1 : h->cur];
  const int Q = h->Q;
  if (ensure_staging(h, sizeof(int64_t) * kChunkSites * Q + 16)) return 1;
  for (int64_t s0 = 0; s0 < h->N; s0 += kChunkSites) {
    const int64_t m = std::min(kChunkSites, h->N - s0);
    soa_to_aos_kernel<<<blocks_for(m * Q), 256>>>(src, (double*)h->staging, s0, m, Q, h->stride, h->perm);
    CU(cudaGetLastError());
    CU(cudaMemcpy(f + s0 * Q, h->staging, sizeof(double) * m * Q, cudaMemcpyDeviceToHost));
  }
  CU(cudaMemcpy(f + h->N * Q, src + (int64_t)Q * h->stride, sizeof(double) * (1 + h->S), cudaMemcpyDeviceToHost));
  return 0;
}

int hlb_gpu_get_halo(hlb_gpu_t h, int which, double* out) {
  if (!h || (!out && h->S)) return fail("null argument");
  CU(cudaSetDevice(h->cfg.device));
  if (join_aux(h)) return 1;
  CU(cudaStreamSynchronize(h->compute));
  const double* src = h->f[which ? h->cur ^ 1 : h->cur] + (int64_t)h->Q * h->stride + 1;
  if (h->S) CU(cudaMemcpy(out, src, sizeof(double) * h->S, cudaMemcpyDeviceToHost));
  return 0;
}

int hlb_gpu_set_halo(hlb_gpu_t h, int which, const double* in) {
  if (!h || (!in && h->S)) return fail("null argument");
  CU(cudaSetDevice(h->cfg.device));
  if (join_aux(h)) return 1;
  CU(cudaStreamSynchronize(h->compute));
  double* dst = h->f[which ? h->cur ^ 1 : h->cur] + (int64_t)h->Q * h->stride + 1;
  if (h->S) CU(cudaMemcpy(dst, in, sizeof(double) * h->S, cudaMemcpyHostToDevice));
  h->edgePending = false;  // the caller moved the halo itself (host-staged exchange)
  if (which == 0) h->haloProvided = true;
  return 0;
}

int hlb_gpu_set_equilibrium(hlb_gpu_t h, double rho, const double* m) {
  if (!h || !m) return fail("null argument");
  CU(cudaSetDevice(h->cfg.device));
  if (join_aux(h)) return 1;
  new_step_state(h);
  // Lattice::CalculateFeq (scalar path) on the host, then broadcast to every site of both arrays
  double feq[27];
  const double density_1 = 1. / rho;
  const double mm = m[0] * m[0] + m[1] * m[1] + m[2] * m[2];
  for (int i = 0; i < h->Q; ++i) {
    int cx, cy, cz;
    double w;
    if (h->Q == 15) { cx = Lat<15>::cx(i); cy = Lat<15>::cy(i); cz = Lat<15>::cz(i); w = Lat<15>::W(i); }
    else if (h->Q == 19) { cx = Lat<19>::cx(i); cy = Lat<19>::cy(i); cz = Lat<19>::cz(i); w = Lat<19>::W(i); }
    else { cx = Lat<27>::cx(i); cy = Lat<27>::cy(i); cz = Lat<27>::cz(i); w = Lat<27>::W(i); }
    const double mde = cx * m[0] + cy * m[1] + cz * m[2];
    feq[i] = w * (rho - (3. / 2.) * mm * density_1 + (9. / 2.) * density_1 * mde * mde + 3. * mde);
  }
  if (ensure_staging(h, 512)) return 1;
  CU(cudaMemcpy(h->staging, feq, sizeof(double) * h->Q, cudaMemcpyHostToDevice));
  if (h->N) fill_planes_kernel<<<blocks_for(h->N * h->Q), 256>>>(h->f[0], h->f[1], h->N, h->stride, h->Q, (const double*)h->staging);
  CU(cudaGetLastError());
  CU(cudaDeviceSynchronize());
  return 0;
}

int hlb_gpu_set_step_scalars(hlb_gpu_t h, uint64_t timeStep, const double* inDens, const double* outDens,
                             uint32_t cacheMask) {
  if (!h) return fail("null argument");
  CU(cudaSetDevice(h->cfg.device));
  {
    // the same values again (every streamer of a step pushes them): nothing to do, and nothing held
    // back has to leave early
    bool same = h->scalarsSet && h->lastStep == timeStep && h->lastMask == cacheMask;
    for (int w = 0; w < 2 && same; ++w) {
      const int n = w ? h->cfg.n_outlets : h->cfg.n_inlets;
      const double* d = w ? outDens : inDens;
      if (n && !d) return fail("iolet densities missing");
      same = (int)h->lastDens[w].size() == n && (n == 0 || std::memcmp(h->lastDens[w].data(), d, sizeof(double) * n) == 0);
    }
    if (same) return 0;
  }
  if (join_aux(h)) return 1;  // deferred launches belong to the step whose scalars are still set
  h->sc[0].done = h->sc[1].done = h->post.done = 0;  // what follows runs with the new scalars
  h->timeStep = timeStep;
  if (ensure_caches(h, cacheMask)) return 1;
  h->cacheMask = cacheMask;
  h->monitorFused = (cacheMask & C_MONITOR) != 0;
  if (upload_densities(h, 0, inDens)) return 1;
  if (upload_densities(h, 1, outDens)) return 1;
  h->scalarsSet = true;
  h->lastStep = timeStep;
  h->lastMask = cacheMask;
  for (int w = 0; w < 2; ++w) {
    const int n = w ? h->cfg.n_outlets : h->cfg.n_inlets;
    const double* d = w ? outDens : inDens;
    h->lastDens[w].assign(d, d + n);
  }
  return 0;
}

int hlb_gpu_stream_and_collide(hlb_gpu_t h, int slot, int64_t first, int64_t count) {
  if (!h) return fail("null argument");
  CU(cudaSetDevice(h->cfg.device));
  return launch_range(h, slot, first, count, false);
}

int hlb_gpu_post_step(hlb_gpu_t h, int slot, int64_t first, int64_t count) {
  if (!h) return fail("null argument");
  CU(cudaSetDevice(h->cfg.device));
  return launch_range(h, slot, first, count, true);
}

int hlb_gpu_exchange_site_halo(hlb_gpu_t h) {
  if (!h) return fail("null argument");
  if (!h->finalised) return fail("handle not finalised");
  CU(cudaSetDevice(h->cfg.device));
  return exchange_site_halo(h);
}

int hlb_gpu_get_gzs_send(hlb_gpu_t h, double* out) {
  if (!h || (!out && h->nGzsServe)) return fail("null argument");
  CU(cudaSetDevice(h->cfg.device));
  if (join_aux(h)) return 1;
  CU(cudaStreamSynchronize(h->compute));
  if (h->nGzsServe) CU(cudaMemcpy(out, h->gzsSendBuf, sizeof(double) * h->Q * h->nGzsServe, cudaMemcpyDeviceToHost));
  return 0;
}

int hlb_gpu_set_gzs_ghost(hlb_gpu_t h, const double* in) {
  if (!h || (!in && h->nGzsNeed)) return fail("null argument");
  CU(cudaSetDevice(h->cfg.device));
  if (join_aux(h)) return 1;
  CU(cudaStreamSynchronize(h->compute));
  if (h->nGzsNeed) CU(cudaMemcpy(h->gzsGhost, in, sizeof(double) * h->Q * h->nGzsNeed, cudaMemcpyHostToDevice));
  return 0;
}

int hlb_gpu_request_comms(hlb_gpu_t h) {
  // LBM::RequestComms only *registers* the sends/receives with the Net (lb.hpp:162-173); they are
  // issued after PreSend.  Mirror that: remember, and post when the edge ranges have been issued.
  // It is also the first call of every time step of lb::LBM: a whole step follows, so each part may
  // leave as one launch at its first request (see hlb_gpu_handle::Deferred).
  if (!h) return fail("null argument");
  h->edgePending = true;
  h->announced = true;
  return 0;
}

int hlb_gpu_edge_done(hlb_gpu_t h) {
  if (!h) return fail("null argument");
  CU(cudaSetDevice(h->cfg.device));
  if (h->edgePending) {
    h->edgePending = false;
    return post_comms(h);
  }
  return 0;
}

int hlb_gpu_copy_received(hlb_gpu_t h) {
  if (!h) return fail("null argument");
  if (!h->finalised) return fail("handle not finalised");
  CU(cudaSetDevice(h->cfg.device));
  if (join_aux(h)) return 1;  // everything streamed before the received populations land
  if (h->edgePending) {  // caller never marked the end of PreSend: post now (no overlap)
    h->edgePending = false;
    if (post_comms(h)) return 1;
  }
  if (h->S == 0) return 0;
  if (h->commPosted) {
    CU(cudaStreamWaitEvent(h->compute, h->evComm, 0));  // Net::Wait
    h->commPosted = false;
  } else if (h->haloProvided) {
    h->haloProvided = false;
  } else {
    return fail("halo neither exchanged (hlb_gpu_comm_init + request_comms) nor provided (hlb_gpu_set_halo)");
  }
  copy_received_kernel<<<blocks_for(h->S), 256, 0, h->compute>>>(
      h->f[h->cur ^ 1], h->f[h->cur] + (int64_t)h->Q * h->stride + 1, h->streamIdx, h->S);
  h->launches++;
  CU(cudaGetLastError());
  return 0;
}

int hlb_gpu_swap(hlb_gpu_t h) {
  if (!h) return fail("null argument");
  if (join_aux(h)) return 1;
  h->cur ^= 1;
  new_step_state(h);
  return 0;
}

int hlb_gpu_get_cache(hlb_gpu_t h, uint32_t which, double* out) {
  if (!h || !out) return fail("null argument");
  static const int per[8] = {1, 3, 1, 1, 1, 9, 3, 3};
  CU(cudaSetDevice(h->cfg.device));
  for (int i = 0; i < 8; ++i)
    if (which == (1u << i)) {
      if (!h->cache[i]) return fail("cache was never requested");
      if (join_aux(h)) return 1;
  CU(cudaStreamSynchronize(h->compute));
      CU(cudaMemcpy(out, h->cache[i], sizeof(double) * per[i] * h->N, cudaMemcpyDeviceToHost));
      return 0;
    }
  return fail("unknown cache");
}

int hlb_gpu_step(hlb_gpu_t h, int nsteps) {
  if (!h) return fail("null argument");
  if (!h->finalised) return fail("handle not finalised");
  CU(cudaSetDevice(h->cfg.device));
  for (int i = 0; i < nsteps; ++i)
    if (one_step(h)) return 1;
  return 0;
}

int hlb_gpu_set_overlap(hlb_gpu_t h, int enabled) {
  if (!h) return fail("null argument");
  if (join_aux(h)) return 1;
  h->schedule = enabled ? h->scheduleDefault : false;
  return 0;
}

int hlb_gpu_get_time_step(hlb_gpu_t h, uint64_t* t) {
  if (!h || !t) return fail("null argument");
  *t = h->timeStep;
  return 0;
}

int hlb_gpu_sync(hlb_gpu_t h) {
  if (!h) return fail("null argument");
  CU(cudaSetDevice(h->cfg.device));
  if (join_aux(h)) return 1;
  CU(cudaStreamSynchronize(h->compute));
  CU(cudaStreamSynchronize(h->comm));
  return 0;
}

int hlb_gpu_time_steps(hlb_gpu_t h, int nsteps, float* ms) {
  if (!h || !ms) return fail("null argument");
  CU(cudaSetDevice(h->cfg.device));
  if (join_aux(h)) return 1;
  CU(cudaStreamSynchronize(h->compute));
  CU(cudaStreamSynchronize(h->comm));
  CU(cudaEventRecord(h->evT0, h->compute));
  for (int i = 0; i < nsteps; ++i)
    if (one_step(h)) return 1;
  CU(cudaEventRecord(h->evT1, h->compute));
  CU(cudaEventSynchronize(h->evT1));
  CU(cudaStreamSynchronize(h->comm));
  CU(cudaEventElapsedTime(ms, h->evT0, h->evT1));
  return 0;
}

int hlb_gpu_time_steps_detail(hlb_gpu_t h, int nsteps, float* total_ms, float* bulk_ms, int64_t* bulk_sites) {
  if (!h || !total_ms || !bulk_ms || !bulk_sites) return fail("null argument");
  h->profileBulk = true;
  h->profUsed = 0;
  h->profSites = 0;
  int rc = hlb_gpu_time_steps(h, nsteps, total_ms);
  h->profileBulk = false;
  if (rc) return rc;
  float acc = 0.f;
  for (size_t i = 0; i + 1 < h->profUsed; i += 2) {
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, h->profEv[i], h->profEv[i + 1]));
    acc += ms;
  }
  *bulk_ms = acc;
  *bulk_sites = h->profSites;
  return 0;
}

int hlb_gpu_monitor(hlb_gpu_t h, double* out4) {
  if (!h || !out4) return fail("null argument");
  if (h->monitorPending) return fail("hlb_gpu_monitor while a read-back begun by hlb_gpu_monitor_begin is outstanding");
  CU(cudaSetDevice(h->cfg.device));
  if (join_aux(h)) return 1;
  if (!h->monitorPinned) CU(cudaMallocHost(&h->monitorPinned, sizeof(double) * 4));
  if (h->monitorFused && h->monitorLaunches > 0) {
    h->monitorLaunches = 0;
    // gathered by the collide-and-stream kernels themselves during the last step(s): one small launch
    // and 32 bytes to pinned memory
    monitor_fold_decode_kernel<<<1, 256, 0, h->compute>>>(h->monitorSlots, h->monitorDev + 4);
    h->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(h->monitorPinned, h->monitorDev + 4, sizeof(double) * 4, cudaMemcpyDeviceToHost, h->compute));
    CU(cudaStreamSynchronize(h->compute));
    for (int k = 0; k < 4; ++k) out4[k] = h->monitorPinned[k];
    return 0;
  }
  unsigned long long init[4] = {~0ull, ~0ull, 0ull, 0ull};
  CU(cudaMemcpyAsync(h->monitorDev, init, sizeof(init), cudaMemcpyHostToDevice, h->compute));
  if (h->N) {
    const unsigned grid = (unsigned)std::min<int64_t>((h->N + 255) / 256, 148 * 16);
    unsigned long long* mo = (unsigned long long*)h->monitorDev;
    switch (h->Q) {
      case 15: monitor_kernel<15><<<grid, 256, 0, h->compute>>>(h->f[h->cur], h->N, h->stride, mo); break;
      case 19: monitor_kernel<19><<<grid, 256, 0, h->compute>>>(h->f[h->cur], h->N, h->stride, mo); break;
      case 27: monitor_kernel<27><<<grid, 256, 0, h->compute>>>(h->f[h->cur], h->N, h->stride, mo); break;
    }
    h->launches++;
  }
  monitor_decode_kernel<<<1, 32, 0, h->compute>>>((unsigned long long*)h->monitorDev);
  h->launches++;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(out4, h->monitorDev + 4, sizeof(double) * 4, cudaMemcpyDeviceToHost, h->compute));
  if (join_aux(h)) return 1;
  CU(cudaStreamSynchronize(h->compute));
  return 0;
}

// The same read-back split in two, so that a driver can issue the next time step before it waits for
// this one's 32 bytes (the values it gets are one step old, as the reference's monitors' are several:
// their PhasedBroadcast cycles span time steps)
int hlb_gpu_monitor_begin(hlb_gpu_t h) {
  if (!h) return fail("null argument");
  if (h->monitorPending) return fail("hlb_gpu_monitor_begin: the previous read-back has not been collected (hlb_gpu_monitor_end)");
  CU(cudaSetDevice(h->cfg.device));
  if (!h->monitorPinnedAsync) CU(cudaMallocHost(&h->monitorPinnedAsync, sizeof(double) * 4));
  if (!h->evMonitor) CU(cudaEventCreateWithFlags(&h->evMonitor, cudaEventDisableTiming));
  if (h->monitorFused && h->monitorLaunches > 0) {
    if (join_aux(h)) return 1;
    h->monitorLaunches = 0;
    monitor_fold_decode_kernel<<<1, 256, 0, h->compute>>>(h->monitorSlots, h->monitorDev + 4);
    h->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(h->monitorPinnedAsync, h->monitorDev + 4, sizeof(double) * 4, cudaMemcpyDeviceToHost, h->compute));
    CU(cudaEventRecord(h->evMonitor, h->compute));
    h->monitorPending = 1;
    return 0;
  }
  // nothing gathered in-kernel since the last read: the one-pass form, which is synchronous
  if (hlb_gpu_monitor(h, h->monitorPinnedAsync)) return 1;
  h->monitorPending = 2;
  return 0;
}

int hlb_gpu_monitor_end(hlb_gpu_t h, double* out4) {
  if (!h || !out4) return fail("null argument");
  if (!h->monitorPending) return fail("hlb_gpu_monitor_end without hlb_gpu_monitor_begin");
  if (h->monitorPending == 1) CU(cudaEventSynchronize(h->evMonitor));
  h->monitorPending = 0;
  for (int k = 0; k < 4; ++k) out4[k] = h->monitorPinnedAsync[k];
  return 0;
}

int hlb_gpu_monitor_global(hlb_gpu_t h, double* out4) {
  // StabilityTester / IncompressibilityChecker pass their values up and down a PhasedBroadcast tree
  // (Code/net/PhasedBroadcastRegular.h); here the four extrema of every rank meet in one
  // ncclAllReduce: min over {min f, min density, -max density, -max |u|}
  if (!h || !out4) return fail("null argument");
  double v[4];
  if (hlb_gpu_monitor(h, v)) return 1;  // this rank's values; the compute stream is drained on return
  if (h->cfg.nranks > 1) {
    if (!h->comm_nccl)
      return fail("hlb_gpu_monitor_global needs hlb_gpu_comm_init (host-staged runs reduce hlb_gpu_monitor's values "
                  "over their own communicator)");
    if (!g_nccl.AllReduce) return fail("ncclAllReduce not found in libnccl");
    v[2] = -v[2];
    v[3] = -v[3];
    // on the halo stream, behind whatever exchange was posted there: every NCCL call of this
    // communicator is issued on one stream, in the same order on every rank
    CU(cudaMemcpyAsync(h->monitorDev, v, sizeof(v), cudaMemcpyHostToDevice, h->comm));
    const int rc = g_nccl.AllReduce(h->monitorDev, h->monitorDev, 4, kNcclDouble, kNcclMin, h->comm_nccl, h->comm);
    if (rc) return fail(std::string("ncclAllReduce: ") + g_nccl.GetErrorString(rc));
    CU(cudaMemcpyAsync(v, h->monitorDev, sizeof(v), cudaMemcpyDeviceToHost, h->comm));
    CU(cudaStreamSynchronize(h->comm));
    v[2] = -v[2];
    v[3] = -v[3];
  }
  for (int k = 0; k < 4; ++k) out4[k] = v[k];
  return 0;
}

int hlb_gpu_stability(hlb_gpu_t h, int with_convergence, double* out2) {
  if (!h || !out2) return fail("null argument");
  if (!h->finalised) return fail("handle not finalised");
  CU(cudaSetDevice(h->cfg.device));
  if (join_aux(h)) return 1;
  unsigned long long init[2] = {0ull, 0x8000000000000000ull};  // {no failing population, mon_enc(+0.0)}
  CU(cudaMemcpyAsync(h->monitorDev, init, sizeof(init), cudaMemcpyHostToDevice, h->compute));
  if (h->N) {
    const unsigned grid = (unsigned)std::min<int64_t>((h->N + 255) / 256, 148 * 16);
    unsigned long long* mo = (unsigned long long*)h->monitorDev;
    const double *fNew = h->f[h->cur ^ 1], *fOld = h->f[h->cur];
    switch (h->Q) {
      case 15: stability_kernel<15><<<grid, 256, 0, h->compute>>>(fNew, fOld, h->N, h->stride, with_convergence, mo); break;
      case 19: stability_kernel<19><<<grid, 256, 0, h->compute>>>(fNew, fOld, h->N, h->stride, with_convergence, mo); break;
      case 27: stability_kernel<27><<<grid, 256, 0, h->compute>>>(fNew, fOld, h->N, h->stride, with_convergence, mo); break;
    }
    h->launches++;
  }
  stability_decode_kernel<<<1, 32, 0, h->compute>>>((unsigned long long*)h->monitorDev);
  h->launches++;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(out2, h->monitorDev + 2, sizeof(double) * 2, cudaMemcpyDeviceToHost, h->compute));
  CU(cudaStreamSynchronize(h->compute));
  return 0;
}

int hlb_gpu_launch_count(hlb_gpu_t h, int64_t* n) {
  if (!h || !n) return fail("null argument");
  *n = h->launches;
  return 0;
}

int hlb_gpu_target_runs(hlb_gpu_t h, int64_t* words_in_runs, int64_t* words) {
  if (!h || !words_in_runs || !words) return fail("null argument");
  if (!h->finalised) return fail("hlb_gpu_target_runs before hlb_gpu_finalise");
  *words_in_runs = h->runWords;
  *words = h->runWordsTotal;
  return 0;
}

}  // extern "C"
