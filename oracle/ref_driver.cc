// ---------------------------------------------------------------------------------------------
// ref_driver.cc -- TEST INFRASTRUCTURE ONLY.  Builds oracle/_ref/libhemelb_ref.so.
//
// Drives the UNMODIFIED reference lattice / kernel / collision / streamer / property-cache / iolet
// sources (compiled where they lie under /root/reference/Code; see oracle/Makefile) through a
// re-stated LBM phase loop (lb.hpp:176-309, FieldData.cc:27-48).  The containers they sit on
// (geometry::Domain, geometry::FieldData, lb::BoundaryValues) are the vector-backed shadow
// headers in oracle/ref_shim/.  Index tables are inputs (from the table builders under test).
//
// Not buildable from the reference and therefore absent here: TRT (TRT.h:42-90 does not compile),
// MRT + Nash iolets (MRT.h:73-86 does not compile).
// ---------------------------------------------------------------------------------------------
#include <cstdint>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <pthread.h>
#include <sched.h>

#include <atomic>
#include <barrier>
#include <thread>
#include <tuple>
#include <vector>

#include "lb/lattices/D3Q15.h"
#include "lb/lattices/D3Q19.h"
#include "lb/lattices/D3Q27.h"
#include "lb/kernels/LBGK.h"
#include "lb/kernels/MRT.h"
#include "lb/kernels/DHumieresD3Q15MRTBasis.h"
#include "lb/kernels/DHumieresD3Q19MRTBasis.h"
#ifdef HLB_REF_TRT
// lb/kernels/TRT.h has bit-rotted in the reference (lb/Kernels.h includes it, but nothing in a default build instantiates TRT): MakeOpposites() counts the
// rest direction as a pair and overruns its array at compile time (TRT.h:48, `iBar >= i`), and the bodies still use
// the `.f` member FVector lost when it became a std::array.  oracle/Makefile passes the file through three
// substitutions into a temporary include directory that comes first on the path (nothing is copied into the
// repository): `iBar >= i` -> `iBar > i`, `f_neq.f[` -> `f_neq[`, `f_eq.f[` -> `f_eq[`.  Only TRT::Collide is
// instantiated below -- its arithmetic is the reference's, token for token; CalculateDensityMomentumFeq /
// CalculateFeq (which call Lattice functions with signatures that no longer exist) never are.
#include "lb/kernels/TRT.h"
#endif
#include "lb/collisions/Normal.h"
#include "lb/iolets/BoundaryValues.h"
#include "lb/iolets/InOutLetCosine.h"
#include "lb/iolets/InOutLetParabolicVelocity.h"
#include "lb/streamers/StreamerTypeFactory.h"
#include "lb/streamers/BulkStreamer.h"
#include "lb/streamers/SimpleBounceBack.h"
#include "lb/streamers/BouzidiFirdaousLallemand.h"
#include "lb/streamers/GuoZhengShi.h"
#include "lb/streamers/NashZerothOrderPressure.h"
#include "lb/streamers/LaddIolet.h"
#include "lb/MacroscopicPropertyCache.h"

using namespace hemelb;

namespace {

struct NeighProc {
  int rank;
  site_t count, first;
};

struct RankState {
  geometry::Domain dom;
  geometry::FieldData fd;
  site_t mid[6], edge[6];
  site_t totalSharedFs = 0;
  std::vector<NeighProc> procs;
  std::vector<site_t> streamingIndices;
  std::vector<site_t> inputIndex;
  std::unique_ptr<lb::MacroscopicPropertyCache> cache;
  geometry::neighbouring::NeighbouringDataManager ndm;
};

struct SimBase {
  int Q = 0, R = 0;
  std::vector<std::unique_ptr<RankState>> ranks;
  lb::SimulationState state{1.0, 1000000000ul};
  lb::LbmParameters params;
  std::vector<std::unique_ptr<lb::InOutLet>> inletStore, outletStore;
  lb::BoundaryValues inletValues, outletValues;
  unsigned cacheMask = 0;
  bool memoryIsRankLocal = false;  // href_sim_step_mt: every rank's arrays re-made by its own (bound) thread
  // coordinate -> (rank, local id, input id) lookup shared by all ranks (GZS)
  std::function<bool(const util::Vector3D<site_t>&, int&, site_t&, site_t&)> lookup;

  virtual ~SimBase() = default;
  virtual void Init() = 0;
  virtual void StreamAndCollide(int r, int slot, site_t first, site_t count) = 0;
  virtual void PostStep(int r, int slot, site_t first, site_t count) = 0;

  void ApplyCacheMask() {
    for (auto& rs : ranks) {
      auto& c = *rs->cache;
      c.ResetRequirements();
      if (cacheMask & 1) c.densityCache.SetRefreshFlag();
      if (cacheMask & 2) c.velocityCache.SetRefreshFlag();
      if (cacheMask & 4) c.wallShearStressMagnitudeCache.SetRefreshFlag();
      if (cacheMask & 8) c.vonMisesStressCache.SetRefreshFlag();
      if (cacheMask & 16) c.shearRateCache.SetRefreshFlag();
      if (cacheMask & 32) c.stressTensorCache.SetRefreshFlag();
      if (cacheMask & 64) c.tractionCache.SetRefreshFlag();
      if (cacheMask & 128) c.tangentialProjectionTractionCache.SetRefreshFlag();
    }
  }

  void Step() {  // lb.hpp:176-309, FieldData.cc:27-48, SimulationMaster.impl.h:218-219
    ApplyCacheMask();
    for (int r = 0; r < R; ++r) {
      RankState& S = *ranks[r];
      site_t off = 0;
      for (int t = 0; t < 6; ++t) off += S.mid[t];
      for (int t = 0; t < 6; ++t) {
        StreamAndCollide(r, t, off, S.edge[t]);
        off += S.edge[t];
      }
    }
    for (int r = 0; r < R; ++r) {
      RankState& S = *ranks[r];
      site_t off = 0;
      for (int t = 0; t < 6; ++t) {
        StreamAndCollide(r, t, off, S.mid[t]);
        off += S.mid[t];
      }
    }
    for (int r = 0; r < R; ++r) {
      RankState& S = *ranks[r];
      for (auto& p : S.procs) {
        RankState& O = *ranks[p.rank];
        for (auto& po : O.procs)
          if (po.rank == r)
            for (site_t i = 0; i < p.count; ++i) O.fd.fOld[po.first + i] = S.fd.fNew[p.first + i];
      }
    }
    for (int r = 0; r < R; ++r) {
      RankState& S = *ranks[r];
      for (site_t i = 0; i < S.totalSharedFs; ++i)
        S.fd.fNew[S.streamingIndices[i]] = S.fd.fOld[S.dom.nSites * Q + 1 + i];
    }
    for (int r = 0; r < R; ++r) {
      RankState& S = *ranks[r];
      site_t off = 0;
      for (int t = 0; t < 6; ++t) off += S.mid[t];
      for (int t = 0; t < 6; ++t) {
        PostStep(r, t, off, S.edge[t]);
        off += S.edge[t];
      }
      off = 0;
      for (int t = 0; t < 6; ++t) {
        PostStep(r, t, off, S.mid[t]);
        off += S.mid[t];
      }
    }
    for (int r = 0; r < R; ++r) ranks[r]->fd.fOld.swap(ranks[r]->fd.fNew);
    state.Increment();
  }
};

template <class KernelT, template <class> class WallLink, template <class> class InLink,
          template <class> class OutLink>
struct RefSim : SimBase {
  using C = lb::Normal<KernelT>;
  using S0 = lb::BulkStreamer<C>;
  using S1 = lb::StreamerTypeFactory<WallLink<C>, lb::NullLink<C>>;
  using S2 = lb::StreamerTypeFactory<lb::NullLink<C>, InLink<C>>;
  using S3 = lb::StreamerTypeFactory<lb::NullLink<C>, OutLink<C>>;
  using S4 = lb::StreamerTypeFactory<WallLink<C>, InLink<C>>;
  using S5 = lb::StreamerTypeFactory<WallLink<C>, OutLink<C>>;
  static_assert(lb::streamer<S0> && lb::streamer<S1> && lb::streamer<S2> && lb::streamer<S4>);
  struct Streamers {
    std::unique_ptr<S0> s0;
    std::unique_ptr<S1> s1;
    std::unique_ptr<S2> s2;
    std::unique_ptr<S3> s3;
    std::unique_ptr<S4> s4;
    std::unique_ptr<S5> s5;
  };
  std::vector<Streamers> st;

  void Init() override {  // lb.hpp:75-114 InitCollisions
    st.resize(R);
    for (int r = 0; r < R; ++r) {
      RankState& S = *ranks[r];
      lb::InitParams ip;
      ip.latDat = &S.dom;
      ip.lbmParams = &params;
      ip.neighbouringDataManager = &S.ndm;
      ip.boundaryObject = nullptr;
      ip.siteRanges.resize(2);
      site_t midFirst = 0, edgeFirst = 0;
      for (int t = 0; t < 6; ++t) edgeFirst += S.mid[t];
      auto setRanges = [&](int t) {
        ip.siteRanges[0] = {midFirst, midFirst + S.mid[t]};
        ip.siteRanges[1] = {edgeFirst, edgeFirst + S.edge[t]};
        ip.siteCount = S.mid[t] + S.edge[t];
        midFirst += S.mid[t];
        edgeFirst += S.edge[t];
      };
      setRanges(0);
      st[r].s0 = std::make_unique<S0>(ip);
      setRanges(1);
      st[r].s1 = std::make_unique<S1>(ip);
      setRanges(2);
      ip.boundaryObject = &inletValues;
      st[r].s2 = std::make_unique<S2>(ip);
      setRanges(3);
      ip.boundaryObject = &outletValues;
      st[r].s3 = std::make_unique<S3>(ip);
      setRanges(4);
      ip.boundaryObject = &inletValues;
      st[r].s4 = std::make_unique<S4>(ip);
      setRanges(5);
      ip.boundaryObject = &outletValues;
      st[r].s5 = std::make_unique<S5>(ip);
    }
  }
  void StreamAndCollide(int r, int slot, site_t first, site_t count) override {
    RankState& S = *ranks[r];
    switch (slot) {
      case 0: st[r].s0->StreamAndCollide(first, count, &params, S.fd, *S.cache); break;
      case 1: st[r].s1->StreamAndCollide(first, count, &params, S.fd, *S.cache); break;
      case 2: st[r].s2->StreamAndCollide(first, count, &params, S.fd, *S.cache); break;
      case 3: st[r].s3->StreamAndCollide(first, count, &params, S.fd, *S.cache); break;
      case 4: st[r].s4->StreamAndCollide(first, count, &params, S.fd, *S.cache); break;
      case 5: st[r].s5->StreamAndCollide(first, count, &params, S.fd, *S.cache); break;
    }
  }
  void PostStep(int r, int slot, site_t first, site_t count) override {
    RankState& S = *ranks[r];
    switch (slot) {
      case 0: st[r].s0->PostStep(first, count, &params, S.fd, *S.cache); break;
      case 1: st[r].s1->PostStep(first, count, &params, S.fd, *S.cache); break;
      case 2: st[r].s2->PostStep(first, count, &params, S.fd, *S.cache); break;
      case 3: st[r].s3->PostStep(first, count, &params, S.fd, *S.cache); break;
      case 4: st[r].s4->PostStep(first, count, &params, S.fd, *S.cache); break;
      case 5: st[r].s5->PostStep(first, count, &params, S.fd, *S.cache); break;
    }
  }
};

#ifdef HLB_REF_TRT
  // The reference's TRT::Collide inside the reference's streamers.  TRT.h's own CalculateDensityMomentumFeq /
  // CalculateFeq call Lattice functions with signatures that no longer exist (no build of the reference instantiates TRT); they state
  // what LBGK.h:28-53 states, so this kernel takes those two from the text of LBGK.h's form and hands the collision
  // itself to the reference's TRT<L>::Collide (see the include above for how TRT.h reaches the compiler).
  template <lb::lattice_type L> class TrtOfTheReference {
  public:
    using LatticeType = L;
    using VarsType = lb::HydroVars<TrtOfTheReference>;
    TrtOfTheReference(lb::InitParams& ip) : inner(ip) {}
    void CalculateDensityMomentumFeq(VarsType& hv, site_t) {
      L::CalculateDensityMomentumFEq(hv.f, hv.density, hv.momentum, hv.velocity, hv.GetFEq());
      for (unsigned i = 0; i < L::NUMVECTORS; ++i) hv.SetFNeq(i, hv.f[i] - hv.GetFEq()[i]);
    }
    void CalculateFeq(VarsType& hv, site_t) {
      L::CalculateFeq(hv.density, hv.momentum, hv.GetFEq());
      for (unsigned i = 0; i < L::NUMVECTORS; ++i) hv.SetFNeq(i, hv.f[i] - hv.GetFEq()[i]);
    }
    void Collide(const lb::LbmParameters* p, VarsType& hv) {
      typename lb::TRT<L>::VarsType t(hv.f);
      t.tau = hv.tau;
      t.density = hv.density;
      t.momentum = hv.momentum;
      t.velocity = hv.velocity;
      for (unsigned i = 0; i < L::NUMVECTORS; ++i) {
        t.SetFEq(i, hv.GetFEq()[i]);
        t.SetFNeq(i, hv.GetFNeq()[i]);
      }
      inner.Collide(p, t);
      for (unsigned i = 0; i < L::NUMVECTORS; ++i) hv.SetFPostCollision(i, t.GetFPostCollision()[i]);
    }
  private:
    lb::TRT<L> inner;
  };
#endif

template <class KernelT, template <class> class WallLink>
SimBase* MakeIo(int inBC, int outBC, bool allowNash) {
  if (inBC == 0 && outBC == 0 && allowNash)
    return new RefSim<KernelT, WallLink, lb::NashZerothOrderPressureLink, lb::NashZerothOrderPressureLink>();
  if (inBC == 1 && outBC == 0 && allowNash)
    return new RefSim<KernelT, WallLink, lb::LaddIoletLink, lb::NashZerothOrderPressureLink>();
  if (inBC == 0 && outBC == 1 && allowNash)
    return new RefSim<KernelT, WallLink, lb::NashZerothOrderPressureLink, lb::LaddIoletLink>();
  if (inBC == 1 && outBC == 1)
    return new RefSim<KernelT, WallLink, lb::LaddIoletLink, lb::LaddIoletLink>();
  return nullptr;
}

// MRT + Nash does not compile in the reference (MRT.h:73-86): only instantiate Ladd/Ladd for MRT.
template <class KernelT, template <class> class WallLink>
SimBase* MakeIoMrt(int inBC, int outBC) {
#ifdef HLB_REF_MRT_NASH
  // (libhemelb_ref_mrtgzs.so only: MRT.h reaches the compiler through the substitutions of oracle/Makefile)
  return MakeIo<KernelT, WallLink>(inBC, outBC, true);
#endif
  if (inBC == 1 && outBC == 1)
    return new RefSim<KernelT, WallLink, lb::LaddIoletLink, lb::LaddIoletLink>();
  return nullptr;
}

template <class KernelT, bool MRTK>
SimBase* MakeWall(int wall, int inBC, int outBC) {
  if constexpr (MRTK) {
    switch (wall) {
      case 0: return MakeIoMrt<KernelT, lb::BounceBackLink>(inBC, outBC);
      case 1: return MakeIoMrt<KernelT, lb::BouzidiFirdaousLallemandLink>(inBC, outBC);
      case 2: return MakeIoMrt<KernelT, lb::GuoZhengShiLink>(inBC, outBC);
    }
  } else {
    switch (wall) {
      case 0: return MakeIo<KernelT, lb::BounceBackLink>(inBC, outBC, true);
      case 1: return MakeIo<KernelT, lb::BouzidiFirdaousLallemandLink>(inBC, outBC, true);
      case 2: return MakeIo<KernelT, lb::GuoZhengShiLink>(inBC, outBC, true);
    }
  }
  return nullptr;
}

SimBase* MakeSim(int Q, int kernel, int wall, int inBC, int outBC) {
  if (kernel == 0) {
    if (Q == 15) return MakeWall<lb::LBGK<lb::D3Q15>, false>(wall, inBC, outBC);
    if (Q == 19) return MakeWall<lb::LBGK<lb::D3Q19>, false>(wall, inBC, outBC);
    if (Q == 27) return MakeWall<lb::LBGK<lb::D3Q27>, false>(wall, inBC, outBC);
  } else if (kernel == 1) {
    if (Q == 15) return MakeWall<lb::MRT<lb::DHumieresD3Q15MRTBasis>, true>(wall, inBC, outBC);
    if (Q == 19) return MakeWall<lb::MRT<lb::DHumieresD3Q19MRTBasis>, true>(wall, inBC, outBC);
  }
#ifdef HLB_REF_TRT
  else if (kernel == 2) {
    if (Q == 15) return MakeWall<TrtOfTheReference<lb::D3Q15>, false>(wall, inBC, outBC);
    if (Q == 19) return MakeWall<TrtOfTheReference<lb::D3Q19>, false>(wall, inBC, outBC);
    if (Q == 27) return MakeWall<TrtOfTheReference<lb::D3Q27>, false>(wall, inBC, outBC);
  }
#endif
  return nullptr;
}

template <class L, class K>
void CollideOne(double dt, double dx, double rho, double eta, const double* f, double* fpost,
                double* feq, double* fneq, double* rmu, const double* rates) {
  lb::LbmParameters p(dt, dx, rho, eta);
  lb::InitParams ip;
  ip.lbmParams = &p;
  K k(ip);
  if constexpr (requires { k.SetMrtRelaxationParameters(std::declval<ConstDistSpan<K::NUMMOMENTS>>()); }) {
    if (rates) k.SetMrtRelaxationParameters(ConstDistSpan<K::NUMMOMENTS>(rates, K::NUMMOMENTS));
  }
  typename K::VarsType hv(f);
  hv.tau = p.GetTau();
  k.CalculateDensityMomentumFeq(hv, 0);
  k.Collide(&p, hv);
  for (unsigned i = 0; i < L::NUMVECTORS; ++i) {
    fpost[i] = hv.GetFPostCollision()[i];
    if (feq) feq[i] = hv.GetFEq()[i];
    if (fneq) fneq[i] = hv.GetFNeq()[i];
  }
  if (rmu) {
    rmu[0] = hv.density;
    for (int k2 = 0; k2 < 3; ++k2) {
      rmu[1 + k2] = hv.momentum[k2];
      rmu[4 + k2] = hv.velocity[k2];
    }
  }
}

#ifdef HLB_REF_TRT
// f_eq / f_neq as the (pinned) LBGK kernel computes them -- TRT.h:62-92 states the same two calls -- handed to
// the reference's TRT::Collide through HydroVars' public setters.
template <class L>
void TrtCollideOne(double dt, double dx, double rho, double eta, const double* f, double* fpost, double* feq,
                   double* fneq, double* rmu) {
  lb::LbmParameters p(dt, dx, rho, eta);
  lb::InitParams ip;
  ip.lbmParams = &p;
  lb::LBGK<L> lbgk(ip);
  typename lb::LBGK<L>::VarsType base(f);
  base.tau = p.GetTau();
  lbgk.CalculateDensityMomentumFeq(base, 0);
  lb::TRT<L> trt(ip);
  typename lb::TRT<L>::VarsType hv(f);
  hv.tau = p.GetTau();
  hv.density = base.density;
  hv.momentum = base.momentum;
  hv.velocity = base.velocity;
  for (unsigned i = 0; i < L::NUMVECTORS; ++i) {
    hv.SetFEq(i, base.GetFEq()[i]);
    hv.SetFNeq(i, base.GetFNeq()[i]);
  }
  trt.Collide(&p, hv);
  for (unsigned i = 0; i < L::NUMVECTORS; ++i) {
    fpost[i] = hv.GetFPostCollision()[i];
    if (feq) feq[i] = hv.GetFEq()[i];
    if (fneq) fneq[i] = hv.GetFNeq()[i];
  }
  if (rmu) {
    rmu[0] = hv.density;
    for (int k2 = 0; k2 < 3; ++k2) {
      rmu[1 + k2] = hv.momentum[k2];
      rmu[4 + k2] = hv.velocity[k2];
    }
  }
}
#endif

template <class L>
void StressOne(double rho, double tau, const double* fneq, const double* normal, double* out) {
  double sp = (1.0 - 1.0 / (2.0 * tau)) / std::sqrt(2.0);
  typename L::const_span f(fneq, L::NUMVECTORS);
  L::CalculateVonMisesStress(f, out[0], sp);
  util::Vector3D<double> n(normal[0], normal[1], normal[2]);
  L::CalculateWallShearStressMagnitude(rho, f, n, out[1], sp);
  out[2] = L::CalculateShearRate(tau, f, rho);
  util::Matrix3D s;
  L::CalculateStressTensor(rho, tau, f, s);
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) out[3 + 3 * a + b] = s[a][b];
  util::Vector3D<LatticeStress> t, tt;
  L::CalculateTractionOnAPoint(rho, tau, f, n, t);
  L::CalculateTangentialProjectionTraction(rho, tau, f, n, tt);
  for (int a = 0; a < 3; ++a) {
    out[12 + a] = t[a];
    out[15 + a] = tt[a];
  }
}

}  // namespace

// link-time stub: lb/SimulationState.cc references reporting::Dict (ctemplate wrapper)
namespace hemelb::reporting {
  Dict::Dict(const std::string&) : raw(nullptr, [](ctemplate::TemplateDictionary*) {}) {}
  Dict::Dict(ctemplate::TemplateDictionary*) : raw(nullptr, [](ctemplate::TemplateDictionary*) {}) {}
  Dict Dict::AddSectionDictionary(const std::string&) { return Dict(std::string()); }
  void Dict::SetValue(const std::string&, const std::string&) {}
  void Dict::SetIntValue(const std::string&, long) {}
  void Dict::SetBoolValue(const std::string&, bool) {}
  template <typename T> void Dict::SetFormattedValue(const std::string&, const char*, const T&) {}
  template void Dict::SetFormattedValue<double>(const std::string&, const char*, const double&);
}

extern "C" {

int href_is_sse3() {
#ifdef HEMELB_USE_SSE3
  return 1;
#else
  return 0;
#endif
}

double href_tau(double dt, double dx, double rho, double eta) {
  return lb::LbmParameters(dt, dx, rho, eta).GetTau();
}

int href_lattice(int Q, int* c, double* w, int* inv) {
  auto fill = [&](auto L) {
    using LT = decltype(L);
    for (unsigned i = 0; i < LT::NUMVECTORS; ++i) {
      c[3 * i] = LT::CX[i];
      c[3 * i + 1] = LT::CY[i];
      c[3 * i + 2] = LT::CZ[i];
      w[i] = LT::EQMWEIGHTS[i];
      inv[i] = LT::INVERSEDIRECTIONS[i];
    }
  };
  if (Q == 15) fill(lb::D3Q15{});
  else if (Q == 19) fill(lb::D3Q19{});
  else if (Q == 27) fill(lb::D3Q27{});
  else return 1;
  return 0;
}

int href_collide(int Q, int kernel, double dt, double dx, double rho, double eta, const double* f,
                 double* fpost, double* feq, double* fneq, double* rmu, const double* mrtRates) {
  if (kernel == 0 && Q == 15) CollideOne<lb::D3Q15, lb::LBGK<lb::D3Q15>>(dt, dx, rho, eta, f, fpost, feq, fneq, rmu, nullptr);
  else if (kernel == 0 && Q == 19) CollideOne<lb::D3Q19, lb::LBGK<lb::D3Q19>>(dt, dx, rho, eta, f, fpost, feq, fneq, rmu, nullptr);
  else if (kernel == 0 && Q == 27) CollideOne<lb::D3Q27, lb::LBGK<lb::D3Q27>>(dt, dx, rho, eta, f, fpost, feq, fneq, rmu, nullptr);
  else if (kernel == 1 && Q == 15) CollideOne<lb::D3Q15, lb::MRT<lb::DHumieresD3Q15MRTBasis>>(dt, dx, rho, eta, f, fpost, feq, fneq, rmu, mrtRates);
  else if (kernel == 1 && Q == 19) CollideOne<lb::D3Q19, lb::MRT<lb::DHumieresD3Q19MRTBasis>>(dt, dx, rho, eta, f, fpost, feq, fneq, rmu, mrtRates);
#ifdef HLB_REF_TRT
  else if (kernel == 2 && Q == 15) TrtCollideOne<lb::D3Q15>(dt, dx, rho, eta, f, fpost, feq, fneq, rmu);
  else if (kernel == 2 && Q == 19) TrtCollideOne<lb::D3Q19>(dt, dx, rho, eta, f, fpost, feq, fneq, rmu);
  else if (kernel == 2 && Q == 27) TrtCollideOne<lb::D3Q27>(dt, dx, rho, eta, f, fpost, feq, fneq, rmu);
#endif
  else return 1;
  return 0;
}

int href_feq(int Q, double rho, const double* m, double* feq) {
  util::Vector3D<double> mom(m[0], m[1], m[2]);
  if (Q == 15) lb::D3Q15::CalculateFeq(rho, mom, lb::D3Q15::mut_span(feq, 15));
  else if (Q == 19) lb::D3Q19::CalculateFeq(rho, mom, lb::D3Q19::mut_span(feq, 19));
  else if (Q == 27) lb::D3Q27::CalculateFeq(rho, mom, lb::D3Q27::mut_span(feq, 27));
  else return 1;
  return 0;
}

int href_mrt_basis(int Q, double tau, double* norms, double* rates) {
  if (Q == 15) {
    using B = lb::DHumieresD3Q15MRTBasis;
    auto s = B::SetUpCollisionMatrix(tau);
    for (unsigned k = 0; k < B::NUMMOMENTS; ++k) { norms[k] = B::BASIS_TIMES_BASIS_TRANSPOSED[k]; rates[k] = s[k]; }
    return B::NUMMOMENTS;
  }
  if (Q == 19) {
    using B = lb::DHumieresD3Q19MRTBasis;
    auto s = B::SetUpCollisionMatrix(tau);
    for (unsigned k = 0; k < B::NUMMOMENTS; ++k) { norms[k] = B::BASIS_TIMES_BASIS_TRANSPOSED[k]; rates[k] = s[k]; }
    return B::NUMMOMENTS;
  }
  return 0;
}

int href_stress_functions(int Q, double rho, double tau, const double* fneq, const double* normal, double* out) {
  if (Q == 15) StressOne<lb::D3Q15>(rho, tau, fneq, normal, out);
  else if (Q == 19) StressOne<lb::D3Q19>(rho, tau, fneq, normal, out);
  else if (Q == 27) StressOne<lb::D3Q27>(rho, tau, fneq, normal, out);
  else return 1;
  return 0;
}

double href_cosine_density(double mean, double amp, double phase, double period, uint64_t t) {
  lb::InOutLetCosine io;
  io.SetDensityMean(mean);
  io.SetDensityAmp(amp);
  io.SetPhase(phase);
  io.SetPeriod(period);
  io.SetWarmup(0);
  io.SetMinimumSimulationDensity(mean - amp);
  return io.GetDensity(t);
}

void href_parabolic_velocity(const double* normal, const double* position, double radius,
                             double maxSpeed, const double* x, uint64_t t, double* v) {
  lb::InOutLetParabolicVelocity io;
  io.SetNormal(util::Vector3D<double>(normal[0], normal[1], normal[2]));
  io.SetPosition(LatticePosition(position[0], position[1], position[2]));
  io.SetRadius(radius);
  io.SetMaxSpeed(maxSpeed);
  io.SetWarmup(0);
  auto r = io.GetVelocity(LatticePosition(x[0], x[1], x[2]), t);
  for (int k = 0; k < 3; ++k) v[k] = r[k];
}

// ---- simulation over supplied tables
void* href_sim_create(int Q, int kernel, int wall, int inBC, int outBC, double dt, double dx,
                      double rho, double eta, int R, int nInlets, const double* inletRec,
                      int nOutlets, const double* outletRec) {
  SimBase* S = MakeSim(Q, kernel, wall, inBC, outBC);
  if (!S) return nullptr;
  S->Q = Q;
  S->R = R;
  S->params = lb::LbmParameters(dt, dx, rho, eta);
  for (int r = 0; r < R; ++r) S->ranks.emplace_back(new RankState());
  auto mk = [&](int n, const double* rec, std::vector<std::unique_ptr<lb::InOutLet>>& store,
                lb::BoundaryValues& bv) {
    for (int i = 0; i < n; ++i) {
      const double* q = rec + 16 * i;
      lb::InOutLet* io;
      if ((int)q[0] == 0) {
        auto* c = new lb::InOutLetCosine();
        c->SetDensityMean(q[9]);
        c->SetDensityAmp(q[10]);
        c->SetPhase(q[11]);
        c->SetPeriod(q[12]);
        c->SetWarmup((unsigned)q[13]);
        io = c;
      } else {
        auto* v = new lb::InOutLetParabolicVelocity();
        v->SetRadius(q[7]);
        v->SetMaxSpeed(q[8]);
        v->SetWarmup((unsigned)q[13]);
        io = v;
      }
      io->SetNormal(util::Vector3D<double>(q[1], q[2], q[3]));
      io->SetPosition(LatticePosition(q[4], q[5], q[6]));
      io->SetMinimumSimulationDensity(q[14]);
      store.emplace_back(io);
      bv.iolets.push_back(io);
    }
    bv.state = &S->state;
  };
  mk(nInlets, inletRec, S->inletStore, S->inletValues);
  mk(nOutlets, outletRec, S->outletStore, S->outletValues);
  return S;
}

double href_sim_get_tau(void* sp) { return ((SimBase*)sp)->params.GetTau(); }

void href_sim_set_domain(void* sp, int r, int64_t N, const int64_t* counts12,
                         const int64_t* neighbourIndices, const uint32_t* wallMask,
                         const uint32_t* ioletMask, const int32_t* siteType, const int32_t* ioletId,
                         const double* distanceToWall, const double* wallNormal,
                         const int64_t* globalCoords, const int64_t* inputIndex,
                         int64_t totalSharedFs, int nprocs, const int64_t* procs3,
                         const int64_t* streamingIndices) {
  SimBase* S = (SimBase*)sp;
  RankState& X = *S->ranks[r];
  const int Q = S->Q;
  X.dom.nSites = N;
  X.dom.numVectors = Q;
  X.dom.localRank = r;
  for (int t = 0; t < 6; ++t) {
    X.mid[t] = counts12[t];
    X.edge[t] = counts12[6 + t];
  }
  X.dom.neighbourIndices.assign(neighbourIndices, neighbourIndices + N * Q);
  X.dom.distanceToWall.assign(distanceToWall, distanceToWall + N * (Q - 1));
  X.dom.wallNormalAtSite.resize(N);
  X.dom.globalSiteCoords.resize(N);
  X.dom.siteData.resize(N);
  X.inputIndex.assign(inputIndex, inputIndex + N);
  for (site_t i = 0; i < N; ++i) {
    X.dom.wallNormalAtSite[i] = util::Vector3D<double>(wallNormal[3 * i], wallNormal[3 * i + 1], wallNormal[3 * i + 2]);
    X.dom.globalSiteCoords[i] = util::Vector3D<site_t>(globalCoords[3 * i], globalCoords[3 * i + 1], globalCoords[3 * i + 2]);
    // Build the reference's SiteData through its own GeometrySite constructor
    // (SiteDataBare.cc:23-73) so that type / masks / ioletId come from reference code.
    geometry::GeometrySite gs(true);
    gs.links.resize(Q - 1);
    for (int d = 1; d < Q; ++d) {
      using CutType = io::formats::geometry::CutType;
      auto& link = gs.links[d - 1];
      if ((wallMask[i] >> (d - 1)) & 1u) link.type = CutType::WALL;
      else if ((ioletMask[i] >> (d - 1)) & 1u) {
        link.type = (siteType[i] == 2) ? CutType::INLET : CutType::OUTLET;
        link.ioletId = ioletId[i];
      }
    }
    X.dom.siteData[i] = geometry::SiteData(gs);
  }
  X.totalSharedFs = totalSharedFs;
  X.procs.clear();
  for (int p = 0; p < nprocs; ++p) X.procs.push_back({(int)procs3[3 * p], procs3[3 * p + 1], procs3[3 * p + 2]});
  X.streamingIndices.assign(streamingIndices, streamingIndices + totalSharedFs);
  X.fd.dom = &X.dom;
  X.fd.fOld.assign(N * Q + 1 + totalSharedFs, 0.0);
  X.fd.fNew.assign(N * Q + 1 + totalSharedFs, 0.0);
  X.fd.force.resize(N);
  X.cache = std::make_unique<lb::MacroscopicPropertyCache>(S->state, X.dom);
}

// call after all domains are set: wires coordinate lookups (GZS) and constructs the streamers
void href_sim_init(void* sp) {
  SimBase* S = (SimBase*)sp;
  struct Key {
    site_t x, y, z;
    bool operator<(const Key& o) const { return std::tie(x, y, z) < std::tie(o.x, o.y, o.z); }
  };
  auto table = std::make_shared<std::map<Key, std::tuple<int, site_t, site_t>>>();
  for (int r = 0; r < S->R; ++r) {
    RankState& X = *S->ranks[r];
    for (site_t i = 0; i < X.dom.nSites; ++i) {
      auto& c = X.dom.globalSiteCoords[i];
      (*table)[Key{c[0], c[1], c[2]}] = {r, i, X.inputIndex[i]};
    }
  }
  for (int r = 0; r < S->R; ++r) {
    RankState& X = *S->ranks[r];
    X.dom.procOf = [table](const util::Vector3D<site_t>& c) -> proc_t {
      auto it = table->find(Key{c[0], c[1], c[2]});
      return it == table->end() ? SITE_OR_BLOCK_SOLID : std::get<0>(it->second);
    };
    X.dom.contigOf = [table](const util::Vector3D<site_t>& c) -> site_t {
      return std::get<1>(table->at(Key{c[0], c[1], c[2]}));
    };
    X.dom.globalIdOf = [table](const util::Vector3D<site_t>& c) -> site_t {
      return std::get<2>(table->at(Key{c[0], c[1], c[2]}));
    };
    // remote f_old rows, resolved at read time (vectors are swapped every step)
    auto byGid = std::make_shared<std::map<site_t, std::pair<int, site_t>>>();
    for (auto& kv : *table) (*byGid)[std::get<2>(kv.second)] = {std::get<0>(kv.second), std::get<1>(kv.second)};
    X.fd.nfields.resolve = [S, byGid](site_t gid) -> const distribn_t* {
      auto& pr = byGid->at(gid);
      return &S->ranks[pr.first]->fd.fOld[pr.second * S->Q];
    };
  }
  S->Init();
}

void href_sim_destroy(void* sp) { delete (SimBase*)sp; }
void href_sim_set_time(void* sp, uint64_t t) {
  SimBase* S = (SimBase*)sp;
  S->state.Reset();
  for (uint64_t i = 1; i < t; ++i) S->state.Increment();
}
void href_sim_set_cache_mask(void* sp, unsigned mask) {
  SimBase* S = (SimBase*)sp;
  S->cacheMask = mask;
  S->ApplyCacheMask();
}
int64_t href_sim_get_cache(void* sp, int r, int bit, double* out) {
  SimBase* S = (SimBase*)sp;
  RankState& X = *S->ranks[r];
  auto& c = *X.cache;
  const site_t N = X.dom.nSites;
  auto scal = [&](auto& cache) { for (site_t i = 0; i < N; ++i) out[i] = cache.Get(i); return (int64_t)N; };
  auto vec = [&](auto& cache) { for (site_t i = 0; i < N; ++i) for (int k = 0; k < 3; ++k) out[3 * i + k] = cache.Get(i)[k]; return (int64_t)3 * N; };
  switch (bit) {
    case 1: return scal(c.densityCache);
    case 2: return vec(c.velocityCache);
    case 4: return scal(c.wallShearStressMagnitudeCache);
    case 8: return scal(c.vonMisesStressCache);
    case 16: return scal(c.shearRateCache);
    case 32:
      for (site_t i = 0; i < N; ++i)
        for (int a = 0; a < 3; ++a)
          for (int b = 0; b < 3; ++b) out[9 * i + 3 * a + b] = c.stressTensorCache.Get(i)[a][b];
      return 9 * N;
    case 64: return vec(c.tractionCache);
    case 128: return vec(c.tangentialProjectionTractionCache);
  }
  return -1;
}
int64_t href_sim_f_size(void* sp, int r) { return (int64_t)((SimBase*)sp)->ranks[r]->fd.fOld.size(); }
void href_sim_set_f(void* sp, int r, int which, const double* f) {
  auto& fd = ((SimBase*)sp)->ranks[r]->fd;
  auto& v = which ? fd.fNew : fd.fOld;
  std::memcpy(v.data(), f, v.size() * 8);
}
void href_sim_get_f(void* sp, int r, int which, double* f) {
  auto& fd = ((SimBase*)sp)->ranks[r]->fd;
  auto& v = which ? fd.fNew : fd.fOld;
  std::memcpy(f, v.data(), v.size() * 8);
}
void href_sim_stream_and_collide(void* sp, int r, int slot, int64_t first, int64_t count) {
  SimBase* S = (SimBase*)sp;
  S->ApplyCacheMask();
  S->StreamAndCollide(r, slot, first, count);
}
void href_sim_post_step(void* sp, int r, int slot, int64_t first, int64_t count) {
  ((SimBase*)sp)->PostStep(r, slot, first, count);
}
// One std::thread per emulated rank for the whole call (the reference is one single-threaded MPI process per
// rank).  Coupling as loose as the reference's: a rank publishes "my domain-edge sites of step k are done" (the
// send of lb.hpp:196-214), does its mid-domain sites, then waits only for ITS neighbours' flags before it pulls
// their slices (the wait of the receive), copies them in and runs PostStep; one barrier per step stands in for
// SimulationState::Increment, which the emulated ranks share.  Used for the CPU baseline only.
void href_sim_step_mt(void* sp, int n) {
  SimBase* S = (SimBase*)sp;
  const int R = S->R, Q = S->Q;
  std::vector<std::atomic<int64_t>> edgeDone(R);
  for (auto& e : edgeDone) e.store(0);
  std::barrier stepEnd(R, [S]() noexcept { S->state.Increment(); });
  std::barrier stepEnd0(R);
  S->ApplyCacheMask();
  // One core per emulated rank, as an MPI launcher binds its ranks: fresh threads all start on the caller's core,
  // and while they poll for their neighbours the scheduler took over a second to spread them (measured: the first
  // 30 steps at the rate of 1.4 cores, then 6x faster).
  std::vector<int> cpus;
  {
    cpu_set_t mask;
    CPU_ZERO(&mask);
    if (sched_getaffinity(0, sizeof mask, &mask) == 0)
      for (int c = 0; c < CPU_SETSIZE; ++c)
        if (CPU_ISSET(c, &mask)) cpus.push_back(c);
  }
  auto body = [&](int r) {
    if (!cpus.empty()) {
      cpu_set_t one;
      CPU_ZERO(&one);
      CPU_SET(cpus[(size_t)r % cpus.size()], &one);
      pthread_setaffinity_np(pthread_self(), sizeof one, &one);
    }
    RankState& X = *S->ranks[r];
    if (!S->memoryIsRankLocal) {
      // An MPI rank allocates and first touches its own arrays; here the caller's thread filled them.  On a host
      // with more than one memory node that would leave most ranks reading remote memory: re-make each array from
      // the rank's own thread once (values unchanged).
      auto own = [](auto& v) {
        std::remove_reference_t<decltype(v)> c(v.begin(), v.end());
        v.swap(c);
      };
      own(X.fd.fOld);
      own(X.fd.fNew);
      own(X.dom.neighbourIndices);
      own(X.dom.distanceToWall);
      own(X.dom.wallNormalAtSite);
      own(X.dom.siteData);
      own(X.dom.globalSiteCoords);
      own(X.streamingIndices);
    }
    site_t edge0 = 0;
    for (int t = 0; t < 6; ++t) edge0 += X.mid[t];
    // (no rank may read a neighbour's arrays while that neighbour is still re-making them)
    if (!S->memoryIsRankLocal) stepEnd0.arrive_and_wait();
    for (int64_t it = 1; it <= n; ++it) {
      site_t off = edge0;
      for (int t = 0; t < 6; ++t) { S->StreamAndCollide(r, t, off, X.edge[t]); off += X.edge[t]; }
      edgeDone[r].store(it, std::memory_order_release);
      off = 0;
      for (int t = 0; t < 6; ++t) { S->StreamAndCollide(r, t, off, X.mid[t]); off += X.mid[t]; }
      for (auto& p : X.procs) {
        while (edgeDone[p.rank].load(std::memory_order_acquire) < it) std::this_thread::yield();
        RankState& O = *S->ranks[p.rank];
        for (auto& po : O.procs)
          if (po.rank == r) std::copy_n(&O.fd.fNew[po.first], p.count, &X.fd.fOld[p.first]);
      }
      for (site_t i = 0; i < X.totalSharedFs; ++i)
        X.fd.fNew[X.streamingIndices[i]] = X.fd.fOld[X.dom.nSites * Q + 1 + i];
      off = edge0;
      for (int t = 0; t < 6; ++t) { S->PostStep(r, t, off, X.edge[t]); off += X.edge[t]; }
      off = 0;
      for (int t = 0; t < 6; ++t) { S->PostStep(r, t, off, X.mid[t]); off += X.mid[t]; }
      // every neighbour must have pulled this step's slices before the arrays change roles
      stepEnd.arrive_and_wait();
      X.fd.fOld.swap(X.fd.fNew);
    }
  };
  std::vector<std::thread> th;
  for (int r = 0; r < R; ++r) th.emplace_back(body, r);
  for (auto& t : th) t.join();
  S->memoryIsRankLocal = true;
}

void href_sim_step(void* sp, int n) {
  for (int i = 0; i < n; ++i) ((SimBase*)sp)->Step();
}

}  // extern "C"

// extraction (.xtr / .off) and checkpoint reading through the reference's own sources
#include "ref_xtr.h"
