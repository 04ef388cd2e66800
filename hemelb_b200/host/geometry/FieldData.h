// geometry/FieldData.h -- device-backed drop-in for HemeLB's geometry::FieldData.
//
// Put hemelb_b200/host FIRST on the include path of a HemeLB build: this header then replaces
// Code/geometry/FieldData.h (same class name, same member signatures; the reference's FieldData
// members are non-virtual and LBM / SimulationMaster hold a concrete geometry::FieldData*, so
// header substitution is the only seam -- Code/geometry/FieldData.h:117-212, Code/lb/lb.h:130,
// Code/SimulationMaster.h:83).  lb/lb.h, lb/lb.hpp, SimulationMaster*.h, SimBuilder, StepManager
// and BoundaryValues stay byte-identical.
//
// The distributions live on the GPU (handle from include/hemelb_b200.h).  GetFOld / GetFNew hand
// out pointers into host mirrors that are allocated and refreshed on demand, one array at a time
// (initial conditions, checkpoints, extraction of distributions; the per-step readers of the
// reference have device-side stand-ins: lb/StabilityTester.h here, the fused monitors);
// SendAndReceive / CopyReceived / SwapOldAndNew forward to the engine.  geometry::Domain declares `friend class FieldData`, which is how the
// index tables are read for upload.
#ifndef HEMELB_GEOMETRY_FIELDDATA_H
#define HEMELB_GEOMETRY_FIELDDATA_H

#include <map>
#include <memory>
#include <mutex>
#include <vector>

#include "Exception.h"
#include "constants.h"
#include "units.h"
#include "geometry/Domain.h"
#include "geometry/Site.h"
#include "geometry/neighbouring/NeighbouringDomain.h"
#include "lb/lattices/LatticeInfo.h"
#include "hemelb_b200.h"

namespace hemelb::net { class Net; }
namespace hemelb::lb { class BoundaryValues; }
namespace hemelb::geometry::neighbouring { class NeighbouringDataManager; }

namespace hemelb::geometry {

  // what the six streamers of one LBM tell the engine before the first step
  struct GpuPolicy {
    int kernel = -1, wall = -1, inlet = -1, outlet = -1;
    double tau = 0.0;
    lb::BoundaryValues* inletValues = nullptr;
    lb::BoundaryValues* outletValues = nullptr;
    // what hlb_gpu_set_step_scalars was last given: the twelve StreamAndCollide calls of one time
    // step (lb.hpp:176-255) push the scalars once, not twelve times -- every push would flush the
    // engine's deferred mid-domain launches and cost a host-to-device copy
    bool scalarsPushed = false;
    unsigned long long pushedStep = 0;
    unsigned pushedMask = 0;
    std::vector<double> pushedIn, pushedOut;
    // LBM calls every streamer twice per phase, domain-edge range first (PreSend / PostReceive's
    // first half), mid-domain range second; the ranges themselves cannot tell the two apart when
    // they are empty, so the outlet-wall streamer (the last of the six) counts its calls
    unsigned lastSlotStreams = 0, lastSlotPostSteps = 0;
    // GuoZhengShi across ranks: one entry per wall link that extrapolates from a site on another rank,
    // in the order GuoZhengShiLink's constructor registers them (GuoZhengShi.h:36-104), and the
    // manager the needs were registered with (its GetNeedsForProc lists give the serve side)
    struct GzsLink { site_t site; int direction; site_t globalId; };
    std::vector<GzsLink> gzsLinks;
    neighbouring::NeighbouringDataManager* gzsManager = nullptr;
    // for the monitors that are constructed with the Domain, not the FieldData (lb/IncompressibilityChecker.h):
    // the engine once it exists, and whether the site kernel should gather density / velocity extrema
    hlb_gpu_t engine = nullptr;
    bool monitorRequested = false;
  };

  // The six streamers are constructed from InitParams (LBM::InitCollisions, lb.hpp:75-114), which
  // names the Domain but not the FieldData; each constructor files its policy here under the
  // Domain, so that the engine -- built when the first range is asked for -- knows the wall and
  // iolet policies of all six.
  // (one rank per process in a HemeLB build; the lock is for harnesses whose ranks are threads of one
  // process -- entries are per Domain, and std::map keeps references to them valid)
  inline std::map<const Domain*, GpuPolicy>& GpuPolicies() {
    static std::map<const Domain*, GpuPolicy> all;
    return all;
  }
  inline std::mutex& GpuPoliciesLock() {
    static std::mutex lock;
    return lock;
  }
  inline GpuPolicy& GpuPolicyFor(const Domain* d) {
    std::lock_guard<std::mutex> guard(GpuPoliciesLock());
    return GpuPolicies()[d];
  }
  inline void ForgetGpuPolicy(const Domain* d) {
    std::lock_guard<std::mutex> guard(GpuPoliciesLock());
    GpuPolicies().erase(d);
  }

  class FieldData {
  public:
    using domain_type = Domain;

    explicit FieldData(std::shared_ptr<domain_type> d) :
        m_domain{d}, m_force(d->GetLocalFluidSiteCount()),
        m_neighbouringFields{std::make_unique<neighbouring::NeighbouringFieldData>(d->neighbouringData)} {}

    ~FieldData() { if (m_gpu) hlb_gpu_destroy(m_gpu); ForgetGpuPolicy(m_domain.get()); }
    FieldData(FieldData const&) = delete;

    domain_type& GetDomain() { return *m_domain; }
    domain_type const& GetDomain() const { return *m_domain; }
    neighbouring::NeighbouringFieldData& GetNeighbouringData() { return *m_neighbouringFields; }
    neighbouring::NeighbouringFieldData const& GetNeighbouringData() const { return *m_neighbouringFields; }

    Site<FieldData> GetSite(site_t i) { return Site<FieldData>(i, *this); }
    Site<const FieldData> GetSite(site_t i) const { return Site<const FieldData>(i, *this); }

    // Host views.  The two host mirrors are allocated when first asked for (a run that never looks at
    // the distributions on the host -- monitors on the device, extraction encoded on the device --
    // never pays for them) and each is refreshed on its own: reading one pulls that array only, and
    // only if the device copy is newer; writing through the non-const overloads marks that array
    // dirty, and it alone is pushed to the device before the next kernel.
    distribn_t* GetFOld(site_t idx) { Pull(0); m_dirty[0] = true; return &m_mirror[0][idx]; }
    distribn_t const* GetFOld(site_t idx) const { const_cast<FieldData*>(this)->Pull(0); return &m_mirror[0][idx]; }
    distribn_t* GetFNew(site_t idx) { Pull(1); m_dirty[1] = true; return &m_mirror[1][idx]; }
    distribn_t const* GetFNew(site_t idx) const { const_cast<FieldData*>(this)->Pull(1); return &m_mirror[1][idx]; }
    template <typename LatticeType> auto GetFNew(site_t site) {
      constexpr auto Q = LatticeType::NUMVECTORS;
      return MutDistSpan<Q>{GetFNew(site * Q), Q};
    }

    void SwapOldAndNew() {  // FieldData.h:165-167
      if (m_gpu) { PushIfDirty(); Check(hlb_gpu_swap(m_gpu)); }
      m_mirror[0].swap(m_mirror[1]);
      std::swap(m_stale[0], m_stale[1]);
    }
    void SendAndReceive(net::Net*) {  // FieldData.cc:27-39 -> NCCL send/recv posted after PreSend
      // LBM::RequestComms is the first call of every time step, the first step included: the
      // engine is built here if it does not exist yet (the streamers were constructed, and filed
      // their policies, in LBM::InitCollisions)
      Check(hlb_gpu_request_comms(Engine()));
    }
    void CopyReceived() {  // FieldData.cc:41-48
      Check(hlb_gpu_copy_received(Engine()));
    }

    void ResetForces(LatticeForceVector const& f = LatticeForceVector(0, 0, 0)) { std::fill(m_force.begin(), m_force.end(), f); }
    LatticeForceVector const& GetForceAtSite(site_t i) const { return m_force[i]; }
    void SetForceAtSite(site_t i, LatticeForceVector const& f) { m_force[i] = f; }
    void AddToForceAtSite(site_t i, LatticeForceVector const& f) { m_force[i] += f; }

    // ---- used by the Gpu*Streamer policy classes ------------------------------------------------
    GpuPolicy& Policy() { return GpuPolicyFor(m_domain.get()); }
    // for callers that are about to change the distributions on the device
    hlb_gpu_t Engine() {
      EnsureEngine();
      PushIfDirty();
      if (m_needFirstSiteHalo) {
        // the first time step's phase 0 (NeighbouringDataManager::RequestComms) ran before the engine
        // existed: its site halo goes now, with the initial condition already on the device
        m_needFirstSiteHalo = false;
        Check(hlb_gpu_exchange_site_halo(m_gpu));
      }
      m_stale[0] = m_stale[1] = true;
      return m_gpu;
    }
    // for callers that only read them there (lb/StabilityTester.h): null until a streamer has run
    hlb_gpu_t EngineIfBuilt() { if (m_gpu) PushIfDirty(); return m_gpu; }
    static void Check(int rc) { if (rc) throw Exception() << "hemelb_b200: " << hlb_gpu_last_error(); }

  private:
    static std::size_t CalcDistSize(Domain const& d) {
      return d.GetLocalFluidSiteCount() * d.latticeInfo.GetNumVectors() + 1 + d.totalSharedFs;
    }
    void Pull(int which) {
      if (m_mirror[which].empty()) m_mirror[which].resize(CalcDistSize(*m_domain));
      if (m_gpu && m_stale[which]) {
        Check(hlb_gpu_get_f(m_gpu, which, m_mirror[which].data()));
        m_stale[which] = false;
      }
    }
    void PushIfDirty() {
      for (int which = 0; m_gpu && which < 2; ++which)
        if (m_dirty[which]) {
          Check(hlb_gpu_set_f(m_gpu, which, m_mirror[which].data()));
          m_dirty[which] = false;
        }
    }
    void EnsureEngine();  // defined in lb/streamers/GpuStreamers.h (needs BoundaryValues)

    std::shared_ptr<domain_type> m_domain;
    std::vector<distribn_t> m_mirror[2];  // [0] f_old, [1] f_new; empty until first asked for
    std::vector<LatticeForceVector> m_force;
    std::unique_ptr<neighbouring::NeighbouringFieldData> m_neighbouringFields;
    hlb_gpu_t m_gpu = nullptr;
    bool m_dirty[2] = {false, false}, m_stale[2] = {false, false}, m_needFirstSiteHalo = false;
    friend struct GpuEngineBuilder;
  };
}
#endif
