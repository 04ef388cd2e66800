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
// out pointers into a host mirror that is refreshed on demand (initial conditions, checkpoints,
// StabilityTester, extraction of distributions); SendAndReceive / CopyReceived / SwapOldAndNew
// forward to the engine.  geometry::Domain declares `friend class FieldData`, which is how the
// index tables are read for upload.
#ifndef HEMELB_GEOMETRY_FIELDDATA_H
#define HEMELB_GEOMETRY_FIELDDATA_H

#include <map>
#include <memory>
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
  };

  // The six streamers are constructed from InitParams (LBM::InitCollisions, lb.hpp:75-114), which
  // names the Domain but not the FieldData; each constructor files its policy here under the
  // Domain, so that the engine -- built when the first range is asked for -- knows the wall and
  // iolet policies of all six.
  inline std::map<const Domain*, GpuPolicy>& GpuPolicies() {
    static std::map<const Domain*, GpuPolicy> all;
    return all;
  }
  inline GpuPolicy& GpuPolicyFor(const Domain* d) { return GpuPolicies()[d]; }

  class FieldData {
  public:
    using domain_type = Domain;

    explicit FieldData(std::shared_ptr<domain_type> d) :
        m_domain{d}, m_mirrorOld(CalcDistSize(*d)), m_mirrorNew(CalcDistSize(*d)),
        m_force(d->GetLocalFluidSiteCount()),
        m_neighbouringFields{std::make_unique<neighbouring::NeighbouringFieldData>(d->neighbouringData)} {}

    ~FieldData() { if (m_gpu) hlb_gpu_destroy(m_gpu); GpuPolicies().erase(m_domain.get()); }
    FieldData(FieldData const&) = delete;

    domain_type& GetDomain() { return *m_domain; }
    domain_type const& GetDomain() const { return *m_domain; }
    neighbouring::NeighbouringFieldData& GetNeighbouringData() { return *m_neighbouringFields; }
    neighbouring::NeighbouringFieldData const& GetNeighbouringData() const { return *m_neighbouringFields; }

    Site<FieldData> GetSite(site_t i) { return Site<FieldData>(i, *this); }
    Site<const FieldData> GetSite(site_t i) const { return Site<const FieldData>(i, *this); }

    // Host views.  Writing through GetFOld/GetFNew marks the mirror dirty; it is pushed to the
    // device before the next kernel.  Reading pulls it back first if the device copy is newer.
    distribn_t* GetFOld(site_t idx) { PullIfStale(); m_hostDirty = true; return &m_mirrorOld[idx]; }
    distribn_t const* GetFOld(site_t idx) const { const_cast<FieldData*>(this)->PullIfStale(); return &m_mirrorOld[idx]; }
    distribn_t* GetFNew(site_t idx) { PullIfStale(); m_hostDirty = true; return &m_mirrorNew[idx]; }
    distribn_t const* GetFNew(site_t idx) const { const_cast<FieldData*>(this)->PullIfStale(); return &m_mirrorNew[idx]; }
    template <typename LatticeType> auto GetFNew(site_t site) {
      constexpr auto Q = LatticeType::NUMVECTORS;
      return MutDistSpan<Q>{GetFNew(site * Q), Q};
    }

    void SwapOldAndNew() {  // FieldData.h:165-167
      if (m_gpu) { PushIfDirty(); Check(hlb_gpu_swap(m_gpu)); m_deviceNewer = true; }
      m_mirrorOld.swap(m_mirrorNew);
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
    hlb_gpu_t Engine() { EnsureEngine(); PushIfDirty(); m_deviceNewer = true; return m_gpu; }
    static void Check(int rc) { if (rc) throw Exception() << "hemelb_b200: " << hlb_gpu_last_error(); }

  private:
    static std::size_t CalcDistSize(Domain const& d) {
      return d.GetLocalFluidSiteCount() * d.latticeInfo.GetNumVectors() + 1 + d.totalSharedFs;
    }
    void PullIfStale() {
      if (m_gpu && m_deviceNewer) {
        Check(hlb_gpu_get_f(m_gpu, 0, m_mirrorOld.data()));
        Check(hlb_gpu_get_f(m_gpu, 1, m_mirrorNew.data()));
        m_deviceNewer = false;
      }
    }
    void PushIfDirty() {
      if (m_gpu && m_hostDirty) {
        Check(hlb_gpu_set_f(m_gpu, 0, m_mirrorOld.data()));
        Check(hlb_gpu_set_f(m_gpu, 1, m_mirrorNew.data()));
        m_hostDirty = false;
      }
    }
    void EnsureEngine();  // defined in lb/streamers/GpuStreamers.h (needs BoundaryValues)

    std::shared_ptr<domain_type> m_domain;
    std::vector<distribn_t> m_mirrorOld, m_mirrorNew;
    std::vector<LatticeForceVector> m_force;
    std::unique_ptr<neighbouring::NeighbouringFieldData> m_neighbouringFields;
    hlb_gpu_t m_gpu = nullptr;
    bool m_hostDirty = true, m_deviceNewer = false;
    friend struct GpuEngineBuilder;
  };
}
#endif
