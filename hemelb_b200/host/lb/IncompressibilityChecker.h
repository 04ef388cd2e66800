// lb/IncompressibilityChecker.h -- lb::IncompressibilityChecker for a build whose distributions live on the B200.
//
// Stands in for Code/lb/IncompressibilityChecker.h / .hpp (same class template, nested DensityTracker,
// constructor and PhasedBroadcast role, so configuration/SimBuilder.h:216-226 and SimulationMaster.h:118
// compile unchanged) when hemelb_b200/host precedes Code/ on the include path.  The reference's checker
// reads propertyCache.densityCache / velocityCache of every local site on the host each cycle
// (IncompressibilityChecker.hpp:184-196); with the engine those caches would have to be pulled from the
// device -- N x 32 bytes per cycle -- only to be reduced to three numbers.  Here the three numbers come
// from the device: hlb_gpu_monitor (gathered inside the site kernel when HLB_CACHE_MONITOR is in the
// step's cache mask -- GpuStreamers.h sets it while a checker is registered -- else one pass), 32 bytes.
// How the trackers of the ranks meet, up and down the PhasedBroadcast tree, is the reference's own
// protocol and untouched; as there, a node's tracker only ever widens (it is never reset).
#ifndef HEMELB_LB_INCOMPRESSIBILITYCHECKER_H
#define HEMELB_LB_INCOMPRESSIBILITYCHECKER_H

#include <algorithm>
#include <array>
#include <cfloat>

#include "geometry/Domain.h"
#include "geometry/FieldData.h"
#include "lb/MacroscopicPropertyCache.h"
#include "net/PhasedBroadcastRegular.h"
#include "reporting/Reportable.h"
#include "reporting/Timers.h"
#include "hassert.h"

namespace hemelb::lb
{
  static const distribn_t REFERENCE_DENSITY = 1.0;  // IncompressibilityChecker.h:26

  template<class BroadcastPolicy>
  class IncompressibilityChecker : public BroadcastPolicy, public reporting::Reportable
  {
    public:
      // {smallest density, largest density, largest velocity magnitude}: over its own storage, or over
      // three doubles of a message buffer
      class DensityTracker
      {
        public:
          static const unsigned DENSITY_TRACKER_SIZE = 3u;
          typedef enum { MIN_DENSITY = 0u, MAX_DENSITY, MAX_VELOCITY_MAGNITUDE } DensityTrackerIndices;

          DensityTracker() : own { DBL_MAX, -DBL_MAX, 0.0 }, values(own.data()) {}
          DensityTracker(distribn_t* const densityValues) : own { }, values(densityValues) {}
          DensityTracker(const DensityTracker& other) : own(other.own), values(other.Wraps() ? other.values : own.data()) {}
          ~DensityTracker() = default;

          void operator=(const DensityTracker& newValues) { std::copy_n(newValues.values, DENSITY_TRACKER_SIZE, values); }
          distribn_t& operator[](DensityTrackerIndices densityIndex) const { return values[densityIndex]; }
          distribn_t* GetDensitiesArray() const { return values; }

          void UpdateDensityTracker(const DensityTracker& newValues)
          {
            Widen(newValues[MIN_DENSITY], newValues[MAX_DENSITY], newValues[MAX_VELOCITY_MAGNITUDE]);
          }
          void UpdateDensityTracker(distribn_t newDensity, distribn_t newVelocityMagnitude)
          {
            Widen(newDensity, newDensity, newVelocityMagnitude);
          }
          // the device's reduction over the local sites: a range of densities at once
          void Widen(distribn_t lowDensity, distribn_t highDensity, distribn_t velocityMagnitude)
          {
            if (lowDensity < values[MIN_DENSITY]) values[MIN_DENSITY] = lowDensity;
            if (highDensity > values[MAX_DENSITY]) values[MAX_DENSITY] = highDensity;
            if (velocityMagnitude > values[MAX_VELOCITY_MAGNITUDE]) values[MAX_VELOCITY_MAGNITUDE] = velocityMagnitude;
          }

        private:
          bool Wraps() const { return values != own.data(); }
          std::array<distribn_t, DENSITY_TRACKER_SIZE> own;
          distribn_t* values;
      };

      IncompressibilityChecker(const geometry::Domain* latticeData, net::Net* net, SimulationState* simState,
                               lb::MacroscopicPropertyCache& propertyCache, reporting::Timers& timings,
                               distribn_t maximumRelativeDensityDifferenceAllowed = 0.05) :
          BroadcastPolicy(net, simState, SPREADFACTOR), mLatDat(latticeData), propertyCache(propertyCache),
          mSimState(simState), timings(timings), maximumRelativeDensityDifferenceAllowed(maximumRelativeDensityDifferenceAllowed),
          globalDensityTracker(nullptr)
      {
        // slots of children that do not exist must not move the extrema (IncompressibilityChecker.hpp:112-139)
        for (unsigned leaf = 0; leaf < SPREADFACTOR; ++leaf)
        {
          distribn_t* slot = childrenDensitiesSerialised + leaf * DensityTracker::DENSITY_TRACKER_SIZE;
          slot[DensityTracker::MIN_DENSITY] = slot[DensityTracker::MAX_DENSITY] = REFERENCE_DENSITY;
          slot[DensityTracker::MAX_VELOCITY_MAGNITUDE] = 0.0;
        }
        // from now on the engine of this Domain gathers the extrema inside its site kernel
        geometry::GpuPolicyFor(mLatDat).monitorRequested = true;
      }
      ~IncompressibilityChecker() noexcept override = default;

      void Report(reporting::Dict& dictionary) override
      {
        if (AreDensitiesAvailable() && !IsDensityDiffWithinRange())
        {
          reporting::Dict incomp = dictionary.AddSectionDictionary("DENSITIES");
          incomp.SetFormattedValue("ALLOWED", "%.1f%%", GetMaxRelativeDensityDifferenceAllowed() * 100);
          incomp.SetFormattedValue("ACTUAL", "%.1f%%", GetMaxRelativeDensityDifference() * 100);
        }
      }

      distribn_t GetGlobalSmallestDensity() const { HASSERT(AreDensitiesAvailable()); return (*globalDensityTracker)[DensityTracker::MIN_DENSITY]; }
      distribn_t GetGlobalLargestDensity() const { HASSERT(AreDensitiesAvailable()); return (*globalDensityTracker)[DensityTracker::MAX_DENSITY]; }
      double GetGlobalLargestVelocityMagnitude() const { HASSERT(AreDensitiesAvailable()); return (*globalDensityTracker)[DensityTracker::MAX_VELOCITY_MAGNITUDE]; }
      double GetMaxRelativeDensityDifference() const
      {
        const distribn_t spread = GetGlobalLargestDensity() - GetGlobalSmallestDensity();
        HASSERT(spread >= 0.0);
        return spread / REFERENCE_DENSITY;
      }
      double GetMaxRelativeDensityDifferenceAllowed() const { return maximumRelativeDensityDifferenceAllowed; }
      bool AreDensitiesAvailable() const { return globalDensityTracker != nullptr; }
      bool IsDensityDiffWithinRange() const { return GetMaxRelativeDensityDifference() < maximumRelativeDensityDifferenceAllowed; }

    protected:
      void ProgressFromChildren(unsigned long) override { this->ReceiveFromChildren(childrenDensitiesSerialised, DensityTracker::DENSITY_TRACKER_SIZE); }
      void ProgressFromParent(unsigned long) override { this->ReceiveFromParent(downwardsDensityTracker.GetDensitiesArray(), DensityTracker::DENSITY_TRACKER_SIZE); }
      void ProgressToChildren(unsigned long) override { this->SendToChildren(downwardsDensityTracker.GetDensitiesArray(), DensityTracker::DENSITY_TRACKER_SIZE); }
      void ProgressToParent(unsigned long) override { this->SendToParent(upwardsDensityTracker.GetDensitiesArray(), DensityTracker::DENSITY_TRACKER_SIZE); }

      void TopNodeAction() override { downwardsDensityTracker = upwardsDensityTracker; }

      void PostReceiveFromChildren(unsigned long) override
      {
        timings[hemelb::reporting::Timers::monitoring].Start();
        for (unsigned child = 0; child < SPREADFACTOR; ++child)
          upwardsDensityTracker.UpdateDensityTracker(DensityTracker(childrenDensitiesSerialised + child * DensityTracker::DENSITY_TRACKER_SIZE));
        timings[hemelb::reporting::Timers::monitoring].Stop();
      }

      // the local sites (IncompressibilityChecker.hpp:184-196): one device reduction instead of the host loop
      void PostSendToParent(unsigned long) override
      {
        timings[hemelb::reporting::Timers::monitoring].Start();
        if (hlb_gpu_t engine = geometry::GpuPolicyFor(mLatDat).engine)
        {
          double extrema[4];  // {smallest population, smallest density, largest density, largest |u|}
          geometry::FieldData::Check(hlb_gpu_monitor(engine, extrema));
          if (mLatDat->GetLocalFluidSiteCount() > 0)
            upwardsDensityTracker.Widen(extrema[1], extrema[2], extrema[3]);
        }
        timings[hemelb::reporting::Timers::monitoring].Stop();
      }

      void Effect() override { globalDensityTracker = &downwardsDensityTracker; }

    private:
      static const unsigned int SPREADFACTOR = 10u;  // the reference's tree width (IncompressibilityChecker.h:227)
      const geometry::Domain* mLatDat;
      lb::MacroscopicPropertyCache& propertyCache;  // (kept for the constructor's signature; not read)
      lb::SimulationState* mSimState;
      reporting::Timers& timings;
      distribn_t maximumRelativeDensityDifferenceAllowed;
      DensityTracker* globalDensityTracker;
      DensityTracker upwardsDensityTracker, downwardsDensityTracker;
      distribn_t childrenDensitiesSerialised[SPREADFACTOR * DensityTracker::DENSITY_TRACKER_SIZE];
  };
}
#endif
