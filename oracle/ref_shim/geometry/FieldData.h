// Shadow header (oracle/_ref build only): vector-backed stand-in for geometry::FieldData
// (Code/geometry/FieldData.h:117-212) -- f_old / f_new storage and Site views.
#pragma once
#include <vector>
#include "geometry/Domain.h"
namespace hemelb::geometry {
  class FieldData {
  public:
    using domain_type = Domain;
    Domain* dom = nullptr;
    std::vector<distribn_t> fOld, fNew;
    std::vector<LatticeForceVector> force;
    neighbouring::NeighbouringFieldData nfields;
    Domain& GetDomain() { return *dom; }
    const Domain& GetDomain() const { return *dom; }
    Site<FieldData> GetSite(site_t i) { return Site<FieldData>(i, *this); }
    Site<const FieldData> GetSite(site_t i) const { return Site<const FieldData>(i, *this); }
    distribn_t* GetFOld(site_t idx) { return &fOld[idx]; }
    const distribn_t* GetFOld(site_t idx) const { return &fOld[idx]; }
    distribn_t* GetFNew(site_t idx) { return &fNew[idx]; }
    const distribn_t* GetFNew(site_t idx) const { return &fNew[idx]; }
    const LatticeForceVector& GetForceAtSite(site_t i) const { return force[i]; }
    void SetForceAtSite(site_t i, const LatticeForceVector& f) { force[i] = f; }
    void AddToForceAtSite(site_t i, const LatticeForceVector& f) { force[i] += f; }
    neighbouring::NeighbouringFieldData& GetNeighbouringData() { return nfields; }
    const neighbouring::NeighbouringFieldData& GetNeighbouringData() const { return nfields; }
  };
}
