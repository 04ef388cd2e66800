"""Shared definitions of the extraction / checkpoint parity cases (golden generator, CPU and GPU tests)."""
from __future__ import annotations

import numpy as np

import oracle as O
from oracle import xtr as X
from hemelb_b200 import geometry as G
from tests.cases import anisotropic_f, geometry, iolets_for

DT, DX, RHO, ETA = 1e-4, 1e-4, 1000.0, 0.004   # LbmParameters(1e-4, 1e-4) => tau = 0.62
ORIGIN = (0.034, 0.001, 0.074)                  # DummyDataSource.h:27
REF_PRESSURE = 80.0

ALL_FIELDS = [("Pressure", "pressure", "float", (80.0,)), ("Velocity", "velocity", "float", ()),
              ("ShearStress", "shearstress", "float", ()), ("VonMises", "vonmisesstress", "double", ()),
              ("ShearRate", "shearrate", "float", ()), ("Stress", "stresstensor", "float", ()),
              ("Traction", "traction", "double", ()), ("TangTraction", "tangentialprojectiontraction", "float", ()),
              ("distributions", "distributions", "double", ()), ("Rank", "mpirank", "int32", ())]
CHECKPOINT_FIELDS = [("distributions", "distributions", "double", ())]


def _c(v):
    return tuple(np.array(ORIGIN) + DX * np.array(v))


# name -> case.  `writes`: steps to advance before each Write (time step = steps so far).
CASES = {
    "four_cube_all": dict(geom="four_cube", Q=15, R=1, wall="SBB", steps=4, fields=ALL_FIELDS, selector="whole", params=(),
                          frequency=2, writes=(0, 1, 1)),
    "four_cube_plane": dict(geom="four_cube", Q=15, R=1, wall="SBB", steps=4, fields=ALL_FIELDS[:2], selector="plane",
                            params=(*_c((2.5, 2.5, 2.5)), 0.0, 0.0, 1.0, 0.0), frequency=1, writes=(0,)),
    "cylinder_surface_r2": dict(geom="cylinder", Q=19, R=2, wall="BFL", steps=3, fields=ALL_FIELDS, selector="surface",
                                params=(), frequency=1, writes=(0, 2)),
    "cylinder_checkpoint_r2": dict(geom="cylinder", Q=19, R=2, wall="BFL", steps=3, fields=CHECKPOINT_FIELDS,
                                   selector="whole", params=(), frequency=1, writes=(0, 2, 2)),
}


def xfields(spec):
    return [X.Field(n, s, t, o) for (n, s, t, o) in spec]


def rank_of(case):
    geom = geometry(case["geom"])
    return None if case["R"] == 1 else G.slab_decomposition(geom, case["R"])


def make_sim(kind, case):
    """('ref' | 'oracle') simulation of a case, advanced `steps` steps with every cache on."""
    geom = geometry(case["geom"])
    Q, R = case["Q"], case["R"]
    inlets, outlets = iolets_for(geom, "NASH", "NASH")
    dom = O.OracleDomains(geom, Q, rank_of(case), R)
    T = [dom.tables(r) for r in range(R)]
    if kind == "ref":
        sim = O.RefSim(T, Q, "LBGK", case["wall"], "NASH", "NASH", dt=DT, dx=DX, rho=RHO, eta=ETA, inlets=inlets, outlets=outlets)
    else:
        sim = O.OracleSim(dom, "LBGK", case["wall"], "NASH", "NASH", tau=tau(), inlets=inlets, outlets=outlets)
        sim._domains_keepalive = dom
    for r in range(R):
        sim.set_f(initial_f(T[r], Q, r), r)
    sim.set_cache_mask(255)
    sim.step(case["steps"])
    return sim, T


def initial_f(t, Q, r):
    return anisotropic_f(t["N"], Q, t["totalSharedFs"], site_offset=3 * r) * 0.05


def tau():
    # LbmParameters.h:34-39 with eta / rho: tau = 0.5 + (dt * eta / rho) / (Cs2 * dx^2)
    return 0.5 + (DT * ETA / RHO) / ((1.0 / 3.0) * DX * DX)


def rank_data(sim, T, Q):
    out = []
    for r, t in enumerate(T):
        d = {"N": int(t["N"]), "globalCoords": t["globalCoords"], "wallMask": t["wallMask"], "wallNormal": t["wallNormal"],
             "f": sim.get_f(r)}
        for name in O.CACHE_BITS:
            d[name] = sim.get_cache(name, r)
        out.append(d)
    return out
