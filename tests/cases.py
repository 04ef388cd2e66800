"""Shared builders for the parity tests: small geometries, iolet records, initial data."""
from __future__ import annotations

import functools

import numpy as np

from hemelb_b200 import geometry as G
from hemelb_b200.capi import iolet_record
from hemelb_b200.lbm import prepare_boundary_objects

TOL_F = 1e-13  # north_star: per-step distributions agree to <= 1e-13 absolute


@functools.lru_cache(maxsize=None)
def geometry(name: str):
    if name == "four_cube":
        return G.four_cube()
    if name == "cylinder":
        return G.cylinder(5.3, 20)
    if name == "cylinder_long":
        return G.cylinder(4.2, 44)
    if name == "tree":
        return G.capsule_tree(3, 5.0, 16.0)
    if name == "sac":
        return G.sac(8, 3, 4, roughness=1.5)
    raise KeyError(name)


def iolets_for(geom, inlet_bc: str, outlet_bc: str):
    """Inlet / outlet records for a geometry's caps: cosine pressure for NASH, parabolic velocity
    for LADD (as a SimConfig would pair them)."""
    meta = geom.meta
    if meta.get("kind") == "four_cube":
        ins = [dict(position=(2.5, 2.5, 0.5), normal=(0, 0, 1), radius=2.5)]
        outs = [dict(position=(2.5, 2.5, 4.5), normal=(0, 0, -1), radius=2.5)]
    else:
        ins = [dict(position=tuple(p.position), normal=tuple(p.normal), radius=p.radius - 2) for p in meta["inlets"]]
        outs = [dict(position=tuple(p.position), normal=tuple(p.normal), radius=p.radius - 2) for p in meta["outlets"]]

    def rec(spec, bc, k, inlet):
        if bc == "LADD":
            return iolet_record(1, spec["normal"], spec["position"], radius=spec["radius"] + 1.0,
                                max_speed=0.02 if inlet else 0.015)
        return iolet_record(0, spec["normal"], spec["position"], radius=spec["radius"],
                            density_mean=1.01 if inlet else 0.995 - 0.0005 * k, density_amp=0.004,
                            phase=0.3 * k, period=64.0)

    inlets = [rec(s, inlet_bc, k, True) for k, s in enumerate(ins)]
    outlets = [rec(s, outlet_bc, k, False) for k, s in enumerate(outs)]
    prepare_boundary_objects(inlets, outlets)
    return inlets, outlets


def anisotropic_f(N, Q, S, site_offset=0):
    """LbTestsHelper.h:164-171: f_old[site][dir] = (dir+1)/10 + site/100."""
    f = np.zeros(N * Q + 1 + S)
    f[:N * Q] = (((np.arange(Q) + 1) / 10)[None, :] + ((np.arange(N) + site_offset) / 100)[:, None]).ravel()
    return f


def perturbed_equilibrium(N, Q, S, weights, seed=20261017):
    """rest equilibrium + seeded +-1e-3 perturbation."""
    rng = np.random.default_rng(seed)
    f = np.zeros(N * Q + 1 + S)
    f[:N * Q] = (weights[None, :] * (1.0 + 1e-3 * rng.uniform(-1, 1, (N, Q)))).ravel()
    return f


def valid_combo(Q, kernel, wall, inlet, outlet, need_ref=False):
    if kernel == "MRT" and Q == 27:
        return False
    if need_ref:
        if kernel == "TRT":
            return False  # TRT.h does not compile
        if kernel == "MRT" and (inlet, outlet) != ("LADD", "LADD"):
            return False  # MRT::CalculateFeq does not compile
        if kernel == "MRT" and wall == "GZS":
            return False  # reference reads an unset m_neq
    return True
