"""The oracle restatement against the UNMODIFIED reference headers (oracle/_ref): bit-identical
distributions and property caches on every policy bundle the reference can build.  Skipped where
oracle/_ref was not built (no /root/reference and no shipped .so)."""
import numpy as np
import pytest

import oracle as O
from hemelb_b200 import geometry as G
from tests.cases import anisotropic_f, geometry, iolets_for, valid_combo

pytestmark = pytest.mark.skipif(O.ref_lib() is None, reason="oracle/_ref not built")

DT, DX, RHO, ETA = 1e-4, 1e-4, 1000.0, 0.004  # LbmParameters(1e-4, 1e-4) => tau = 0.62

COMBOS = [(Q, k, w, i, o)
          for Q in (15, 19, 27) for k in ("LBGK", "MRT") for w in ("SBB", "BFL", "GZS")
          for (i, o) in (("NASH", "NASH"), ("LADD", "NASH"), ("LADD", "LADD"))
          if valid_combo(Q, k, w, i, o, need_ref=True)]


def _pair(geom, Q, kernel, wall, inlet, outlet, rank, R, sse3=False):
    inlets, outlets = iolets_for(geom, inlet, outlet)
    dom = O.OracleDomains(geom, Q, rank, R)
    T = [dom.tables(r) for r in range(R)]
    ref = O.RefSim(T, Q, kernel, wall, inlet, outlet, dt=DT, dx=DX, rho=RHO, eta=ETA, inlets=inlets, outlets=outlets,
                   sse3=sse3)
    sim = O.OracleSim(dom, kernel, wall, inlet, outlet, tau=ref.tau, inlets=inlets, outlets=outlets)
    for r in range(R):
        f = anisotropic_f(T[r]["N"], Q, T[r]["totalSharedFs"], site_offset=3 * r)
        sim.set_f(f, r)
        ref.set_f(f, r)
    return sim, ref, T


@pytest.mark.parametrize("Q,kernel,wall,inlet,outlet", COMBOS)
def test_four_cube_bit_identical(Q, kernel, wall, inlet, outlet):
    sim, ref, T = _pair(geometry("four_cube"), Q, kernel, wall, inlet, outlet, None, 1)
    sim.set_cache_mask(255)
    ref.set_cache_mask(255)
    sim.step(5)
    ref.step(5)
    n = T[0]["N"] * Q
    assert np.array_equal(sim.get_f()[:n], ref.get_f()[:n])
    for name in O.CACHE_BITS:
        assert np.array_equal(sim.get_cache(name), ref.get_cache(name)), name


@pytest.mark.parametrize("R", (1, 3))
@pytest.mark.parametrize("Q,kernel,wall,inlet,outlet", [
    (19, "LBGK", "BFL", "NASH", "NASH"), (19, "LBGK", "GZS", "LADD", "NASH"), (15, "LBGK", "SBB", "NASH", "NASH"),
    (27, "LBGK", "BFL", "LADD", "LADD"), (19, "MRT", "BFL", "LADD", "LADD"), (15, "MRT", "SBB", "LADD", "LADD")])
def test_cylinder_multi_rank_bit_identical(R, Q, kernel, wall, inlet, outlet):
    """Emulated ranks with the tables of the restated Domain builder: edge / mid phases, halo
    exchange, CopyReceived and PostStep all reproduce the reference's streamers exactly."""
    geom = geometry("cylinder")
    rank = None if R == 1 else G.slab_decomposition(geom, R)
    sim, ref, T = _pair(geom, Q, kernel, wall, inlet, outlet, rank, R)
    sim.step(8)
    ref.step(8)
    for r in range(R):
        n = T[r]["N"] * Q
        assert np.array_equal(sim.get_f(r)[:n], ref.get_f(r)[:n])


def test_scalar_vs_sse3_reference_paths_within_budget():
    """The reference's two code paths (scalar and the x86-64 default SSE3 intrinsics) differ in
    summation order; both must sit inside the 1e-13 per-step budget of each other."""
    if O.ref_lib(True) is None:
        pytest.skip("SSE3 build absent")
    geom = geometry("cylinder")
    a, ra, T = _pair(geom, 19, "LBGK", "BFL", "NASH", "NASH", None, 1)
    b, rb, _ = _pair(geom, 19, "LBGK", "BFL", "NASH", "NASH", None, 1, sse3=True)
    _, w, _ = O.lattice(19)
    from tests.cases import perturbed_equilibrium
    f = perturbed_equilibrium(T[0]["N"], 19, 0, w)
    ra.set_f(f)
    rb.set_f(f)
    ra.step(1)
    rb.step(1)
    n = T[0]["N"] * 19
    assert np.abs(ra.get_f()[:n] - rb.get_f()[:n]).max() <= 1e-13


@pytest.mark.parametrize("name,R,kind", [("cylinder", 4, "slab"), ("tree", 8, "basic"), ("tree", 5, "ragged")])
def test_threaded_reference_step_equals_serial(name, R, kind):
    """href_sim_step_mt (the timed CPU baseline: one thread per emulated rank for the whole call, neighbour flags
    instead of joins, one barrier per step) against the serial phase loop, bit for bit."""
    from tests.test_domain_vs_ref import decomposition
    geom = geometry(name)
    rank = decomposition(geom, R, kind)
    a, ra, T = _pair(geom, 19, "LBGK", "BFL", "NASH", "NASH", rank, R)
    b, rb, _ = _pair(geom, 19, "LBGK", "BFL", "NASH", "NASH", rank, R)
    ra.step(9)
    rb.step_mt(4)
    rb.step_mt(5)
    for r in range(R):
        assert np.array_equal(ra.get_f(r), rb.get_f(r))
