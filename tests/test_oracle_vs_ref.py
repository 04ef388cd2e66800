"""The oracle restatement against the UNMODIFIED reference headers (oracle/_ref): bit-identical
distributions and property caches on every policy bundle the reference can build.  Skipped where
oracle/_ref was not built (no /root/reference and no shipped .so)."""
import ctypes as C

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


@pytest.mark.parametrize("Q", (15, 19, 27))
def test_trt_collision_against_the_reference_file(Q):
    """TRT::Collide as lb/kernels/TRT.h:94-121 states it, compiled from the reference file itself.  The file has
    bit-rotted there (lb/Kernels.h includes it, no build instantiates it): it reaches the compiler through three substitutions made at build time --
    `iBar >= i` -> `iBar > i` (MakeOpposites counted the rest direction as a pair and overran its array), and
    `f_neq.f[` / `f_eq.f[` -> `f_neq[` / `f_eq[` (FVector became a std::array) -- see oracle/Makefile and
    oracle/ref_driver.cc.  Collide's arithmetic is untouched; f_eq and f_neq come from the pinned LBGK kernel, as
    TRT.h:62-92 computes them.  Bit-identical to the oracle's TRT, from which the CUDA kernel is checked."""
    L = O.ref_lib()
    if L is None:
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(Q)
    _, w, _ = O.lattice(Q)
    probe = np.ascontiguousarray(w)
    try:
        O.ref_collide(L, Q, "TRT", 100.0, 1.0, 1000.0, 0.004, probe)
    except ValueError:
        pytest.skip("oracle/_ref was built without the TRT path")
    worst = 0.0
    for tau0 in (0.51, 0.62, 0.8, 1.0, 1.4, 2.5):
        dt = (tau0 - 0.5) / 3.0 * 1000.0 / 0.004
        tau = float(L.href_tau(C.c_double(dt), C.c_double(1.0), C.c_double(1000.0), C.c_double(0.004)))
        for _ in range(40):
            f = w * (1.0 + 0.2 * rng.uniform(-1, 1, Q))
            a = O.collide(Q, "TRT", tau, f)
            b = O.ref_collide(L, Q, "TRT", dt, 1.0, 1000.0, 0.004, f)
            assert np.array_equal(a["fpost"], b["fpost"]), (tau0, float(np.abs(a["fpost"] - b["fpost"]).max()))
            assert np.array_equal(a["feq"], b["feq"]) and np.array_equal(a["fneq"], b["fneq"])
            worst = max(worst, float(np.abs(a["fpost"] - f).max()))
    assert worst > 1e-3  # (the collision did something)
    # and it is not LBGK in disguise: at tau != 1 the two differ
    f = w * (1.0 + 0.2 * rng.uniform(-1, 1, Q))
    assert not np.array_equal(O.collide(Q, "TRT", 0.62, f)["fpost"], O.collide(Q, "LBGK", 0.62, f)["fpost"])


TRT_COMBOS = [(Q, w, i, o) for Q in (15, 19, 27) for w in ("SBB", "BFL", "GZS")
              for (i, o) in (("NASH", "NASH"), ("LADD", "NASH"), ("LADD", "LADD"))]


def _trt_built():
    L = O.ref_lib()
    if L is None:
        return False
    try:
        O.ref_collide(L, 19, "TRT", 100.0, 1.0, 1000.0, 0.004, np.ascontiguousarray(O.lattice(19)[1]))
    except ValueError:
        return False
    return True


@pytest.mark.parametrize("Q,wall,inlet,outlet", TRT_COMBOS)
def test_trt_through_the_reference_streamers_four_cube(Q, wall, inlet, outlet):
    """The reference's TRT::Collide (see test_trt_collision_against_the_reference_file for how TRT.h is compiled)
    inside the reference's own streamers -- every wall and iolet rule, all three lattices, all eight caches --
    against the oracle's TRT: configs[4]'s kernel on the whole path, bit for bit."""
    if not _trt_built():
        pytest.skip("oracle/_ref was built without the TRT path")
    sim, ref, T = _pair(geometry("four_cube"), Q, "TRT", wall, inlet, outlet, None, 1)
    sim.set_cache_mask(255)
    ref.set_cache_mask(255)
    sim.step(5)
    ref.step(5)
    n = T[0]["N"] * Q
    assert np.array_equal(sim.get_f()[:n], ref.get_f()[:n])
    for name in O.CACHE_BITS:
        assert np.array_equal(sim.get_cache(name), ref.get_cache(name), equal_nan=True), name


@pytest.mark.parametrize("name,R,Q,wall,inlet,outlet", [
    ("sac", 1, 27, "BFL", "NASH", "NASH"),          # configs[4]: D3Q27 TRT + BFL on the rough-walled sac
    ("sac", 3, 27, "BFL", "NASH", "NASH"),
    ("cylinder", 3, 19, "GZS", "LADD", "NASH"),
    ("tree", 3, 19, "BFL", "NASH", "NASH"),
    ("cylinder", 3, 15, "SBB", "LADD", "LADD")])
def test_trt_through_the_reference_streamers_multi_rank(name, R, Q, wall, inlet, outlet):
    if not _trt_built():
        pytest.skip("oracle/_ref was built without the TRT path")
    geom = geometry(name)
    rank = None if R == 1 else G.slab_decomposition(geom, R)
    sim, ref, T = _pair(geom, Q, "TRT", wall, inlet, outlet, rank, R)
    sim.step(8)
    ref.step(8)
    for r in range(R):
        n = T[r]["N"] * Q
        assert np.array_equal(sim.get_f(r)[:n], ref.get_f(r)[:n]), r


@pytest.mark.parametrize("name,R,Q", [("four_cube", 1, 15), ("four_cube", 1, 19), ("cylinder", 1, 19), ("cylinder", 3, 19),
                                      ("tree", 3, 19), ("cylinder", 2, 15)])
def test_mrt_with_guo_zheng_shi_walls_is_the_reference_plus_one_projection(name, R, Q):
    """configs[3]'s collision and wall rule.  In the reference GuoZhengShi.h:279 collides the wall node's HydroVars
    without setting its m_neq, which MRT::Collide reads: undefined behaviour, so there is nothing to be identical
    to.  oracle/_ref/libhemelb_ref_mrtgzs.so is the same reference code with that ONE line inserted before the
    collision (m_neq = M f_neq of the wall node, by MRT's own ProjectVelsIntoMomentSpace; oracle/Makefile).  The
    oracle equals it bit for bit -- so the missing projection is the only thing the oracle (and the CUDA kernel
    checked against it) adds to the reference's text: the extrapolated wall node, its equilibrium, the moment-space
    collision and the streaming are the reference's arithmetic.  Velocity iolets on both sides (MRT with Nash
    iolets does not compile in the reference)."""
    if O.ref_lib("mrtgzs") is None:
        pytest.skip("oracle/_ref/libhemelb_ref_mrtgzs.so not built")
    geom = geometry(name)
    rank = None if R == 1 else G.slab_decomposition(geom, R)
    sim, ref, T = _pair(geom, Q, "MRT", "GZS", "LADD", "LADD", rank, R, sse3="mrtgzs")
    sim.set_cache_mask(255)
    ref.set_cache_mask(255)
    sim.step(8)
    ref.step(8)
    for r in range(R):
        n = T[r]["N"] * Q
        a, b = sim.get_f(r)[:n], ref.get_f(r)[:n]
        assert np.isfinite(b).all()
        assert np.array_equal(a, b), (r, float(np.abs(a - b).max()))
        for cname in O.CACHE_BITS:
            assert np.array_equal(sim.get_cache(cname, r), ref.get_cache(cname, r), equal_nan=True), cname


@pytest.mark.parametrize("name,R,Q,wall,inlet,outlet", [
    ("tree", 1, 19, "GZS", "LADD", "NASH"),        # configs[3]'s exact bundle: MRT + GuoZhengShi, Ladd inlet, Nash outlets
    ("tree", 3, 19, "GZS", "LADD", "NASH"),
    ("four_cube", 1, 19, "GZS", "LADD", "NASH"),
    ("four_cube", 1, 15, "BFL", "NASH", "NASH"),
    ("cylinder", 3, 19, "BFL", "NASH", "NASH"),
    ("cylinder", 2, 15, "SBB", "LADD", "NASH"),
    ("four_cube", 1, 19, "GZS", "NASH", "NASH")])
def test_mrt_with_nash_iolets_through_the_reference(name, R, Q, wall, inlet, outlet):
    """MRT with Nash iolets does not compile in the reference: MRT::CalculateFeq (MRT.h:73-86), which the Nash link
    calls, kept the `.f` member FVector lost.  In oracle/_ref/libhemelb_ref_mrtgzs.so four substitutions bring it to
    the form of the CalculateDensityMomentumFeq right above it (oracle/Makefile); together with the inserted m_neq
    projection (previous test) that makes configs[3]'s own bundle buildable from the reference's text -- and the
    oracle equals it bit for bit, all eight caches included."""
    if O.ref_lib("mrtgzs") is None:
        pytest.skip("oracle/_ref/libhemelb_ref_mrtgzs.so not built")
    geom = geometry(name)
    rank = None if R == 1 else G.slab_decomposition(geom, R)
    try:
        sim, ref, T = _pair(geom, Q, "MRT", wall, inlet, outlet, rank, R, sse3="mrtgzs")
    except Exception:
        pytest.skip("oracle/_ref/libhemelb_ref_mrtgzs.so was built without the MRT + Nash bundles")
    sim.set_cache_mask(255)
    ref.set_cache_mask(255)
    sim.step(8)
    ref.step(8)
    for r in range(R):
        n = T[r]["N"] * Q
        a, b = sim.get_f(r)[:n], ref.get_f(r)[:n]
        assert np.isfinite(b).all()
        assert np.array_equal(a, b), (r, float(np.abs(a - b).max()))
        for cname in O.CACHE_BITS:
            assert np.array_equal(sim.get_cache(cname, r), ref.get_cache(cname, r), equal_nan=True), cname
