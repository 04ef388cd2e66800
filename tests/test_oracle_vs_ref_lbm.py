"""The oracle's WHOLE time step against the reference's own ``lb::LBM`` (oracle/_ref/libhemelb_reflbm.so).

oracle/ref_lbm_driver.cc builds, on R emulated ranks, what configuration/SimBuilder.h builds -- the reference's
unmodified geometry::Domain, geometry::FieldData, NeighbouringDataManager, lb::LBM<Traits> with the reference's CPU
streamers, lb::BoundaryValues and net::phased::StepManager (two phases, separated concerns) -- and steps it as
SimulationMaster::DoTimeStep does.  Held against it here, bit for bit after several steps: the restatement
(oracle/hemelb_oracle.cc), i.e. what every GPU parity test compares the product with.  This pins, beyond the
streamers one range at a time (tests/test_oracle_vs_ref.py): the phase schedule (lb.hpp:162-314), the halo exchange
(FieldData.cc:14-48), GuoZhengShi's site halo (NeighbouringDataManager.cc), the iolet densities over time
(BoundaryValues.cc, InOutLetCosine.cc, LBM::PrepareBoundaryObjects) and the initial condition
(InitialCondition.hpp).

One thing the reference does that nobody restates: NashZerothOrderPressure.h:39, LaddIolet.h:46 and GuoZhengShi.h:168
look the site's iolet up with ``GetLocalIolet(site.GetIoletId())`` -- a position in the rank's LOCAL iolet list
indexed with the GLOBAL id.  On the rank that owns the boundary-condition task (rank 0: every iolet is local) and on
any rank whose local list is 0..k-1 the two agree; elsewhere the reference takes another iolet's normal or reads past
the end of the list.  ``test_where_the_reference_indexes_its_local_iolet_list_with_a_global_id`` shows that this is
the only difference; the oracle and the GPU engine use the site's own iolet everywhere."""
import os

import numpy as np
import pytest

import oracle as O
from hemelb_b200.domain import build_domains
from hemelb_b200.lbm import BoundaryValues, SimulationState, prepare_boundary_objects
from tests.cases import geometry, iolets_for, perturbed_equilibrium
from tests.test_domain_vs_ref import decomposition
from tests.test_host_lbm import DX, physical_dt, reference_tau

pytestmark = pytest.mark.skipif(O.ref_lbm_lib() is None, reason="oracle/_ref/libhemelb_reflbm.so not built (needs the reference sources)")

IOLET_RANGES = (2, 3, 4, 5, 8, 9, 10, 11)  # inlet, outlet, inlet+wall, outlet+wall: mid-domain, then domain-edge


def local_lists_are_prefixes(geom, rank_of_site, R):
    """Per rank: does BoundaryValues' local iolet list (BoundaryValues.cc:21-43: the iolets with a site on this
    rank, in global order; all of them on rank 0, which owns the boundary-condition task) equal 0..k-1 for inlets
    and for outlets?  Then GetLocalIolet(global id) is the iolet meant."""
    ok = [True] * R
    if R == 1:
        return ok
    site_rank = np.asarray(rank_of_site)[np.asarray(geom.bsite)]  # (per boundary record; links in the .gmy's 26 directions)
    for r in range(1, R):
        for typ in (2, 3):  # cut by an inlet / an outlet (io/formats/geometry.h), the site's type in SiteData
            links = (np.asarray(geom.btype) == typ) & (site_rank == r)[:, None]
            here = np.unique(np.asarray(geom.biolet)[links])
            ok[r] = ok[r] and list(here) == list(range(here.size))
    return ok


def both(name, R, kind, Q, wall, inlet, steps, equilibrium=None, warmup=0, seed=11, kernel="LBGK", outlet="NASH"):
    geom = geometry(name)
    inlets, outlets = iolets_for(geom, inlet, outlet)
    if warmup:
        for rec in inlets + outlets:
            rec[13] = warmup
        prepare_boundary_objects(inlets, outlets)
    rank = decomposition(geom, R, kind)
    doms = build_domains(geom, Q, rank, R)
    dt = physical_dt(0.8)
    sim = O.OracleSim(O.OracleDomains(geom, Q, rank, R), kernel, wall, inlet, outlet, tau=reference_tau(dt),
                      inlets=inlets, outlets=outlets)
    f0 = None
    if equilibrium is None:
        w = O.lattice(Q)[1]
        f0 = [perturbed_equilibrium(d.N, Q, d.totalSharedFs, w, seed=seed + r) for r, d in enumerate(doms)]
        for r, f in enumerate(f0):
            sim.set_f(f, r)
            sim.set_f(f, r, 1)
    else:
        sim.set_equilibrium(*equilibrium)
    if steps:
        sim.step(steps)
    ref, dens = O.ref_lbm_run(geom, Q, wall, inlet, inlets, outlets, dt, DX, steps, [d.N for d in doms], rank, R,
                              f0=f0, equilibrium=equilibrium, kernel=kernel, outlet=outlet)
    mine = [sim.get_f(r)[:d.N * Q] for r, d in enumerate(doms)]
    return geom, rank, doms, mine, ref, dens, (inlets, outlets)


@pytest.mark.parametrize("name,R,kind,Q,wall,inlet", [
    ("four_cube", 1, None, 19, "SBB", "NASH"),
    ("cylinder", 1, None, 19, "BFL", "NASH"),
    ("cylinder", 2, "slab", 19, "BFL", "NASH"),
    ("cylinder", 3, "ragged", 19, "BFL", "NASH"),    # site-granular, what an optimised decomposition hands to Domain
    ("tree", 2, "slab", 19, "BFL", "NASH"),          # 1 inlet, 4 outlets
    ("tree", 3, "slab", 19, "BFL", "NASH"),
    ("cylinder", 2, "slab", 15, "SBB", "NASH"),
    ("tree", 2, "slab", 27, "BFL", "NASH"),
    ("cylinder", 2, "slab", 19, "BFL", "LADD"),
    ("cylinder", 3, "ragged", 19, "GZS", "NASH"),    # wall links that need f_old of sites on other ranks
    ("tree", 3, "slab", 19, "GZS", "NASH"),
    ("tree", 2, "slab", 19, "GZS", "LADD"),          # configs[3]'s link rules (with LBGK: MRT + GZS is UB in the reference)
])
def test_seven_steps_of_the_reference_lbm_bit_for_bit(name, R, kind, Q, wall, inlet):
    geom, rank, doms, mine, ref, _, _ = both(name, R, kind, Q, wall, inlet, steps=7)
    assert all(local_lists_are_prefixes(geom, rank, R)), "pick a decomposition the reference indexes correctly"
    for r in range(R):
        assert np.isfinite(ref[r]).all()
        assert np.array_equal(mine[r], ref[r]), (r, float(np.abs(mine[r] - ref[r]).max()))


@pytest.mark.parametrize("name,R,kind,Q,kernel,wall", [
    ("cylinder", 2, "slab", 19, "MRT", "BFL"),
    ("tree", 2, "slab", 19, "MRT", "BFL"),
    ("cylinder", 3, "ragged", 15, "MRT", "SBB"),
    ("tree", 2, "slab", 19, "LBGK", "BFL"),
])
def test_mrt_with_velocity_iolets_through_the_reference_lbm(name, R, kind, Q, kernel, wall):
    """MRT (MRT.h:56-105, d'Humieres bases) inside the whole step.  Velocity iolets on both sides: MRT with Nash
    iolets does not compile in the reference (MRT.h:73-86)."""
    geom, rank, doms, mine, ref, _, _ = both(name, R, kind, Q, wall, "LADD", steps=7, kernel=kernel, outlet="LADD")
    assert all(local_lists_are_prefixes(geom, rank, R))
    for r in range(R):
        assert np.isfinite(ref[r]).all()
        assert np.array_equal(mine[r], ref[r]), (r, float(np.abs(mine[r] - ref[r]).max()))


def _defect_case():
    """(run in a child process: the reference reads past the end of a vector here, which may also end in SIGSEGV)"""
    geom, rank, doms, mine, ref, _, _ = both("tree", 3, "basic", 19, "BFL", "NASH", steps=1)
    ok = local_lists_are_prefixes(geom, np.asarray(rank), 3)
    assert ok[0] and not all(ok)
    differing = 0
    for r, d in enumerate(doms):
        a, b = mine[r].reshape(d.N, 19), ref[r].reshape(d.N, 19)
        bad = np.nonzero((a != b).any(axis=1))[0]
        if ok[r]:
            assert bad.size == 0
            continue
        t = d.tables()
        bounds = np.cumsum(np.concatenate([t["mid"], t["edge"]]))
        assert set(np.searchsorted(bounds, bad, side="right")) <= set(IOLET_RANGES)
        differing += bad.size
    assert differing > 0  # (if the reference ever looks the iolet up by its global id this test says so)
    print("DEFECT-CONFINED-TO-IOLET-SITES", differing)


def test_where_the_reference_indexes_its_local_iolet_list_with_a_global_id():
    """The tree over BasicDecomposition on 3 ranks: ranks 1 and 2 hold outlets that are not 0..k-1.  After one step
    the oracle differs from the reference on those ranks at iolet-typed sites and nowhere else; rank 0 and every
    non-iolet site are identical.  The out-of-range read is undefined behaviour in the reference: on larger trees
    over 8 ranks it ends in SIGSEGV inside StreamerTypeFactory<NullLink, NashZerothOrderPressureLink>::StreamAndCollide
    about one run in three (seen while timing it), so the case runs in a child process and a
    child killed by SIGSEGV counts as the defect showing itself too."""
    import signal
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    run = subprocess.run([sys.executable, "-c", "import tests.test_oracle_vs_ref_lbm as t; t._defect_case()"], cwd=root,
                         capture_output=True, text=True, timeout=300)
    if run.returncode == -signal.SIGSEGV:
        return
    assert run.returncode == 0, run.stderr[-2000:]
    assert "DEFECT-CONFINED-TO-IOLET-SITES" in run.stdout


@pytest.mark.parametrize("name,R,kind,Q", [("cylinder", 1, None, 19), ("tree", 2, "slab", 19), ("cylinder", 2, "slab", 27),
                                           ("cylinder", 1, None, 15)])
def test_equilibrium_initial_condition_with_momentum(name, R, kind, Q):
    """lb::EquilibriumInitialCondition::SetFs (InitialCondition.hpp): f_old = f_new = f_eq(rho, m) -- then three
    steps from there."""
    eq = (1.02, (0.011, -0.02, 0.005))
    for steps in (0, 3):
        _, _, _, mine, ref, _, _ = both(name, R, kind, Q, "BFL", "NASH", steps=steps, equilibrium=eq)
        for r in range(R):
            assert np.array_equal(mine[r], ref[r]), (steps, r)
    assert len(np.unique(ref[0])) > Q  # (the flow has started: not the initial state any more)


@pytest.mark.parametrize("inlet", ["NASH", "LADD"])
def test_warm_up_ramp_of_the_iolets(inlet):
    """Warm-up of 5 steps: the cosine iolets go from the minimum simulation density to their target
    (InOutLetCosine.cc:26-43, LBM::PrepareBoundaryObjects lb.hpp:128-152), the parabolic inlet from rest to its
    speed (InOutLetParabolicVelocity.cc:23-43).  One rank: with more, the reference takes the minimum over each
    rank's local iolets (lb.hpp:132-140), the oracle's emulated ranks share one record list."""
    steps = 8
    _, _, _, mine, ref, dens, (inlets, outlets) = both("cylinder", 1, None, 19, "BFL", inlet, steps=steps, warmup=5)
    assert np.array_equal(mine[0], ref[0])
    # the product's host-side scalar provider (hemelb_b200/lbm.py) against the reference's BoundaryValues
    state = SimulationState()
    bi, bo = BoundaryValues(inlets, state), BoundaryValues(outlets, state)
    for s in range(steps):
        want = np.concatenate([bi.densities(), bo.densities()])
        assert np.array_equal(want, dens[s]), (s, want, dens[s])
        state.increment()
    assert dens[0, -1] != dens[6, -1]


@pytest.mark.parametrize("name", ["four_cube", "large_cylinder", "fedosov1c", "cyl_l100_r5"])
def test_reference_inputs_through_the_reference_lbm(name):
    """The reference's own .gmy / .xml fixtures (tests/golden/ref_inputs: voxelised walls with genuine cut distances
    and normals, the XML's step length, voxel size and pressures) through the reference's whole lb::LBM from an
    lb::EquilibriumInitialCondition at the initial pressure, one bundle per fixture as in
    tests/test_reference_inputs.py -- on one rank and on as many ranks as keep every rank's iolet list 0..k-1."""
    import xml.etree.ElementTree as ET
    from hemelb_b200 import geometry as G
    from tests.ref_inputs import HERE, load
    from tests.test_reference_inputs import POLICIES
    geom, tau, rho0, inlets, outlets = load(name)
    Q, kernel, wall, inlet, outlet = POLICIES[name]
    sim_xml = ET.parse(os.path.join(HERE, name + ".xml")).getroot().find("simulation")
    dt, dx = float(sim_xml.find("step_length").get("value")), float(sim_xml.find("voxel_size").get("value"))
    steps = 20 if geom.n_sites < 50000 else 6
    ran_on = []
    for R in (1, 2, 3):
        if R > 1 and geom.n_sites < 1000:
            continue
        rank = None if R == 1 else G.basic_decomposition(geom, R)
        if R > 1 and not all(local_lists_are_prefixes(geom, rank, R)):
            continue
        doms = build_domains(geom, Q, rank, R)
        sim = O.OracleSim(O.OracleDomains(geom, Q, rank, R), kernel, wall, inlet, outlet, tau=tau, inlets=inlets, outlets=outlets)
        sim.set_equilibrium(rho0)
        sim.step(steps)
        ref, _ = O.ref_lbm_run(geom, Q, wall, inlet, inlets, outlets, dt, dx, steps, [d.N for d in doms], rank, R,
                               equilibrium=(rho0, (0.0, 0.0, 0.0)), kernel=kernel, outlet=outlet)
        for r, d in enumerate(doms):
            assert np.array_equal(sim.get_f(r)[:d.N * Q], ref[r]), (name, R, r)
        ran_on.append(R)
    assert 1 in ran_on and (geom.n_sites < 1000 or len(ran_on) > 1), ran_on
