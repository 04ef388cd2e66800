"""The reference's own test inputs (Code/tests/resources/{four_cube,large_cylinder,fedosov1c,cyl_l100_r5}.{gmy,xml},
copied to tests/golden/ref_inputs by tests/golden/make_reference_inputs.py) on the CPU side: the reader,
the XML numbers in lattice units, and the oracle against the unmodified reference streamers (oracle/_ref)
on geometry the setup tool voxelised -- cut distances and wall normals the synthetic generators never make."""
import numpy as np
import pytest

import oracle as O
from hemelb_b200 import geometry as G
from hemelb_b200.domain import build_domains
from tests.ref_inputs import load

# fluid sites held by each fixture (the .gmy headers; four_cube is the 4x4x4 box the reference's
# FourCubeBasedTestFixture runs on, Code/tests/helpers/FourCubeBasedTestFixture.h)
SITES = {"four_cube": 64, "large_cylinder": 5576, "fedosov1c": 15222, "cyl_l100_r5": 212400}
POLICIES = {  # one bundle per fixture, together covering every wall rule and both iolet rules
    "four_cube": (15, "LBGK", "SBB", "NASH", "NASH"),
    "large_cylinder": (27, "LBGK", "BFL", "NASH", "NASH"),   # MRT + Nash does not compile in the reference
    "fedosov1c": (19, "LBGK", "GZS", "NASH", "NASH"),
    "cyl_l100_r5": (19, "LBGK", "BFL", "NASH", "NASH"),
}


@pytest.mark.parametrize("name", sorted(SITES))
def test_reader_and_units(name):
    geom, tau, rho0, inlets, outlets = load(name)
    assert geom.n_sites == SITES[name]
    assert tau > 0.5 and np.isfinite(rho0)
    assert len(inlets) >= 1 and len(outlets) >= 1
    # cut distances of a voxelised wall are genuine fractions, not the synthetic generators' grid values
    dom = build_domains(geom, 19)[0]
    cut = dom.distance_to_wall()
    cut = cut[cut >= 0]
    assert cut.size and cut.min() >= 0.0 and cut.max() <= 1.0
    if name != "four_cube":
        assert np.unique(np.round(cut, 6)).size > 20


def test_four_cube_xml_in_lattice_units():
    """four_cube.xml: 80.1 / 80.0 mmHg, dt = 0.0857 s, dx = 0.01 m; tau = 0.5 + dt nu / (cs2 dx2)
    with nu = 4e-6 m2/s (Code/lb/LbmParameters.h:35)."""
    geom, tau, rho0, inlets, outlets = load("four_cube")
    assert tau == pytest.approx(0.5 + 3.0 * 4e-6 * 0.0857 / 1e-4, rel=1e-12)
    assert inlets[0][9] > outlets[0][9] > 1.0     # 80.1 mmHg > 80.0 mmHg > the 0 mmHg reference pressure
    assert tuple(inlets[0][1:4]) == (0.0, 0.0, 1.0) and tuple(outlets[0][1:4]) == (0.0, 0.0, -1.0)


@pytest.mark.skipif(O.ref_lib() is None, reason="oracle/_ref not built")
@pytest.mark.parametrize("name", sorted(SITES))
def test_oracle_equals_reference_on_reference_inputs(name):
    """XML parameters, equilibrium start at the initial pressure, the fixture's own iolets: the restatement and
    the reference's classes stay bit-identical, single rank and over three emulated ranks."""
    geom, tau, rho0, inlets, outlets = load(name)
    Q, kernel, wall, inlet, outlet = POLICIES[name]
    import xml.etree.ElementTree as ET, os
    from tests.ref_inputs import HERE
    sim_xml = ET.parse(os.path.join(HERE, name + ".xml")).getroot().find("simulation")
    dt, dx = float(sim_xml.find("step_length").get("value")), float(sim_xml.find("voxel_size").get("value"))
    steps = 40 if geom.n_sites < 50000 else 12
    for R in (1, 3):
        if R > 1 and geom.n_sites < 1000:
            continue
        rank = None if R == 1 else G.basic_decomposition(geom, R)
        dom = O.OracleDomains(geom, Q, rank, R)
        T = [dom.tables(r) for r in range(R)]
        ref = O.RefSim(T, Q, kernel, wall, inlet, outlet, dt=dt, dx=dx, inlets=inlets, outlets=outlets)
        assert ref.tau == pytest.approx(tau, rel=1e-14)
        sim = O.OracleSim(dom, kernel, wall, inlet, outlet, tau=ref.tau, inlets=inlets, outlets=outlets)
        sim.set_equilibrium(rho0)       # the restated InitialCondition (checked against the lattice's own in test_gpu_parity)
        for r in range(R):
            ref.set_f(sim.get_f(r), r)
        sim.step(steps)
        ref.step(steps)
        for r in range(R):
            n = T[r]["N"] * Q
            a, b = sim.get_f(r)[:n], ref.get_f(r)[:n]
            assert np.isfinite(a).all()
            assert np.array_equal(a, b), "%s rank %d/%d" % (name, r, R)
