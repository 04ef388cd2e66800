"""GPU parity on the reference's own test inputs (tests/golden/ref_inputs: geometry the setup tool voxelised,
parameters and iolets from the accompanying XML): the CUDA path through the C ABI against the oracle,
bit for bit, one rank and three emulated ranks, every wall rule."""
import numpy as np
import pytest

import oracle as O
from hemelb_b200 import geometry as G
from hemelb_b200.domain import build_domains
from hemelb_b200.lbm import GpuLBM
from tests.ref_inputs import load
from tests.test_gpu_parity import _check

pytestmark = pytest.mark.gpu

CASES = [
    ("four_cube", 15, "LBGK", "SBB"), ("four_cube", 19, "MRT", "BFL"),
    ("large_cylinder", 19, "MRT", "BFL"), ("large_cylinder", 27, "LBGK", "BFL"), ("large_cylinder", 19, "TRT", "GZS"),
    ("fedosov1c", 19, "LBGK", "GZS"), ("fedosov1c", 15, "LBGK", "SBB"), ("fedosov1c", 27, "TRT", "BFL"),
    ("cyl_l100_r5", 19, "LBGK", "BFL"), ("cyl_l100_r5", 19, "MRT", "GZS"),
]


@pytest.mark.parametrize("name,Q,kernel,wall", CASES)
def test_reference_inputs_single_rank(name, Q, kernel, wall):
    """From the XML's initial pressure, with the XML's cosine-pressure iolets and tau, caches on."""
    geom, tau, rho0, inlets, outlets = load(name)
    dom = build_domains(geom, Q)[0]
    sim = O.OracleSim(O.OracleDomains(geom, Q), kernel, wall, "NASH", "NASH", tau=tau, inlets=inlets, outlets=outlets)
    gpu = GpuLBM(dom, kernel, wall, "NASH", "NASH", tau=tau, inlets=inlets, outlets=outlets)
    sim.set_equilibrium(rho0)
    gpu.set_equilibrium(rho0, (0.0, 0.0, 0.0))
    _check(gpu.get_f()[:dom.N * Q], sim.get_f()[:dom.N * Q], "initial condition")
    sim.set_cache_mask(255)
    gpu.set_cache_mask(255)
    steps = 60 if dom.N < 50000 else 25
    sim.step(steps)
    gpu.step(steps)
    a = gpu.get_f()[:dom.N * Q]
    assert np.isfinite(a).all()
    _check(a, sim.get_f()[:dom.N * Q], "%s f_old after %d steps" % (name, steps))
    for cache in O.CACHE_BITS:
        _check(gpu.get_cache(cache), sim.get_cache(cache), "cache " + cache)
    gpu.close()


@pytest.mark.parametrize("name,wall", [("cyl_l100_r5", "BFL"), ("fedosov1c", "SBB"), ("large_cylinder", "BFL")])
def test_reference_inputs_three_ranks_host_staged(name, wall):
    """The reference's BasicDecomposition over the fixture's blocks, three engines on one GPU, halo staged
    through hlb_gpu_get_halo / set_halo in the phase order."""
    geom, tau, rho0, inlets, outlets = load(name)
    Q, R = 19, 3
    rank = G.basic_decomposition(geom, R)
    doms = build_domains(geom, Q, rank, R)
    sim = O.OracleSim(O.OracleDomains(geom, Q, rank, R), "LBGK", wall, tau=tau, inlets=inlets, outlets=outlets)
    gpus = [GpuLBM(d, "LBGK", wall, tau=tau, inlets=inlets, outlets=outlets) for d in doms]
    sim.set_equilibrium(rho0)
    for g in gpus:
        g.set_equilibrium(rho0, (0.0, 0.0, 0.0))
    for _ in range(10):
        for g in gpus:
            g.request_comms()
            g.pre_send()
            g.pre_receive()
        sends = [g.get_halo(which=1) for g in gpus]
        for r, d in enumerate(doms):
            recv = np.zeros(d.totalSharedFs)
            for (p, cnt, first) in d.procs:
                op = doms[p].procs
                j = int(np.nonzero(op[:, 0] == r)[0][0])
                o_first = int(op[j, 2]) - (doms[p].N * Q + 1)
                m_first = int(first) - (d.N * Q + 1)
                recv[m_first:m_first + cnt] = sends[p][o_first:o_first + cnt]
            gpus[r].set_halo(recv, which=0)
        for g in gpus:
            g.post_receive()
            g.end_iteration()
            g.swap_old_and_new()
            g.state.increment()
    sim.step(10)
    for r, d in enumerate(doms):
        _check(gpus[r].get_f()[:d.N * Q], sim.get_f(r)[:d.N * Q], "%s rank %d" % (name, r))
    for g in gpus:
        g.close()
