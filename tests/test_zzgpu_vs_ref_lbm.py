"""The CUDA path straight against the reference's own ``lb::LBM`` (oracle/_ref/libhemelb_reflbm.so, prebuilt: it
travels with the snapshot) -- no restatement in between.  tests/test_oracle_vs_ref_lbm.py holds the oracle against
the same library on CPU; these close the triangle on the GPU for the headline bundle, the other lattices and link
rules, and R ranks with the halo staged through the host."""
import numpy as np
import pytest

import oracle as O
from hemelb_b200 import geometry as G
from hemelb_b200.domain import build_domains
from hemelb_b200.lbm import GpuLBM
from tests.cases import geometry, iolets_for, perturbed_equilibrium
from tests.test_host_lbm import DX, physical_dt, reference_tau

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(O.ref_lbm_lib() is None, reason="oracle/_ref/libhemelb_reflbm.so not built / shipped")]


@pytest.mark.parametrize("name,Q,kernel,wall,inlet,outlet", [
    ("cylinder", 19, "LBGK", "BFL", "NASH", "NASH"),   # the headline bundle
    ("tree", 19, "LBGK", "BFL", "NASH", "NASH"),       # configs[2]'s shape: 1 inlet, 4 outlets
    ("four_cube", 19, "LBGK", "SBB", "NASH", "NASH"),
    ("cylinder", 15, "LBGK", "SBB", "NASH", "NASH"),
    ("tree", 27, "LBGK", "BFL", "NASH", "NASH"),
    ("tree", 19, "LBGK", "GZS", "LADD", "NASH"),
    ("tree", 19, "MRT", "BFL", "LADD", "LADD"),
])
def test_gpu_equals_the_reference_lbm_after_seven_steps(name, Q, kernel, wall, inlet, outlet):
    geom = geometry(name)
    inlets, outlets = iolets_for(geom, inlet, outlet)
    dom = build_domains(geom, Q)[0]
    steps, dt = 7, physical_dt(0.8)
    f0 = perturbed_equilibrium(dom.N, Q, 0, O.lattice(Q)[1], seed=5)
    gpu = GpuLBM(dom, kernel, wall, inlet, outlet, tau=reference_tau(dt), inlets=inlets, outlets=outlets)
    gpu.set_f(f0)
    gpu.step(steps)
    ref, _ = O.ref_lbm_run(geom, Q, wall, inlet, inlets, outlets, dt, DX, steps, [dom.N], f0=[f0], kernel=kernel, outlet=outlet)
    got = gpu.get_f()[:dom.N * Q]
    assert np.abs(got - ref[0]).max() <= 1e-13
    assert np.array_equal(got, ref[0])


def test_gpu_ranks_with_a_host_staged_halo_equal_the_reference_lbm():
    """Three ranks' engines on one GPU, the halo moved as net::Net would (hlb_gpu_get_halo / set_halo), against the
    reference's own three-rank run over its FieldData::SendAndReceive / CopyReceived."""
    geom, Q, R = geometry("tree"), 19, 3
    rank = G.slab_decomposition(geom, R)
    inlets, outlets = iolets_for(geom, "NASH", "NASH")
    doms = build_domains(geom, Q, rank, R)
    steps, dt = 6, physical_dt(0.8)
    w = O.lattice(Q)[1]
    f0 = [perturbed_equilibrium(d.N, Q, d.totalSharedFs, w, seed=3 + r) for r, d in enumerate(doms)]
    gpus = [GpuLBM(d, "LBGK", "BFL", tau=reference_tau(dt), inlets=inlets, outlets=outlets) for d in doms]
    for g, f in zip(gpus, f0):
        g.set_f(f)
    for _ in range(steps):
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
    ref, _ = O.ref_lbm_run(geom, Q, "BFL", "NASH", inlets, outlets, dt, DX, steps, [d.N for d in doms], rank, R, f0=f0)
    for r, d in enumerate(doms):
        assert np.array_equal(gpus[r].get_f()[:d.N * Q], ref[r]), r


def _larger_golden_keys():
    from tests.test_golden import KEYS
    return [k for k in KEYS if "_R1_" in k and not k.startswith("four_cube_")]


@pytest.mark.parametrize("key", _larger_golden_keys())
def test_gpu_reproduces_the_larger_golden_vectors(key):
    """configs[3]'s bundle (MRT + GuoZhengShi + Ladd inlet + Nash outlets, tree) and configs[4]'s (D3Q27 TRT + BFL,
    sac) against vectors written by the reference's own streamers -- around its TRT::Collide, and with its one
    missing m_neq projection inserted (tests/golden/make_golden_trt.py)."""
    from tests.cases import anisotropic_f
    from tests.test_golden import GOLD, STEPS, _parse
    gname, R, Q, k, w, i, o = _parse(key)
    geom = geometry(gname)
    inlets, outlets = iolets_for(geom, i, o)
    dom = build_domains(geom, Q)[0]
    gpu = GpuLBM(dom, k, w, i, o, tau=float(GOLD[key + "_tau"][0]), inlets=inlets, outlets=outlets)
    gpu.set_f(anisotropic_f(dom.N, Q, 0))
    gpu.set_cache_mask(3)
    gpu.step(STEPS)
    f = gpu.get_f()[:dom.N * Q]
    assert np.abs(f - GOLD[key + "_f0"]).max() <= 1e-13
    assert np.array_equal(f, GOLD[key + "_f0"])
    assert np.array_equal(gpu.get_cache("density"), GOLD[key + "_rho0"])
