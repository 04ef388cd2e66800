"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Bar (BASELINE.json north_star): per-step distributions <= 1e-13 absolute.  The kernels are built
with -fmad=false and follow the reference's scalar operation order, so in practice the match is
bit-exact; EXACT=True asserts that too.
"""
import numpy as np
import pytest

import oracle as O
from hemelb_b200 import geometry as G
from hemelb_b200.domain import build_domains
from hemelb_b200.lbm import GpuLBM
from tests.cases import TOL_F, anisotropic_f, geometry, iolets_for, perturbed_equilibrium, valid_combo

pytestmark = pytest.mark.gpu
EXACT = True


def _check(a, b, what):
    # NaN where the reference gives NaN (e.g. the wall shear stress of a wall site at rest: the square root
    # of a rounding-negative difference, Lattice.h:652-700) counts as equal
    nan = np.isnan(b)
    assert np.array_equal(np.isnan(a), nan), "%s: NaN at other places than the oracle" % what
    err = np.abs(a[~nan] - b[~nan]).max() if (~nan).any() else 0.0
    assert err <= TOL_F, "%s: max abs err %g" % (what, err)
    if EXACT:
        assert np.array_equal(a, b, equal_nan=True), "%s: not bit-identical (max abs err %g)" % (what, err)


def _run_pair(geom, Q, kernel, wall, inlet, outlet, steps, tau=0.62, mask=255, init="anisotropic"):
    inlets, outlets = iolets_for(geom, inlet, outlet)
    dom = build_domains(geom, Q)[0]
    odom = O.OracleDomains(geom, Q)
    sim = O.OracleSim(odom, kernel, wall, inlet, outlet, tau=tau, inlets=inlets, outlets=outlets)
    gpu = GpuLBM(dom, kernel, wall, inlet, outlet, tau=tau, inlets=inlets, outlets=outlets)
    if init == "anisotropic":
        f0 = anisotropic_f(dom.N, Q, 0)
    else:
        _, w, _ = O.lattice(Q)
        f0 = perturbed_equilibrium(dom.N, Q, 0, w)
    sim.set_f(f0)
    gpu.set_f(f0)
    sim.set_cache_mask(mask)
    gpu.set_cache_mask(mask)
    sim.step(steps)
    gpu.step(steps)
    return sim, gpu, dom


COMBOS = [(Q, k, w, i, o)
          for Q in (15, 19, 27) for k in ("LBGK", "MRT", "TRT") for w in ("SBB", "BFL", "GZS")
          for (i, o) in (("NASH", "NASH"), ("LADD", "NASH"), ("LADD", "LADD"))
          if valid_combo(Q, k, w, i, o)]


@pytest.mark.parametrize("Q,kernel,wall,inlet,outlet", COMBOS)
def test_four_cube_all_policies(Q, kernel, wall, inlet, outlet):
    """configs[0] geometry, every policy bundle, 5 steps from LbTestsHelper's anisotropic data."""
    sim, gpu, dom = _run_pair(geometry("four_cube"), Q, kernel, wall, inlet, outlet, 5)
    _check(gpu.get_f()[:dom.N * Q], sim.get_f()[:dom.N * Q], "f_old after 5 steps")
    for name in O.CACHE_BITS:
        _check(gpu.get_cache(name), sim.get_cache(name), "cache " + name)


@pytest.mark.parametrize("geom_name", ["cylinder", "tree", "sac"])
@pytest.mark.parametrize("Q,kernel,wall,inlet,outlet", [
    (19, "LBGK", "BFL", "NASH", "NASH"),   # configs[1], configs[2]
    (19, "MRT", "GZS", "LADD", "NASH"),    # configs[3]
    (27, "TRT", "BFL", "NASH", "NASH"),    # configs[4]
    (15, "LBGK", "SBB", "NASH", "NASH"),   # configs[0] policies
    (19, "LBGK", "GZS", "LADD", "LADD"),
    (15, "MRT", "BFL", "LADD", "LADD"),
])
def test_baseline_configs_small(geom_name, Q, kernel, wall, inlet, outlet):
    sim, gpu, dom = _run_pair(geometry(geom_name), Q, kernel, wall, inlet, outlet, 20, tau=0.8, init="equilibrium")
    _check(gpu.get_f()[:dom.N * Q], sim.get_f()[:dom.N * Q], "f_old after 20 steps")
    _check(gpu.get_cache("density"), sim.get_cache("density"), "density")
    _check(gpu.get_cache("velocity"), sim.get_cache("velocity"), "velocity")


def test_thousand_steps_density_velocity():
    """north_star: density / velocity agree to <= 1e-10 relative after 1000 steps."""
    geom = geometry("cylinder")
    sim, gpu, dom = _run_pair(geom, 19, "LBGK", "BFL", "NASH", "NASH", 1000, tau=0.8, mask=3, init="equilibrium")
    rho_g, rho_o = gpu.get_cache("density"), sim.get_cache("density")
    u_g, u_o = gpu.get_cache("velocity"), sim.get_cache("velocity")
    assert np.abs(rho_g / rho_o - 1).max() <= 1e-10
    scale = np.abs(u_o).max()
    assert scale > 1e-6  # the flow actually developed
    assert np.abs(u_g - u_o).max() / scale <= 1e-10
    if EXACT:
        assert np.array_equal(gpu.get_f()[:dom.N * 19], sim.get_f()[:dom.N * 19])


@pytest.mark.parametrize("slot", range(6))
def test_single_ranges_like_streamer_tests(slot):
    """StreamerTests.cc pattern: one streamer over one site range, then PostStep, on the four-cube
    fixture (D3Q15 LBGK; BFL walls so PostStep does work)."""
    geom, Q = geometry("four_cube"), 15
    inlets, outlets = iolets_for(geom, "NASH", "NASH")
    dom = build_domains(geom, Q)[0]
    odom = O.OracleDomains(geom, Q)
    sim = O.OracleSim(odom, "LBGK", "BFL", tau=0.62, inlets=inlets, outlets=outlets)
    gpu = GpuLBM(dom, "LBGK", "BFL", tau=0.62, inlets=inlets, outlets=outlets)
    f0 = anisotropic_f(dom.N, Q, 0)
    for s in (sim, gpu):
        s.set_f(f0, which=0)
        s.set_f(np.full_like(f0, -7.0), which=1)
    first = int(dom.mid[:slot].sum())
    count = int(dom.mid[slot])
    assert count > 0
    # a sub-range first, then the rest of the range
    half = count // 2
    for s in (sim, gpu):
        s.stream_and_collide(slot, first, half)
        s.stream_and_collide(slot, first + half, count - half)
        s.post_step(slot, first, count)
    _check(gpu.get_f(which=1)[:dom.N * Q], sim.get_f(which=1)[:dom.N * Q], "f_new")


def test_phase_api_equals_whole_step():
    """RequestComms / PreSend / PreReceive / PostReceive / EndIteration driven from the host equal
    hlb_gpu_step."""
    geom, Q = geometry("cylinder"), 19
    inlets, outlets = iolets_for(geom, "NASH", "NASH")
    dom = build_domains(geom, Q)[0]
    a = GpuLBM(dom, "LBGK", "BFL", tau=0.8, inlets=inlets, outlets=outlets)
    b = GpuLBM(dom, "LBGK", "BFL", tau=0.8, inlets=inlets, outlets=outlets)
    f0 = anisotropic_f(dom.N, Q, 0)
    a.set_f(f0)
    b.set_f(f0)
    for _ in range(7):
        a.do_time_step()
    b.step(7)
    assert np.array_equal(a.get_f(), b.get_f())


def test_tables_round_trip_on_device():
    """neighbourIndices uploaded in reference form come back bit-identical."""
    geom = geometry("tree")
    for Q in (15, 19, 27):
        dom = build_domains(geom, Q)[0]
        inlets, outlets = iolets_for(geom, "NASH", "NASH")
        gpu = GpuLBM(dom, inlets=inlets, outlets=outlets)
        assert np.array_equal(gpu.get_neighbour_indices(), dom.neighbour_indices())


@pytest.mark.parametrize("R,decomp", [(2, "slab"), (3, "slab"), (4, "basic")])
def test_multi_rank_host_staged_halo(R, decomp):
    """R emulated ranks on one GPU, halo moved through hlb_gpu_get_halo / set_halo (the host-staged
    exchange a reference build can keep using net::Net for): matches the oracle's R-rank run."""
    geom, Q = geometry("cylinder_long"), 19
    rank = G.slab_decomposition(geom, R) if decomp == "slab" else G.basic_decomposition(geom, R)
    inlets, outlets = iolets_for(geom, "NASH", "NASH")
    doms = build_domains(geom, Q, rank, R)
    odom = O.OracleDomains(geom, Q, rank, R)
    sim = O.OracleSim(odom, "LBGK", "BFL", tau=0.8, inlets=inlets, outlets=outlets)
    gpus = [GpuLBM(d, "LBGK", "BFL", tau=0.8, inlets=inlets, outlets=outlets) for d in doms]
    for r, d in enumerate(doms):
        f0 = anisotropic_f(d.N, Q, d.totalSharedFs, site_offset=7 * r)
        sim.set_f(f0, r)
        gpus[r].set_f(f0)
    for _ in range(6):
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
    sim.step(6)
    for r, d in enumerate(doms):
        _check(gpus[r].get_f()[:d.N * Q], sim.get_f(r)[:d.N * Q], "rank %d f_old" % r)


@pytest.mark.parametrize("geom_name,Q,kernel,wall,inlet,outlet", [
    ("tree", 19, "LBGK", "BFL", "NASH", "NASH"), ("sac", 27, "TRT", "GZS", "LADD", "NASH"),
    ("cylinder", 15, "MRT", "SBB", "LADD", "LADD")])
def test_internal_renumbering_is_invisible(geom_name, Q, kernel, wall, inlet, outlet):
    """cfg.reorder renumbers sites on the device only: distributions, caches, sub-range calls and
    the neighbour table read back in reference form are identical with and without it."""
    geom = geometry(geom_name)
    inlets, outlets = iolets_for(geom, inlet, outlet)
    dom = build_domains(geom, Q)[0]
    a = GpuLBM(dom, kernel, wall, inlet, outlet, tau=0.7, inlets=inlets, outlets=outlets, reorder=True)
    b = GpuLBM(dom, kernel, wall, inlet, outlet, tau=0.7, inlets=inlets, outlets=outlets, reorder=False)
    assert np.array_equal(a.get_neighbour_indices(), dom.neighbour_indices())
    assert np.array_equal(b.get_neighbour_indices(), dom.neighbour_indices())
    f0 = anisotropic_f(dom.N, Q, 0)
    for g in (a, b):
        g.set_f(f0)
        assert np.array_equal(g.get_f(), f0)
        g.set_cache_mask(255)
        g.step(4)
        # a ragged sub-range of the wall streamer on top (goes through the site-list path)
        first, count = int(dom.mid[0]) + 3, int(dom.mid[1]) - 7
        g.stream_and_collide(1, first, count)
        g.post_step(1, first, count)
    assert np.array_equal(a.get_f(), b.get_f())
    assert np.array_equal(a.get_f(which=1), b.get_f(which=1))
    for name in O.CACHE_BITS:
        assert np.array_equal(a.get_cache(name), b.get_cache(name)), name


@pytest.mark.parametrize("kernel,inlet", [("LBGK", "NASH"), ("MRT", "LADD")])
def test_multi_rank_gzs_site_halo_host_staged(kernel, inlet):
    """GZS links that extrapolate from a site on another rank: the phase-0 site halo
    (NeighbouringDataManager) staged through the host between 3 emulated ranks on one GPU."""
    geom, Q, R = geometry("cylinder_long"), 19, 3
    rank = G.slab_decomposition(geom, R)
    inlets, outlets = iolets_for(geom, inlet, "NASH")
    doms = build_domains(geom, Q, rank, R)
    odom = O.OracleDomains(geom, Q, rank, R)
    sim = O.OracleSim(odom, kernel, "GZS", inlet, "NASH", tau=0.8, inlets=inlets, outlets=outlets)
    gpus = [GpuLBM(d, kernel, "GZS", inlet, "NASH", tau=0.8, inlets=inlets, outlets=outlets) for d in doms]
    assert sum(g.gzs_need.shape[0] for g in gpus) > 0
    # several links extrapolate from the same remote site: they share a ghost row
    assert sum(g.gzs_row_owner.size for g in gpus) < sum(g.gzs_need.shape[0] for g in gpus)
    for r, d in enumerate(doms):
        f0 = anisotropic_f(d.N, Q, d.totalSharedFs, site_offset=5 * r)
        sim.set_f(f0, r)
        gpus[r].set_f(f0)
    for _ in range(5):
        for g in gpus:
            g.exchange_site_halo()  # packs the serve rows
        sends = [g.get_gzs_send() for g in gpus]
        for r, g in enumerate(gpus):
            rows = np.zeros((g.gzs_row_owner.size, Q))
            for p in range(R):
                mine = np.nonzero(g.gzs_row_owner == p)[0]
                theirs = np.nonzero(gpus[p].gzs_serve[:, 0] == r)[0]
                assert mine.size == theirs.size
                rows[mine] = sends[p][theirs]
            g.set_gzs_ghost(rows)
        for g in gpus:
            g.request_comms()
            g.pre_send()
            g.pre_receive()
        halo = [g.get_halo(which=1) for g in gpus]
        for r, d in enumerate(doms):
            recv = np.zeros(d.totalSharedFs)
            for (p, cnt, first) in d.procs:
                op = doms[p].procs
                j = int(np.nonzero(op[:, 0] == r)[0][0])
                o_first = int(op[j, 2]) - (doms[p].N * Q + 1)
                m_first = int(first) - (d.N * Q + 1)
                recv[m_first:m_first + cnt] = halo[p][o_first:o_first + cnt]
            gpus[r].set_halo(recv, which=0)
        for g in gpus:
            g.post_receive()
            g.swap_old_and_new()
            g.state.increment()
    sim.step(5)
    for r, d in enumerate(doms):
        _check(gpus[r].get_f()[:d.N * Q], sim.get_f(r)[:d.N * Q], "rank %d f_old" % r)


def test_error_paths():
    from hemelb_b200.capi import HlbError
    geom = geometry("four_cube")
    dom = build_domains(geom, 27)[0]
    with pytest.raises(HlbError, match="No MRT basis for D3Q27"):
        GpuLBM(dom, "MRT")
    dom = build_domains(geom, 15)[0]
    with pytest.raises(HlbError, match="outside the iolet table"):
        GpuLBM(dom)  # iolet sites, but no iolet records: the kernels would index an empty table
    inlets, outlets = iolets_for(geom, "NASH", "NASH")
    gpu = GpuLBM(dom, inlets=inlets, outlets=outlets)
    with pytest.raises(HlbError, match="outside the local fluid sites"):
        gpu.stream_and_collide(0, 0, dom.N + 1)
    with pytest.raises(HlbError, match="bulk-typed"):
        gpu.stream_and_collide(1, 0, 4)


def test_monitor_matches_numpy():
    geom, Q = geometry("cylinder"), 19
    dom = build_domains(geom, Q)[0]
    inlets, outlets = iolets_for(geom, "NASH", "NASH")
    gpu = GpuLBM(dom, inlets=inlets, outlets=outlets)
    f0 = anisotropic_f(dom.N, Q, 0)
    gpu.set_f(f0)
    m = gpu.monitor()
    assert gpu.monitor_global() == m  # one rank: nothing to reduce
    f = f0[:dom.N * Q].reshape(dom.N, Q)
    rho = f.sum(1)
    assert m["min_f"] == f.min()
    assert abs(m["min_density"] - rho.min()) < 1e-12 and abs(m["max_density"] - rho.max()) < 1e-12
    # fused variant: gathered by the collide kernels from the distributions entering the step
    gpu.set_cache_mask(256)
    gpu.step(1)
    m2 = gpu.monitor()
    assert m2["min_f"] == f.min()
    assert abs(m2["min_density"] - rho.min()) < 1e-12 and abs(m2["max_density"] - rho.max()) < 1e-12
    assert abs(m2["max_speed"] - m["max_speed"]) < 1e-12
    # and it re-arms: the next step reports the state that entered it, equal to a stand-alone pass
    f1 = gpu.get_f()[:dom.N * Q].reshape(dom.N, Q)
    gpu.step(1)
    m3 = gpu.monitor()
    rho1 = f1.sum(1)
    assert m3["min_f"] == f1.min()
    assert abs(m3["min_density"] - rho1.min()) < 1e-12 and abs(m3["max_density"] - rho1.max()) < 1e-12
    c = O.lattice(Q)[0].astype(np.float64)
    u1 = (f1 @ c) / rho1[:, None]
    assert abs(m3["max_speed"] - np.sqrt((u1 * u1).sum(1)).max()) < 1e-12
    assert m3["min_f"] != m2["min_f"]
    # the stand-alone pass (monitor not fused) over the current state
    gpu.set_cache_mask(0)
    gpu.step(0)
    f2 = gpu.get_f()[:dom.N * Q].reshape(dom.N, Q)
    m4 = gpu.monitor()
    assert m4["min_f"] == f2.min()
    assert abs(m4["min_density"] - f2.sum(1).min()) < 1e-12


def test_monitor_read_back_in_two_halves():
    """hlb_gpu_monitor_begin / _end: the values of hlb_gpu_monitor, collected after the next step has been
    issued; one read-back outstanding at a time."""
    from hemelb_b200.capi import HlbError
    geom, Q = geometry("cylinder"), 19
    dom = build_domains(geom, Q)[0]
    inlets, outlets = iolets_for(geom, "NASH", "NASH")
    a = GpuLBM(dom, inlets=inlets, outlets=outlets)
    b = GpuLBM(dom, inlets=inlets, outlets=outlets)
    f0 = anisotropic_f(dom.N, Q, 0)
    for g in (a, b):
        g.set_f(f0)
        g.set_cache_mask(256)
    blocking, lagged = [], []
    for _ in range(4):
        a.step(1)
        blocking.append(a.monitor())
    pending = False
    for _ in range(4):
        b.step(1)
        if pending:
            lagged.append(b.monitor_end())
        b.monitor_begin()
        pending = True
    lagged.append(b.monitor_end())
    assert lagged == blocking
    assert np.array_equal(a.get_f(), b.get_f())
    # nothing gathered since the last read: the one-pass form answers, through the same pair of calls
    b.set_cache_mask(0)
    b.monitor_begin()
    assert b.monitor_end() == a.monitor()
    with pytest.raises(HlbError, match="without hlb_gpu_monitor_begin"):
        b.monitor_end()
    b.set_cache_mask(256)
    b.step(1)
    b.monitor_begin()
    with pytest.raises(HlbError, match="has not been collected"):
        b.monitor_begin()
    with pytest.raises(HlbError, match="outstanding"):
        b.monitor()
    b.monitor_end()


@pytest.mark.parametrize("Q", (15, 19, 27))
def test_equilibrium_initial_condition_matches_the_oracle(Q):
    """EquilibriumInitialCondition::SetFs (Code/lb/InitialCondition.hpp:40-52): f_old = f_new =
    f_eq(rho0, m0) everywhere -- with a non-zero momentum, bit for bit against the oracle, and the
    trajectories that start there stay identical."""
    geom = geometry("cylinder")
    dom = build_domains(geom, Q)[0]
    inlets, outlets = iolets_for(geom, "NASH", "NASH")
    gpu = GpuLBM(dom, "LBGK", "BFL", tau=0.8, inlets=inlets, outlets=outlets)
    sim = O.OracleSim(O.OracleDomains(geom, Q), "LBGK", "BFL", tau=0.8, inlets=inlets, outlets=outlets)
    rho0, m0 = 1.0125, (0.011, -0.007, 0.023)
    gpu.set_equilibrium(rho0, m0)
    sim.set_equilibrium(rho0, m0)
    for which in (0, 1):
        assert np.array_equal(gpu.get_f(which)[:dom.N * Q], sim.get_f(0, which)[:dom.N * Q]), which
    f = gpu.get_f()[:dom.N * Q].reshape(dom.N, Q)
    assert np.all(f == f[0]) and abs(f[0].sum() - rho0) < 1e-14
    gpu.step(4)
    sim.step(4)
    assert np.array_equal(gpu.get_f()[:dom.N * Q], sim.get_f()[:dom.N * Q])


@pytest.mark.parametrize("geom_name,Q,kernel,wall,inlet,outlet", [
    ("tree", 19, "LBGK", "BFL", "NASH", "NASH"), ("tree", 19, "MRT", "GZS", "LADD", "NASH"),
    ("sac", 27, "TRT", "BFL", "NASH", "NASH"), ("cylinder", 15, "LBGK", "SBB", "LADD", "NASH")])
def test_tma_staged_site_kernel_is_bit_identical(monkeypatch, geom_name, Q, kernel, wall, inlet, outlet):
    """HLB_TMA=1: the mid-domain part through the persistent, TMA-staged form of the site kernel (2-D bulk
    tensor copies into per-warp shared-memory stages, mbarrier completion) -- opt-in because it is
    slower (profiles/r02_tma_experiments.md), and bit-identical to the direct form and the oracle."""
    geom = geometry(geom_name)
    dom = build_domains(geom, Q)[0]
    inlets, outlets = iolets_for(geom, inlet, outlet)
    f0 = anisotropic_f(dom.N, Q, 0)
    direct = GpuLBM(dom, kernel, wall, inlet, outlet, tau=0.8, inlets=inlets, outlets=outlets)
    monkeypatch.setenv("HLB_TMA", "1")
    staged = GpuLBM(dom, kernel, wall, inlet, outlet, tau=0.8, inlets=inlets, outlets=outlets)
    monkeypatch.delenv("HLB_TMA")
    sim = O.OracleSim(O.OracleDomains(geom, Q), kernel, wall, inlet, outlet, tau=0.8, inlets=inlets, outlets=outlets)
    for g in (direct, staged, sim):
        g.set_f(f0)
        g.set_cache_mask(255)
    for g in (direct, staged, sim):
        g.step(5)
    assert np.array_equal(staged.get_f(), direct.get_f())
    assert np.abs(staged.get_f()[:dom.N * Q] - sim.get_f()[:dom.N * Q]).max() <= TOL_F
    for name in O.CACHE_BITS:
        assert np.array_equal(staged.get_cache(name), direct.get_cache(name)), name


@pytest.mark.parametrize("geom_name,Q,kernel,wall,inlet,outlet,R", [
    ("cylinder_long", 19, "LBGK", "BFL", "NASH", "NASH", 1), ("tree", 19, "MRT", "GZS", "LADD", "NASH", 1),
    ("sac", 27, "TRT", "BFL", "NASH", "NASH", 1), ("tree", 15, "LBGK", "SBB", "LADD", "LADD", 1),
    ("cylinder_long", 19, "LBGK", "BFL", "NASH", "NASH", 3)])
def test_streaming_targets_as_runs(monkeypatch, geom_name, Q, kernel, wall, inlet, outlet, R):
    """The whole-part launches read the streaming targets of 32 consecutive sites as at most two runs per
    direction where the sites allow it (hlb_gpu_target_runs) and from the index planes elsewhere;
    HLB_NBR_RUNS=0 reads the planes everywhere.  Same bits both ways and as the oracle, on rank 1 of R
    too (its domain-edge part streams into the halo slots)."""
    geom = geometry(geom_name)
    rank = None if R == 1 else G.slab_decomposition(geom, R)
    r = R // 2
    dom = build_domains(geom, Q, rank, R)[r]
    inlets, outlets = iolets_for(geom, inlet, outlet)
    f0 = anisotropic_f(dom.N, Q, dom.totalSharedFs)
    runs = GpuLBM(dom, kernel, wall, inlet, outlet, tau=0.8, inlets=inlets, outlets=outlets)
    monkeypatch.setenv("HLB_NBR_RUNS", "0")
    planes = GpuLBM(dom, kernel, wall, inlet, outlet, tau=0.8, inlets=inlets, outlets=outlets)
    monkeypatch.delenv("HLB_NBR_RUNS")
    inRuns, words = runs.target_runs()
    assert words == (dom.N + 31) // 32 and 0 <= inRuns <= words
    assert inRuns > 0 or R > 1   # (a thin slab's rows are a few sites long: no group of 32 is two runs)
    if geom_name == "cylinder_long" and R == 1:
        assert inRuns > 0.5 * words
    assert planes.target_runs() == (0, 0)
    for g in (runs, planes):
        g.set_f(f0)
        g.set_cache_mask(255)
        if R == 1:
            g.step(5)
        else:  # no peers here: the received slots are given, the sends are compared
            for k in range(3):
                g.request_comms()
                g.pre_send()
                g.pre_receive()
                g.set_halo(0.05 + 0.001 * k + 1e-5 * np.arange(dom.totalSharedFs), which=0)
                g.post_receive()
                g.end_iteration()
                g.swap_old_and_new()
                g.state.increment()
    assert np.array_equal(runs.get_f(), planes.get_f())
    for which in (0, 1):
        assert np.array_equal(runs.get_halo(which=which), planes.get_halo(which=which))
    for name in O.CACHE_BITS:
        assert np.array_equal(runs.get_cache(name), planes.get_cache(name)), name
    if R == 1:
        sim = O.OracleSim(O.OracleDomains(geom, Q), kernel, wall, inlet, outlet, tau=0.8, inlets=inlets, outlets=outlets)
        sim.set_f(f0)
        sim.step(5)
        _check(runs.get_f()[:dom.N * Q], sim.get_f()[:dom.N * Q], "f_old, targets as runs")


@pytest.mark.parametrize("axis", (0, 2))
def test_runs_with_long_rows_and_a_halo(axis):
    """A cylinder whose lattice rows (along z) are 160 sites long, cut into two ranks across the rows (axis 2:
    every row loses an end to the domain-edge part) or along them (axis 0: whole rows are domain-edge and
    push into halo slots): most groups of 32 sites are in runs, the groups next to the cut are not, and
    both ranks match the oracle bit for bit through the host-staged halo."""
    geom, Q, R = G.cylinder(12.3, 160), 19, 2
    rank = G.slab_decomposition(geom, R, axis=axis)
    inlets, outlets = iolets_for(geom, "NASH", "NASH")
    doms = build_domains(geom, Q, rank, R)
    sim = O.OracleSim(O.OracleDomains(geom, Q, rank, R), "LBGK", "BFL", tau=0.8, inlets=inlets, outlets=outlets)
    gpus = [GpuLBM(d, "LBGK", "BFL", tau=0.8, inlets=inlets, outlets=outlets) for d in doms]
    for r, d in enumerate(doms):
        inRuns, words = gpus[r].target_runs()
        assert 0.5 * words < inRuns < words, (axis, r, inRuns, words)
        f0 = anisotropic_f(d.N, Q, d.totalSharedFs, site_offset=11 * r)
        sim.set_f(f0, r)
        gpus[r].set_f(f0)
    for _ in range(4):
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
    sim.step(4)
    for r, d in enumerate(doms):
        _check(gpus[r].get_f()[:d.N * Q], sim.get_f(r)[:d.N * Q], "axis %d rank %d f_old" % (axis, r))


def test_stability_reduction_matches_the_reference_loop():
    """hlb_gpu_stability = the site loop of lb::StabilityTester::PostSendToParent
    (Code/lb/StabilityTester.h:97-141) run where the reference runs it: after the step's streaming,
    before the swap."""
    geom, Q = geometry("cylinder"), 19
    dom = build_domains(geom, Q)[0]
    inlets, outlets = iolets_for(geom, "NASH", "NASH")
    gpu = GpuLBM(dom, "LBGK", "BFL", tau=0.8, inlets=inlets, outlets=outlets)
    f0 = perturbed_equilibrium(dom.N, Q, 0, O.lattice(Q)[1])
    gpu.set_f(f0)
    gpu.step(3)
    gpu.request_comms()
    gpu.pre_send()
    gpu.pre_receive()
    gpu.post_receive()
    bad, du = gpu.stability(True)
    f_old = gpu.get_f(0)[:dom.N * Q].reshape(dom.N, Q)
    f_new = gpu.get_f(1)[:dom.N * Q].reshape(dom.N, Q)
    c = O.lattice(Q)[0].astype(np.float64)
    u_old = (f_old @ c) / f_old.sum(1)[:, None]
    u_new = (f_new @ c) / f_new.sum(1)[:, None]
    want = np.sqrt(((u_new - u_old) ** 2).sum(1)).max()
    assert bad == 0 and want > 0
    assert abs(du - want) <= 1e-15 + 1e-12 * want
    assert gpu.stability(False) == (0, 0.0)
    # a negative and a NaN population in f_new are both "not > 0"
    f1 = gpu.get_f(1)
    f1[5], f1[Q * 7 + 3] = -1e-3, np.nan
    gpu.set_f(f1, which=1)
    assert gpu.stability(False)[0] == 2


@pytest.mark.parametrize("geom_name,Q,kernel,wall,inlet,outlet", [
    ("cylinder", 19, "MRT", "BFL", "NASH", "NASH"), ("tree", 19, "MRT", "BFL", "LADD", "NASH"),
    ("sac", 27, "TRT", "BFL", "NASH", "NASH"), ("cylinder", 27, "LBGK", "SBB", "NASH", "NASH"),
    ("cylinder", 19, "LBGK", "BFL", "NASH", "NASH"), ("tree", 19, "MRT", "GZS", "LADD", "NASH")])
def test_schedules_are_interchangeable(geom_name, Q, kernel, wall, inlet, outlet):
    """The product schedule (fused mid-domain kernel for MRT / D3Q27, second stream otherwise), the
    serial one (hlb_gpu_set_overlap(0)) and the phase API with its deferred whole-range launches --
    interleaved with sub-range calls, cache extraction and monitor read-backs -- give identical
    distributions, caches and monitors, and all equal the oracle."""
    geom = geometry(geom_name)
    inlets, outlets = iolets_for(geom, inlet, outlet)
    dom = build_domains(geom, Q)[0]
    odom = O.OracleDomains(geom, Q)
    sim = O.OracleSim(odom, kernel, wall, inlet, outlet, tau=0.8, inlets=inlets, outlets=outlets)
    f0 = perturbed_equilibrium(dom.N, Q, 0, O.lattice(Q)[1])
    engines = [GpuLBM(dom, kernel, wall, inlet, outlet, tau=0.8, inlets=inlets, outlets=outlets) for _ in range(3)]
    engines[1].set_overlap(False)
    for e in engines + [sim]:
        e.set_f(f0)
        e.set_cache_mask(255 if e is sim else 255 | 256)
    sim.step(6)
    engines[0].step(6)
    engines[1].step(6)
    c = engines[2]
    mid_total = int(dom.mid.sum())
    for it in range(6):
        c._push_scalars()
        c.request_comms()
        c.pre_send()
        off = 0
        for t in range(6):  # PreReceive, with the wall range split in two on odd steps
            n = int(dom.mid[t])
            if t == 1 and it % 2 and n > 3:
                c.stream_and_collide(t, off, n // 3)
                c.stream_and_collide(t, off + n // 3, n - n // 3)
            else:
                c.stream_and_collide(t, off, n)
            off += n
        if it == 2:
            c.monitor()  # a read-back in the middle of a step must flush what was deferred
        c.post_receive()
        c.swap_old_and_new()
        c.state.increment()
    want = sim.get_f()[:dom.N * Q]
    mons = []
    for e in engines:
        _check(e.get_f()[:dom.N * Q], want, "f")
        for name in O.CACHE_BITS:
            assert np.array_equal(e.get_cache(name), sim.get_cache(name)), name
        mons.append(e.monitor())
    assert mons[0] == mons[1]
