"""The C++ host-side drop-in RUNS: tests/host_lbm_run.cc drives hemelb_b200/host's Gpu*Streamer
policy classes and device-backed geometry::FieldData the way lb::LBM<Traits> drives its streamers
(Code/lb/lb.hpp:75-114, 162-314), with the reference's own LbmParameters / SimulationState /
InOutLet / SiteData / MacroscopicPropertyCache classes.

* CPU: linked against a recording stand-in for the C ABI (tests/host_mock_abi.cc); the recorded
  call sequence must be the one LBM's phases imply, with the policy of all six streamers known when
  the engine is created and the step scalars pushed once per step.
* GPU: linked against libhemelb_b200.so; distributions, density and velocity after K steps against
  the oracle.

The binaries need the reference's headers to build, so __graft_entry__.build() builds them into
tests/_build/ where /root/reference exists and they travel to the GPU box prebuilt."""
import os
import subprocess

import numpy as np
import pytest

from hemelb_b200 import geometry as G
from hemelb_b200.domain import build_domains
from tests.cases import anisotropic_f, geometry, iolets_for
from tests.host_build import BUILD, build_host_binaries

RHO, ETA, DX = 1000.0, 0.004, 1.0
KERNELS = {"LBGK": 0, "MRT": 1}
WALLS = {"SBB": 0, "BFL": 1, "GZS": 2}
IOLETS = {"NASH": 0, "LADD": 1}


def physical_dt(tau):
    return (tau - 0.5) / 3.0 * RHO / ETA


def reference_tau(dt):
    """LbmParameters.h:35."""
    cs2 = 1.0 / 3.0
    return 0.5 + (dt * ETA / RHO) / (cs2 * DX * DX)


def write_case(path, dom, kernel, wall, inlet, outlet, inlets, outlets, f0, steps, want, dt, rank=0, nranks=1,
               where_is=None):
    """``where_is``: (extent[3], rows of {x, y, z, rank, local id}) of every fluid site of the geometry --
    what geometry::Domain answers GetProcIdFromGlobalCoords / GetLocalContiguousId... from (GZS across ranks)."""
    t = dom.tables()
    Q, N = int(t["Q"]), int(t["N"])
    head = np.zeros(32, np.int64)
    head[0] = 0x484C4231
    head[1:7] = [Q, KERNELS[kernel], WALLS[wall], IOLETS[inlet], IOLETS[outlet], N]
    head[7:13] = t["mid"]
    head[13:19] = t["edge"]
    head[19] = t["totalSharedFs"]
    head[20], head[21], head[22], head[23] = len(inlets), len(outlets), steps, want
    procs = np.asarray(t["procs"], np.int64).reshape(-1, 3)
    head[24], head[25], head[26] = rank, nranks, procs.shape[0]
    head[27] = 0 if where_is is None else where_is[1].shape[0]
    with open(path, "wb") as fh:
        fh.write(head.tobytes())
        fh.write(np.array([dt, DX, RHO, ETA], np.float64).tobytes())
        for a, dt_ in ((t["neighbourIndices"], np.int64), (t["wallMask"], np.uint32), (t["ioletMask"], np.uint32),
                       (t["ioletId"], np.int32), (t["siteType"], np.int32), (t["distanceToWall"], np.float64),
                       (t["wallNormal"], np.float64), (t["globalCoords"], np.int64)):
            fh.write(np.ascontiguousarray(np.asarray(a).reshape(-1), dt_).tobytes())
        for recs in (inlets, outlets):
            for r in recs:
                fh.write(np.asarray(r, np.float64).tobytes())
        fh.write(np.asarray(f0, np.float64).tobytes())
        fh.write(procs.tobytes())
        fh.write(np.ascontiguousarray(t["streamingIndices"], np.int64).tobytes())
        if where_is is not None:
            fh.write(np.ascontiguousarray(where_is[0], np.int64).tobytes())
            fh.write(np.ascontiguousarray(where_is[1], np.int64).tobytes())
    return Q, N


def where_is_table(builder):
    """The site -> (rank, local contiguous id) table of a ``DomainBuilder``, for ``write_case``."""
    g = builder.geom
    ext = g.block_dims.astype(np.int64) * g.block_size
    rows = np.concatenate([g.coords.astype(np.int64), builder.rank_of_site.astype(np.int64)[:, None],
                           builder.local_of_input.astype(np.int64)[:, None]], 1)
    return ext, rows


def gzs_two_rank_cases(tmp_path, steps):
    """Two ranks of a slab-cut cylinder with GuoZhengShi walls: case files, domains, builder."""
    from hemelb_b200.domain import DomainBuilder
    geom, Q, R = geometry("cylinder_long"), 19, 2
    ros = G.slab_decomposition(geom, R)
    builder = DomainBuilder(geom, Q, ros, R)
    doms = builder.domains
    inlets, outlets = iolets_for(geom, "LADD", "NASH")
    dt = physical_dt(0.8)
    f0s = []
    for r, dom in enumerate(doms):
        f0s.append(anisotropic_f(dom.N, Q, dom.totalSharedFs, site_offset=5 * r))
        write_case(tmp_path / ("case%d.bin" % r), dom, "LBGK", "GZS", "LADD", "NASH", inlets, outlets, f0s[r], steps, 0, dt,
                   rank=r, nranks=R, where_is=where_is_table(builder))
    return geom, Q, R, ros, builder, doms, inlets, outlets, f0s, dt



def expected_calls(dom, steps, n_in, n_out, want):
    """What lb::LBM's phases ask of the C ABI, per time step (lb.hpp:162-309 + FieldData.cc:27-48 +
    SimulationMaster.impl.h:218-219), given the six site counts of the two halves of the domain."""
    mid, edge = [int(x) for x in dom.mid], [int(x) for x in dom.edge]
    mid_total = sum(mid)
    calls = []
    for s in range(steps):
        mask = want if s == steps - 1 else 0
        calls.append("request_comms")
        off = mid_total
        for t in range(6):
            if t == 0:
                calls.append(("set_step_scalars", s + 1, mask))
            calls.append("stream_and_collide %d %d %d" % (t, off, edge[t]))
            off += edge[t]
        calls.append("edge_done")
        off = 0
        for t in range(6):
            calls.append("stream_and_collide %d %d %d" % (t, off, mid[t]))
            off += mid[t]
        calls.append("copy_received")
        for first, counts in ((mid_total, edge), (0, mid)):
            off = first
            for t in range(6):
                calls.append("post_step %d %d %d" % (t, off, counts[t]))
                off += counts[t]
        if mask & 1:
            calls.append("get_cache 1")
        if mask & 2:
            calls.append("get_cache 2")
        calls.append("swap")
    return calls


needs_reference_or_prebuilt = pytest.mark.skipif(
    not os.path.isdir("/root/reference/Code") and not os.path.exists(os.path.join(BUILD, "host_lbm_run_mock")),
    reason="reference checkout absent and no prebuilt tests/_build")


@needs_reference_or_prebuilt
@pytest.mark.parametrize("name,Q,kernel,wall,inlet,outlet", [
    ("four_cube", 15, "LBGK", "SBB", "NASH", "NASH"),
    ("cylinder", 19, "LBGK", "BFL", "NASH", "NASH"),
    ("tree", 19, "MRT", "BFL", "LADD", "NASH"),
])
def test_cxx_host_call_sequence(tmp_path, name, Q, kernel, wall, inlet, outlet):
    build_host_binaries()
    geom = geometry(name)
    dom = build_domains(geom, Q)[0]
    inlets, outlets = iolets_for(geom, inlet, outlet)
    if inlet == "LADD":
        inlets[0][13] = 7  # InOutLetParabolicVelocity::SetWarmup
    steps, want, dt = 3, 3, physical_dt(0.8)
    f0 = anisotropic_f(dom.N, Q, 0)
    write_case(tmp_path / "case.bin", dom, kernel, wall, inlet, outlet, inlets, outlets, f0, steps, want, dt)
    env = dict(os.environ, HLB_MOCK_LOG=str(tmp_path / "calls.log"))
    r = subprocess.run([os.path.join(BUILD, "host_lbm_run_mock"), str(tmp_path / "case.bin"), str(tmp_path / "out.bin")],
                       env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    log = open(tmp_path / "calls.log").read().splitlines()

    # ---- engine construction: every policy known at hlb_gpu_create, tables handed over once
    # (the initial condition is written through the host view before the engine exists; the engine
    # is built by the first phase call, LBM::RequestComms -> FieldData::SendAndReceive)
    rest = log
    create = [ln for ln in log if ln.startswith("create ")]
    assert len(create) == 1
    kv = dict(x.split("=") for x in create[0].split()[1:])
    assert (int(kv["lattice"]), int(kv["kernel"]), int(kv["wall"]), int(kv["inlet"]), int(kv["outlet"])) == \
        (Q, KERNELS[kernel], WALLS[wall], IOLETS[inlet], IOLETS[outlet])
    assert float(kv["tau"]) == reference_tau(dt)
    assert int(kv["n_sites"]) == dom.N and int(kv["nranks"]) == 1 and int(kv["shared"]) == 0
    assert [int(x) for x in kv["mid"].split(",")] == [int(x) for x in dom.mid]
    assert [int(x) for x in kv["edge"].split(",")] == [int(x) for x in dom.edge]
    assert (int(kv["inlets"]), int(kv["outlets"])) == (len(inlets), len(outlets))
    assert rest[0] == create[0]
    build = rest[:rest.index("finalise") + 1]
    assert build[1] == "set_neighbour_indices 0 %d first=%d" % (dom.N, int(dom.neighbour_indices(0, 1)[0]))
    n_bulk = int(dom.mid[0])
    mid_total = int(sum(dom.mid))
    if mid_total > n_bulk:  # cut-link tables only for the boundary-typed range
        assert any(ln.startswith("set_site_data %d %d " % (n_bulk, mid_total - n_bulk)) for ln in build)
        assert "set_wall_distances %d %d" % (n_bulk, mid_total - n_bulk) in build
    assert "set_site_coords 0 %d" % dom.N in build
    min_density = min(float(r_[14]) for r_ in list(inlets) + list(outlets))
    # (a velocity iolet's warm-up length has no getter: the host classes read it off the ramp)
    assert "set_iolets 0 %d kind0=%d min_density=%.17g warmup0=%g" % (
        len(inlets), IOLETS[inlet], min_density, inlets[0][13] if inlet == "LADD" else 0) in build
    assert "set_iolets 1 %d kind0=%d min_density=%.17g warmup0=0" % (len(outlets), IOLETS[outlet], min_density) in build
    # the host mirror (initial condition) goes up once, right after the engine exists
    after = rest[rest.index("finalise") + 1:]
    assert after[0].startswith("set_f 0 f0=%.17g" % f0[0]) and after[1].startswith("set_f 1 ")

    # ---- the per-step sequence
    got = [ln for ln in rest if ln not in build and not ln.startswith("set_f ")]
    want_calls = expected_calls(dom, steps, len(inlets), len(outlets), want)
    # the final read-back of the distributions (f_old only: the mirrors refresh one array at a time)
    # and the destructor
    assert got[-2:] == ["get_f 0", "destroy"]
    got = got[:-2]
    assert len(got) == len(want_calls), (len(got), len(want_calls))
    for g_, w_ in zip(got, want_calls):
        if isinstance(w_, tuple):
            parts = g_.split()
            assert parts[0] == "set_step_scalars" and parts[1] == "t=%d" % w_[1] and parts[2] == "mask=%d" % w_[2], g_
            assert len(parts) == 3 + len(inlets) + len(outlets)
        else:
            assert g_ == w_
    # cosine iolet densities come from the reference's InOutLetCosine at the 0-indexed time step
    scal = [ln for ln in got if ln.startswith("set_step_scalars")]
    assert len(scal) == steps  # once per step, not once per streamer call
    if outlet == "NASH":
        import oracle as O
        r0 = outlets[0]
        for s, ln in enumerate(scal):
            v = float(dict(x.split("=") for x in ln.split()[1:])["out0"])
            assert v == O.ref_lib().href_cosine_density(*[O.C.c_double(float(x)) for x in r0[9:13]], O.C.c_uint64(s))


@needs_reference_or_prebuilt
@pytest.mark.parametrize("mode", [1, 2])
def test_cxx_stability_tester_moves_sixteen_bytes_a_step(tmp_path, mode):
    """hemelb_b200/host/lb/StabilityTester.h in place of the reference's: with a stability check every
    time step (mode 2: plus the velocity convergence check) the host never asks for the distribution
    arrays -- no hlb_gpu_get_f between the steps, one hlb_gpu_stability per step, after the step's
    last PostStep and before the swap -- and the verdict reaches SimulationState."""
    build_host_binaries()
    geom, Q = geometry("cylinder"), 19
    dom = build_domains(geom, Q)[0]
    inlets, outlets = iolets_for(geom, "NASH", "NASH")
    steps, dt = 4, physical_dt(0.8)
    f0 = anisotropic_f(dom.N, Q, 0)
    write_case(tmp_path / "case.bin", dom, "LBGK", "BFL", "NASH", "NASH", inlets, outlets, f0, steps, 0, dt)
    env = dict(os.environ, HLB_MOCK_LOG=str(tmp_path / "calls.log"), HLB_HOST_STABILITY=str(mode))
    r = subprocess.run([os.path.join(BUILD, "host_lbm_run_mock"), str(tmp_path / "case.bin"), str(tmp_path / "out.bin")],
                       env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    log = open(tmp_path / "calls.log").read().splitlines()
    per_step = log[log.index("finalise") + 1:]
    assert per_step.count("stability convergence=%d" % (1 if mode == 2 else 0)) == steps
    assert [ln for ln in per_step if ln.startswith("get_f")] == ["get_f 0"]  # the harness's own final read-back
    assert per_step.index("get_f 0") > len(per_step) - 3
    k = 0
    for _ in range(steps):
        k = per_step.index("stability convergence=%d" % (1 if mode == 2 else 0), k)
        assert per_step[k - 1].startswith("post_step 5 ") and per_step[k + 1] == "swap"
        k += 1
    # lb::Stability: Stable = 1; StableAndConverged = 2 needs |du| / reference <= tolerance: the stand-in
    # reports 1e-3 / 0.01 against 1e-9, so "stable, not converged"
    verdicts = [ln for ln in r.stderr.splitlines() if "stability" in ln]
    assert len(verdicts) == steps and all(v.endswith("stability 1") for v in verdicts)


@needs_reference_or_prebuilt
def test_cxx_incompressibility_checker_moves_thirty_two_bytes_a_step(tmp_path):
    """hemelb_b200/host/lb/IncompressibilityChecker.h in place of the reference's (constructed with the Domain
    and the property cache, as configuration/SimBuilder.h:216-226 does): one hlb_gpu_monitor per step after
    the step's last PostStep and before the swap, the gathering switched on in the step scalars
    (HLB_CACHE_MONITOR), no cache or distribution read-back; the tracker only ever widens."""
    build_host_binaries()
    geom, Q = geometry("cylinder"), 19
    dom = build_domains(geom, Q)[0]
    inlets, outlets = iolets_for(geom, "NASH", "NASH")
    steps, dt = 4, physical_dt(0.8)
    write_case(tmp_path / "case.bin", dom, "LBGK", "BFL", "NASH", "NASH", inlets, outlets, anisotropic_f(dom.N, Q, 0), steps, 0, dt)
    env = dict(os.environ, HLB_MOCK_LOG=str(tmp_path / "calls.log"), HLB_HOST_INCOMPRESSIBILITY="1")
    r = subprocess.run([os.path.join(BUILD, "host_lbm_run_mock"), str(tmp_path / "case.bin"), str(tmp_path / "out.bin")],
                       env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    log = open(tmp_path / "calls.log").read().splitlines()
    per_step = log[log.index("finalise") + 1:]
    assert per_step.count("monitor") == steps
    assert not [ln for ln in per_step if ln.startswith("get_cache")]
    assert [ln for ln in per_step if ln.startswith("get_f")] == ["get_f 0"]  # the harness's own final read-back
    scal = [ln for ln in per_step if ln.startswith("set_step_scalars")]
    assert len(scal) == steps and all(" mask=256" in ln for ln in scal)
    k = 0
    for _ in range(steps):
        k = per_step.index("monitor", k)
        assert per_step[k - 1].startswith("post_step 5 ") and per_step[k + 1] == "swap"
        k += 1
    lines = [ln.split() for ln in r.stderr.splitlines() if "densities" in ln]
    assert len(lines) == steps
    for n, w in enumerate(lines, 1):  # the stand-in ABI reports [1 - 0.001 n, 1 + 0.002 n], speed 0.003 n on its n-th call
        assert float(w[4]) == 1.0 - 0.001 * n and float(w[5]) == 1.0 + 0.002 * n and float(w[7]) == 0.003 * n
        assert int(w[9]) == 1  # within the 5 % allowed


@pytest.mark.gpu
@pytest.mark.parametrize("name,Q,kernel,wall,inlet,outlet", [
    ("four_cube", 15, "LBGK", "SBB", "NASH", "NASH"),
    ("cylinder", 19, "LBGK", "BFL", "NASH", "NASH"),
    ("tree", 19, "MRT", "BFL", "LADD", "NASH"),
])
def test_cxx_host_runs_on_the_gpu_and_matches_the_oracle(tmp_path, name, Q, kernel, wall, inlet, outlet):
    exe = os.path.join(BUILD, "host_lbm_run")
    if not os.path.exists(exe):
        if not os.path.isdir("/root/reference/Code"):
            pytest.skip("tests/_build/host_lbm_run was not prebuilt (needs the reference headers)")
        build_host_binaries()
    import oracle as O
    geom = geometry(name)
    dom = build_domains(geom, Q)[0]
    inlets, outlets = iolets_for(geom, inlet, outlet)
    if inlet == "LADD":
        inlets[0][13] = 4  # warm-up ramp over the first steps (InOutLetParabolicVelocity.cc:33-39)
    steps, want, dt = 5, 3, physical_dt(0.8)
    tau = reference_tau(dt)
    f0 = anisotropic_f(dom.N, Q, 0)
    write_case(tmp_path / "case.bin", dom, kernel, wall, inlet, outlet, inlets, outlets, f0, steps, want, dt)
    # libhemelb_b200.so needs libcudart.so.12: the CUDA toolkit's, or the one torch ships
    import sysconfig
    extra = ["/usr/local/cuda/lib64", os.path.join(sysconfig.get_paths()["purelib"], "nvidia", "cuda_runtime", "lib")]
    env = dict(os.environ, LD_LIBRARY_PATH=":".join([os.environ.get("LD_LIBRARY_PATH", "")] + extra).strip(":"),
               HLB_HOST_STABILITY="2",  # lb::StabilityTester (device-side stand-in) assessing every step
               HLB_HOST_INCOMPRESSIBILITY="1")  # and an lb::IncompressibilityChecker (likewise)
    r = subprocess.run([exe, str(tmp_path / "case.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True,
                       timeout=300, env=env)
    assert r.returncode == 0, r.stderr
    verdicts = [int(ln.split()[-1]) for ln in r.stderr.splitlines() if "stability" in ln]
    # the checker's tracker after every step: extrema, over all steps so far, of the density and speed of
    # the distributions that entered the step (what UpdateCachePostCollision would have cached)
    tracked = [[float(w[4]), float(w[5]), float(w[7])] for w in (ln.split() for ln in r.stderr.splitlines() if "densities" in ln)]
    probe = O.OracleSim(O.OracleDomains(geom, Q), kernel, wall, inlet, outlet, tau=reference_tau(dt), inlets=inlets, outlets=outlets)
    probe.set_f(f0)
    cx = np.asarray(O.lattice(Q)[0], np.float64).reshape(Q, 3) if np.asarray(O.lattice(Q)[0]).size == 3 * Q else None
    # (as in the reference, the slots of children that do not exist hold REFERENCE_DENSITY = 1 and take part in
    # the merge, IncompressibilityChecker.hpp:112-139: the tracked range always contains 1)
    lo, hi, sp = 1.0, 1.0, 0.0
    assert len(tracked) == steps
    for s_ in range(steps):
        fs = probe.get_f()[:dom.N * Q].reshape(dom.N, Q)
        rho_ = fs.sum(1)
        lo, hi = min(lo, rho_.min()), max(hi, rho_.max())
        assert abs(tracked[s_][0] - lo) <= 1e-12 * abs(lo) and abs(tracked[s_][1] - hi) <= 1e-12 * abs(hi), (s_, tracked[s_], lo, hi)
        if cx is not None:
            sp = max(sp, (np.linalg.norm(fs @ cx, axis=1) / rho_).max())
            assert abs(tracked[s_][2] - sp) <= 1e-10 * max(sp, 1e-300), (s_, tracked[s_], sp)
        probe.step(1)
    out = np.fromfile(tmp_path / "out.bin", np.float64)
    # lb::Unstable = 0, Stable = 1, StableAndConverged = 2.  The anisotropic start is far from equilibrium
    # (populations do go negative in the first steps) and, as in the reference, Unstable sticks until Reset()
    assert len(verdicts) == 5 and set(verdicts) <= {0, 1, 2}, verdicts
    assert all(b == 0 for a, b in zip(verdicts, verdicts[1:]) if a == 0), verdicts
    N = dom.N
    assert out.size == N * Q + N + 3 * N
    sim = O.OracleSim(O.OracleDomains(geom, Q), kernel, wall, inlet, outlet, tau=tau, inlets=inlets, outlets=outlets)
    sim.set_f(f0)
    sim.step(steps - 1)
    sim.set_cache_mask(3)
    sim.step(1)
    assert np.abs(out[:N * Q] - sim.get_f()[:N * Q]).max() <= 1e-13
    rho, vel = sim.get_cache("density"), sim.get_cache("velocity")
    assert np.abs(out[N * Q:N * Q + N] - rho).max() <= 1e-10 * np.abs(rho).max()
    assert np.abs(out[N * Q + N:] - vel.reshape(-1)).max() <= 1e-10 * max(np.abs(vel).max(), 1e-300) + 1e-15


@needs_reference_or_prebuilt
def test_cxx_host_multi_rank_construction_and_sequence(tmp_path):
    """Two ranks of a slab-cut cylinder: what FieldData::EnsureEngine tells the engine about the halo
    (neighbouringProcs, streamingIndicesForReceivedDistributions, the NCCL id broadcast over the
    communicator) and the per-step calls with non-empty domain-edge ranges."""
    build_host_binaries()
    geom, Q, R = geometry("cylinder"), 19, 2
    doms = build_domains(geom, Q, G.slab_decomposition(geom, R, axis=2), R)
    inlets, outlets = iolets_for(geom, "NASH", "NASH")
    steps, dt = 2, physical_dt(0.8)
    for r, dom in enumerate(doms):
        assert dom.totalSharedFs > 0 and sum(dom.edge) > 0
        f0 = np.zeros(dom.N * Q + 1 + dom.totalSharedFs)
        f0[:dom.N * Q] = anisotropic_f(dom.N, Q, 0)[:dom.N * Q]
        case = tmp_path / ("case%d.bin" % r)
        write_case(case, dom, "LBGK", "BFL", "NASH", "NASH", inlets, outlets, f0, steps, 0, dt, rank=r, nranks=R)
        env = dict(os.environ, HLB_MOCK_LOG=str(tmp_path / ("calls%d.log" % r)), HLB_HOST_ID_FILE=str(tmp_path / "nccl_id"))
        p = subprocess.run([os.path.join(BUILD, "host_lbm_run_mock"), str(case), str(tmp_path / "out.bin")], env=env,
                           capture_output=True, text=True, timeout=120)
        assert p.returncode == 0, p.stderr
        log = open(tmp_path / ("calls%d.log" % r)).read().splitlines()
        kv = dict(x.split("=") for x in log[0].split()[1:])
        assert (int(kv["rank"]), int(kv["nranks"]), int(kv["shared"]), int(kv["neighbours"])) == \
            (r, R, int(dom.totalSharedFs), dom.procs.shape[0])
        assert [int(x) for x in kv["edge"].split(",")] == [int(x) for x in dom.edge]
        build = log[:log.index("finalise") + 1]
        assert "set_neighbours " + " ".join("%d:%d:%d" % tuple(int(v) for v in row) for row in dom.procs) in build
        ws = int(((np.arange(dom.totalSharedFs, dtype=np.int64) + 1) * np.asarray(dom.streamingIndices, np.int64)).sum())
        assert "set_streaming_indices weighted_sum=%d" % ws in build
        # boundary-typed tables for both halves of the site order
        mid_total, n_bulk, e_bulk = int(sum(dom.mid)), int(dom.mid[0]), int(dom.edge[0])
        assert "set_wall_distances %d %d" % (n_bulk, mid_total - n_bulk) in build
        assert "set_wall_distances %d %d" % (mid_total + e_bulk, dom.N - mid_total - e_bulk) in build
        # the NCCL communicator comes up right after the tables, before the distributions go up
        after = log[log.index("finalise") + 1:]
        assert after[0] == "comm_init" and after[1].startswith("set_f 0 ")
        got = [ln for ln in after[3:] if not ln.startswith("set_f ")]
        want_calls = expected_calls(dom, steps, 1, 1, 0)
        assert got[-2:] == ["get_f 0", "destroy"]
        got = got[:-2]
        assert len(got) == len(want_calls)
        for g_, w_ in zip(got, want_calls):
            assert g_ == w_ if not isinstance(w_, tuple) else g_.startswith("set_step_scalars t=%d mask=0" % w_[1])
    assert os.path.getsize(tmp_path / "nccl_id") == 128


@needs_reference_or_prebuilt
def test_cxx_host_gzs_across_ranks_lists_and_sequence(tmp_path):
    """GuoZhengShi walls on two ranks through the C++ classes: the Gpu streamers' constructors register
    the remote needs with the NeighbouringDataManager (GuoZhengShi.h:36-104), ShareNeeds tells each rank
    what to serve, and the engine is given exactly the link and serve lists of the Python mirror; the
    site halo is exchanged once per time step, before the step's first range."""
    build_host_binaries()
    steps = 3
    geom, Q, R, ros, builder, doms, inlets, outlets, f0s, dt = gzs_two_rank_cases(tmp_path, steps)
    need, serve = builder.gzs_site_halo()
    procs = []
    for r in range(R):
        env = dict(os.environ, HLB_MOCK_LOG=str(tmp_path / ("calls%d.log" % r)), HLB_HOST_ID_FILE=str(tmp_path / "id"))
        procs.append(subprocess.Popen([os.path.join(BUILD, "host_lbm_run_mock"), str(tmp_path / ("case%d.bin" % r)),
                                       str(tmp_path / ("out%d.bin" % r))], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    for p in procs:
        _, err = p.communicate(timeout=120)
        assert p.returncode == 0, err
    ext = geom.block_dims.astype(np.int64) * geom.block_size
    total_links = 0
    for r, dom in enumerate(doms):
        log = open(tmp_path / ("calls%d.log" % r)).read().splitlines()
        build = log[:log.index("finalise") + 1]
        # one row per link: site, direction, owner rank -- as the Python mirror lists them
        got = [ln for ln in build if ln.startswith("set_gzs_remote ")]
        nd = need[r]
        total_links += nd.shape[0]
        if nd.shape[0]:
            rows = [tuple(int(v) for v in x.split(":")) for x in got[0].split()[2:]]
            assert len(rows) == nd.shape[0]
            assert sorted(x[:3] for x in rows) == sorted((int(a), int(b), int(c)) for a, b, c, _ in nd)
            # grouped by owner rank; the key is the neighbour's global non-contiguous id
            assert [x[2] for x in rows] == sorted(x[2] for x in rows)
            inp_of = {(int(builder.rank_of_site[i]), int(builder.local_of_input[i])): i for i in range(geom.n_sites)}
            for site, direction, owner, key in rows:
                c = geom.coords[inp_of[(r, site)]].astype(np.int64) + builder.c[direction]
                assert key == (c[0] * ext[1] + c[1]) * ext[2] + c[2]
        else:
            assert not got
        # what is served: each (requester, site) once, in the requester's order
        got = [ln for ln in build if ln.startswith("set_gzs_serve ")]
        sv = serve[r]
        if sv.shape[0]:
            assert sorted(tuple(int(v) for v in x.split(":")) for x in got[0].split()[2:]) == \
                sorted((int(a), int(b)) for a, b in sv)
        # per step: the site halo first
        after = [ln for ln in log[log.index("finalise") + 1:] if not ln.startswith("set_f ") and ln != "comm_init"]
        assert after.count("exchange_site_halo") == steps
        firsts = [i for i, ln in enumerate(after) if ln == "exchange_site_halo"]
        for i in firsts:
            nxt = after[i + 1]
            assert nxt.startswith("set_step_scalars") or nxt.startswith("stream_and_collide") or nxt == "request_comms"
    assert total_links > 0


@pytest.mark.gpu
def test_cxx_host_gzs_two_ranks_over_nccl(tmp_path):
    """The same two ranks on two GPUs: site halo and distribution halo over NCCL, each rank's
    distributions against the oracle's emulated 2-rank run.  Needs 2 GPUs; skipped otherwise."""
    from tests.test_gpu_multi import _gpu_count
    if _gpu_count() < 2:
        pytest.skip("needs 2 GPUs")
    exe = os.path.join(BUILD, "host_lbm_run")
    if not os.path.exists(exe):
        if not os.path.isdir("/root/reference/Code"):
            pytest.skip("tests/_build/host_lbm_run was not prebuilt (needs the reference headers)")
        build_host_binaries()
    import sysconfig
    import oracle as O
    steps = 6
    geom, Q, R, ros, builder, doms, inlets, outlets, f0s, dt = gzs_two_rank_cases(tmp_path, steps)
    tau = reference_tau(dt)
    extra = ["/usr/local/cuda/lib64", os.path.join(sysconfig.get_paths()["purelib"], "nvidia", "cuda_runtime", "lib"),
             os.path.join(sysconfig.get_paths()["purelib"], "nvidia", "nccl", "lib")]
    env = dict(os.environ, LD_LIBRARY_PATH=":".join([os.environ.get("LD_LIBRARY_PATH", "")] + extra).strip(":"),
               HLB_HOST_ID_FILE=str(tmp_path / "nccl_id"))
    procs = [subprocess.Popen([exe, str(tmp_path / ("case%d.bin" % r)), str(tmp_path / ("out%d.bin" % r))], env=env,
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(R)]
    for p in procs:
        _, err = p.communicate(timeout=300)
        assert p.returncode == 0, err
    sim = O.OracleSim(O.OracleDomains(geom, Q, ros, R), "LBGK", "GZS", "LADD", "NASH", tau=tau, inlets=inlets, outlets=outlets)
    for r in range(R):
        sim.set_f(f0s[r], r)
    sim.step(steps)
    for r, dom in enumerate(doms):
        out = np.fromfile(tmp_path / ("out%d.bin" % r), np.float64)
        assert np.abs(out[:dom.N * Q] - sim.get_f(r)[:dom.N * Q]).max() <= 1e-13, r


@pytest.mark.gpu
def test_cxx_host_two_ranks_over_nccl(tmp_path):
    """Two harness processes, one per GPU (rank r -> device r), the NCCL id broadcast through the
    stand-in communicator; each rank's distributions against the oracle's emulated 2-rank run.
    Needs 2 GPUs (gpurun --gpus 2); skipped otherwise."""
    from tests.test_gpu_multi import _gpu_count
    if _gpu_count() < 2:
        pytest.skip("needs 2 GPUs")
    exe = os.path.join(BUILD, "host_lbm_run")
    if not os.path.exists(exe):
        if not os.path.isdir("/root/reference/Code"):
            pytest.skip("tests/_build/host_lbm_run was not prebuilt (needs the reference headers)")
        build_host_binaries()
    import sysconfig
    import oracle as O
    geom, Q, R = geometry("cylinder"), 19, 2
    ros = G.slab_decomposition(geom, R, axis=2)
    doms = build_domains(geom, Q, ros, R)
    inlets, outlets = iolets_for(geom, "NASH", "NASH")
    steps, dt = 6, physical_dt(0.8)
    tau = reference_tau(dt)
    extra = ["/usr/local/cuda/lib64", os.path.join(sysconfig.get_paths()["purelib"], "nvidia", "cuda_runtime", "lib"),
             os.path.join(sysconfig.get_paths()["purelib"], "nvidia", "nccl", "lib")]
    env = dict(os.environ, LD_LIBRARY_PATH=":".join([os.environ.get("LD_LIBRARY_PATH", "")] + extra).strip(":"),
               HLB_HOST_ID_FILE=str(tmp_path / "nccl_id"))
    f0s, procs = [], []
    for r, dom in enumerate(doms):
        f0s.append(anisotropic_f(dom.N, Q, dom.totalSharedFs, site_offset=7 * r))
        write_case(tmp_path / ("case%d.bin" % r), dom, "LBGK", "BFL", "NASH", "NASH", inlets, outlets, f0s[r], steps, 0, dt,
                   rank=r, nranks=R)
        procs.append(subprocess.Popen([exe, str(tmp_path / ("case%d.bin" % r)), str(tmp_path / ("out%d.bin" % r))], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for r, p in enumerate(procs):
        try:
            out, _ = p.communicate(timeout=300)
        except subprocess.TimeoutExpired:
            p.kill()
            out, _ = p.communicate()
        assert p.returncode == 0, "rank %d failed:\n%s" % (r, out[-3000:])
    sim = O.OracleSim(O.OracleDomains(geom, Q, ros, R), "LBGK", "BFL", "NASH", "NASH", tau=tau, inlets=inlets,
                      outlets=outlets)
    for r in range(R):
        sim.set_f(f0s[r], r)
    sim.step(steps)
    for r, dom in enumerate(doms):
        got = np.fromfile(tmp_path / ("out%d.bin" % r), np.float64)
        assert got.size == dom.N * Q
        assert np.abs(got - sim.get_f(r)[:dom.N * Q]).max() <= 1e-13


@pytest.mark.skipif(not os.path.isdir("/root/reference/Code") and not os.path.exists(os.path.join(BUILD, "host_xtr_run_mock")),
                    reason="reference checkout absent and no prebuilt tests/_build")
def test_cxx_property_encoder_translates_the_reference_spec(tmp_path):
    """extraction/GpuPropertyEncoder.h: the reference's PropertyOutputFile objects (one per selector
    class, every source type among the fields) as they reach hlb_xtr_create, and the sizes / header /
    record calls LocalPropertyOutput's members map to."""
    build_host_binaries()
    env = dict(os.environ, HLB_MOCK_LOG=str(tmp_path / "xtr.log"))
    r = subprocess.run([os.path.join(BUILD, "host_xtr_run_mock")], env=env, capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
    log = open(tmp_path / "xtr.log").read().splitlines()
    creates = [ln for ln in log if ln.startswith("xtr_create")]
    assert len(creates) == 6

    def parse(ln):
        parts = ln.split()
        kv = dict(p.split("=", 1) for p in parts[1:] if not p.startswith("field="))
        return int(kv["selector"]), [float(x) for x in kv["params"].split(",")], kv["units"], kv["coords0"], \
            [p[6:] for p in parts if p.startswith("field=")]

    f32 = lambda x: float(np.float32(x))
    # HLB_XTR_* selectors: whole 0, surface 1, line 3, surface point 4, plane with the constructor-normalised normal 5
    sel, par, units, c0, fields = parse(creates[0])
    assert (sel, c0) == (0, "3,4,5")
    assert units == "0.0001,0.00020000000000000001,0.10000000000000001,0.20000000000000001,0.29999999999999999,1000,80"
    # name : source (OutputField.h:33-44 order) : type code : offsets
    assert fields == ["pressure:0:0:1:80", "velocity:1:0:0", "distributions:8:1:0"]
    sel, par, _, _, fields = parse(creates[1])
    assert sel == 1 and fields == ["shearstress:2:0:0", "traction:6:1:0"]
    sel, par, _, _, fields = parse(creates[2])
    assert sel == 5 and fields == ["velocity:1:1:0", "stresstensor:5:0:0", "rank:9:2:0"]
    assert np.allclose(par, [f32(0.001), f32(0.002), f32(0.003), 0, 0, 1, f32(0.004)], rtol=1e-8, atol=0)
    sel, par, _, _, fields = parse(creates[3])  # infinite plane: radius 0, normal (3,0,4)/5 in float
    assert sel == 5 and np.allclose(par, [0.5, 0, 0, f32(0.6), 0, f32(0.8), 0], rtol=1e-7, atol=0)
    sel, par, _, _, fields = parse(creates[4])
    assert sel == 3 and np.allclose(par, [0, 0, 0, 0, 0, f32(0.01), 0], rtol=1e-8, atol=0)
    assert fields == ["vonmisesstress:3:0:0", "shearrate:4:0:0"]
    sel, par, _, _, fields = parse(creates[5])
    assert sel == 4 and np.allclose(par[:3], [f32(0.01), f32(0.02), f32(0.03)], rtol=1e-8, atol=0)
    assert fields == ["tangentialprojectiontraction:7:0:0"]
    # per encoder: header of HeaderLength bytes for the global count, records of count x site length
    assert log[1:4] == ["xtr_header global=1234 capacity=72", "xtr_encode first=0 n=3 capacity=72", "xtr_destroy"]
    out = r.stdout.splitlines()
    assert out[0] == "sites=3 site_len=24 header_len=72 header0=H records0=R caches=103"


@pytest.mark.skipif(not os.path.isdir("/root/reference/Code") and not os.path.exists(os.path.join(BUILD, "libhost_lbm_real.so")),
                    reason="reference checkout absent and no prebuilt tests/_build/libhost_lbm_real.so")
@pytest.mark.parametrize("name,R,kind", [("cylinder", 1, None), ("cylinder", 2, "slab"), ("tree", 3, "basic"), ("tree", 4, "ragged")])
def test_the_reference_lbm_itself_drives_the_gpu_classes(tmp_path, name, R, kind):
    """tests/host_lbm_real.cc: the reference's own lb::LBM<Traits> (lb.h / lb.hpp, unmodified), constructed and
    stepped as SimulationMaster steps it, over the reference's own geometry::Domain, net::Net, lb::BoundaryValues,
    SimulationState, LbmParameters and Timers (all compiled unmodified; R ranks = threads over oracle/fake_mpi.cc),
    with hemelb_b200/host's streamers in the Traits and its FieldData in the reference's place.  Behind the C ABI
    sits the recording stand-in: on every rank the engine must be created with that rank's tables -- the ones
    hemelb_b200.domain builds, bit for bit what the reference's Domain holds -- and every time step must be
    the call sequence of lb.hpp:162-309."""
    import ctypes as C
    from hemelb_b200 import geometry as G
    from tests.test_domain_vs_ref import decomposition
    build_host_binaries()
    L = C.CDLL(os.path.join(BUILD, "libhost_lbm_real.so"))
    geom, Q = geometry(name), 19
    inlets, outlets = iolets_for(geom, "NASH", "NASH")
    rank = decomposition(geom, R, kind)
    doms = build_domains(geom, Q, rank, R)
    steps, dt = 3, physical_dt(0.8)
    os.environ["HLB_MOCK_LOG"] = str(tmp_path / "calls.log")
    try:
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        arrs = [np.ascontiguousarray(geom.coords, np.int32), np.ascontiguousarray(geom.bsite, np.int64),
                np.ascontiguousarray(geom.btype, np.uint8), np.ascontiguousarray(geom.biolet, np.int32),
                np.ascontiguousarray(geom.bdist, np.float32), np.ascontiguousarray(geom.bnavail, np.uint8),
                np.ascontiguousarray(geom.bnormal, np.float32)]
        bd = np.ascontiguousarray(geom.block_dims, np.int32)
        rk = None if rank is None else np.ascontiguousarray(rank, np.int32)
        inr, outr = np.ascontiguousarray(np.stack(inlets)), np.ascontiguousarray(np.stack(outlets))
        rc = L.hreal_run(R, p(bd), int(geom.block_size), C.c_int64(geom.n_sites), p(arrs[0]), C.c_int64(arrs[1].size),
                         *[p(a) for a in arrs[1:]], None if rk is None else p(rk), C.c_double(dt), C.c_double(DX),
                         len(inlets), p(inr), len(outlets), p(outr), C.c_int64(steps), None, None, 1, 0, None, 0)
        assert rc == 0
    finally:
        del os.environ["HLB_MOCK_LOG"]
    for r, dom in enumerate(doms):
        log = open(str(tmp_path / "calls.log") + ".rank%d" % r).read().splitlines()
        kv = dict(x.split("=") for x in log[0].split()[1:])
        assert log[0].startswith("create ")
        assert (int(kv["lattice"]), int(kv["kernel"]), int(kv["wall"]), int(kv["inlet"]), int(kv["outlet"])) == (19, 0, 1, 0, 0)
        assert float(kv["tau"]) == reference_tau(dt)
        assert (int(kv["rank"]), int(kv["nranks"]), int(kv["n_sites"]), int(kv["shared"])) == (r, R, dom.N, dom.totalSharedFs)
        assert [int(x) for x in kv["mid"].split(",")] == [int(x) for x in dom.mid]
        assert [int(x) for x in kv["edge"].split(",")] == [int(x) for x in dom.edge]
        assert int(kv["neighbours"]) == dom.procs.shape[0]
        build = log[:log.index("finalise") + 1]
        assert build[1] == "set_neighbour_indices 0 %d first=%d" % (dom.N, int(dom.neighbour_indices(0, 1)[0]))
        if dom.procs.shape[0]:
            assert "set_neighbours " + " ".join("%d:%d:%d" % tuple(int(x) for x in row) for row in dom.procs) in build
            w = int((np.arange(1, dom.totalSharedFs + 1, dtype=np.int64) * dom.streamingIndices).sum())
            assert "set_streaming_indices weighted_sum=%d" % w in build
        assert "set_site_coords 0 %d" % dom.N in build
        # (the rank that owns the boundary-condition task keeps every iolet; the table always has them all)
        assert any(ln.startswith("set_iolets 0 %d kind0=0 " % len(inlets)) for ln in build)
        assert any(ln.startswith("set_iolets 1 %d kind0=0 " % len(outlets)) for ln in build)
        got = [ln for ln in log if ln not in build and not ln.startswith("set_f ")]
        assert got[-1] == "destroy"
        got = got[:-1]
        if R > 1:  # the NCCL id travelled over the reference's own communicator (MpiCommunicator::Broadcast)
            boot = [ln for ln in got if ln.startswith("comm_")]
            assert boot == ["comm_init"], boot
            got = [ln for ln in got if not ln.startswith("comm_")]
        want_calls = expected_calls(dom, steps, len(inlets), len(outlets), 0)
        assert len(got) == len(want_calls), (r, len(got), len(want_calls))
        for g_, w_ in zip(got, want_calls):
            if isinstance(w_, tuple):
                parts = g_.split()
                assert parts[0] == "set_step_scalars" and parts[1] == "t=%d" % w_[1] and parts[2] == "mask=%d" % w_[2], g_
                assert len(parts) == 3 + len(inlets) + len(outlets)
            else:
                assert g_ == w_, (r, g_, w_)


@pytest.mark.gpu
def test_the_reference_lbm_itself_on_the_gpu_matches_the_oracle():
    """The same harness linked against libhemelb_b200.so: the reference's lb::LBM over its own Domain, net::Net and
    BoundaryValues steps the engine on the GPU; the distributions after five steps equal the oracle's (1e-13)."""
    import ctypes as C
    lib = os.path.join(BUILD, "libhost_lbm_real_gpu.so")
    if not os.path.exists(lib):
        if not os.path.isdir("/root/reference/Code"):
            pytest.skip("tests/_build/libhost_lbm_real_gpu.so was not prebuilt (needs the reference sources)")
        build_host_binaries()
    import oracle as O
    from hemelb_b200 import capi
    capi.lib()  # libhemelb_b200.so first: the harness library names it as a dependency
    L = C.CDLL(lib)
    geom, Q = geometry("cylinder"), 19
    inlets, outlets = iolets_for(geom, "NASH", "NASH")
    dom = build_domains(geom, Q)[0]
    steps, dt = 5, physical_dt(0.8)
    f0 = anisotropic_f(dom.N, Q, 0)
    fin = np.ascontiguousarray(f0[:dom.N * Q])
    out = np.zeros(dom.N * Q)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    arrs = [np.ascontiguousarray(geom.coords, np.int32), np.ascontiguousarray(geom.bsite, np.int64),
            np.ascontiguousarray(geom.btype, np.uint8), np.ascontiguousarray(geom.biolet, np.int32),
            np.ascontiguousarray(geom.bdist, np.float32), np.ascontiguousarray(geom.bnavail, np.uint8),
            np.ascontiguousarray(geom.bnormal, np.float32)]
    bd = np.ascontiguousarray(geom.block_dims, np.int32)
    inr, outr = np.ascontiguousarray(np.stack(inlets)), np.ascontiguousarray(np.stack(outlets))
    rc = L.hreal_run(1, p(bd), int(geom.block_size), C.c_int64(geom.n_sites), p(arrs[0]), C.c_int64(arrs[1].size),
                     *[p(a) for a in arrs[1:]], None, C.c_double(dt), C.c_double(DX), len(inlets), p(inr), len(outlets),
                     p(outr), C.c_int64(steps), p(fin), p(out), 1, 0, None, 0)
    assert rc == 0
    sim = O.OracleSim(O.OracleDomains(geom, Q), "LBGK", "BFL", "NASH", "NASH", tau=reference_tau(dt), inlets=inlets, outlets=outlets)
    sim.set_f(f0)
    sim.step(steps)
    assert np.abs(out - sim.get_f()[:dom.N * Q]).max() <= 1e-13


@pytest.mark.skipif(not os.path.isdir("/root/reference/Code") and not os.path.exists(os.path.join(BUILD, "libhost_lbm_real.so")),
                    reason="reference checkout absent and no prebuilt tests/_build/libhost_lbm_real.so")
@pytest.mark.parametrize("R", (2, 3))
def test_the_reference_lbm_with_guo_zheng_shi_walls_across_ranks(tmp_path, R):
    """The same harness with GuoZhengShi walls: the streamers' constructors ask the reference's own Domain which rank
    owns the neighbour of every extrapolating wall link (Domain::GetProcIdFromGlobalCoords -> the
    DistributedStore's one-sided windows) and register the remote ones with hemelb_b200/host's
    NeighbouringDataManager, whose ShareNeeds runs over the reference's own net::Net (all-to-all + point-to-point).
    Every rank's engine must be given the link and serve lists of the Python mirror, and the site halo must be
    exchanged once per time step, before the step's first range."""
    import ctypes as C
    from hemelb_b200 import geometry as G
    from hemelb_b200.domain import DomainBuilder
    build_host_binaries()
    L = C.CDLL(os.path.join(BUILD, "libhost_lbm_real.so"))
    geom, Q = geometry("cylinder_long"), 19
    ros = G.slab_decomposition(geom, R)
    builder = DomainBuilder(geom, Q, ros, R)
    doms = builder.domains
    need, serve = builder.gzs_site_halo()
    inlets, outlets = iolets_for(geom, "NASH", "NASH")
    steps, dt = 3, physical_dt(0.8)
    os.environ["HLB_MOCK_LOG"] = str(tmp_path / "calls.log")
    try:
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        arrs = [np.ascontiguousarray(geom.coords, np.int32), np.ascontiguousarray(geom.bsite, np.int64),
                np.ascontiguousarray(geom.btype, np.uint8), np.ascontiguousarray(geom.biolet, np.int32),
                np.ascontiguousarray(geom.bdist, np.float32), np.ascontiguousarray(geom.bnavail, np.uint8),
                np.ascontiguousarray(geom.bnormal, np.float32)]
        bd = np.ascontiguousarray(geom.block_dims, np.int32)
        rk = np.ascontiguousarray(ros, np.int32)
        inr, outr = np.ascontiguousarray(np.stack(inlets)), np.ascontiguousarray(np.stack(outlets))
        rc = L.hreal_run(R, p(bd), int(geom.block_size), C.c_int64(geom.n_sites), p(arrs[0]), C.c_int64(arrs[1].size),
                         *[p(a) for a in arrs[1:]], p(rk), C.c_double(dt), C.c_double(DX), len(inlets), p(inr),
                         len(outlets), p(outr), C.c_int64(steps), None, None, 2, 0, None, 0)
        assert rc == 0
    finally:
        del os.environ["HLB_MOCK_LOG"]
    ext = geom.block_dims.astype(np.int64) * geom.block_size
    total_links = 0
    for r, dom in enumerate(doms):
        log = open(str(tmp_path / "calls.log") + ".rank%d" % r).read().splitlines()
        assert " wall=2 " in log[0]
        build = log[:log.index("finalise") + 1]
        got = [ln for ln in build if ln.startswith("set_gzs_remote ")]
        nd = need[r]
        total_links += nd.shape[0]
        if nd.shape[0]:
            rows = [tuple(int(v) for v in x.split(":")) for x in got[0].split()[2:]]
            assert sorted(x[:3] for x in rows) == sorted((int(a), int(b), int(c)) for a, b, c, _ in nd)
            assert [x[2] for x in rows] == sorted(x[2] for x in rows)  # grouped by owner rank
            inp_of = {(int(builder.rank_of_site[i]), int(builder.local_of_input[i])): i for i in range(geom.n_sites)}
            for site, direction, owner, key in rows:  # the key is the neighbour's global non-contiguous id
                c = geom.coords[inp_of[(r, site)]].astype(np.int64) + builder.c[direction]
                assert key == (c[0] * ext[1] + c[1]) * ext[2] + c[2]
        else:
            assert not got
        got = [ln for ln in build if ln.startswith("set_gzs_serve ")]
        sv = serve[r]
        if sv.shape[0]:
            assert sorted(tuple(int(v) for v in x.split(":")) for x in got[0].split()[2:]) == \
                sorted((int(a), int(b)) for a, b in sv)
        after = [ln for ln in log[log.index("finalise") + 1:] if not ln.startswith("set_f ") and ln != "comm_init"]
        assert after.count("exchange_site_halo") == steps
        for i in [k for k, ln in enumerate(after) if ln == "exchange_site_halo"]:
            nxt = after[i + 1]
            assert nxt.startswith("set_step_scalars") or nxt.startswith("stream_and_collide") or nxt == "request_comms"
    assert total_links > 0


@pytest.mark.skipif(not os.path.isdir("/root/reference/Code") and not os.path.exists(os.path.join(BUILD, "libhost_lbm_real.so")),
                    reason="reference checkout absent and no prebuilt tests/_build/libhost_lbm_real.so")
@pytest.mark.parametrize("R", (2, 3, 5))
def test_the_monitor_stand_ins_over_the_reference_broadcast_tree(tmp_path, R):
    """hemelb_b200/host's lb::StabilityTester and lb::IncompressibilityChecker as actions beside the reference's
    lb::LBM, their up-and-down passes carried by the reference's own net::PhasedBroadcastRegular / PhasedBroadcast
    over its net::Net (R ranks as threads).  The recording ABI gives every rank extrema of its own
    ([1 - 0.001 (r + 1), 1 + 0.002 (r + 1)], speed 0.003 (r + 1)): after a few cycles every rank must hold the
    extrema of all of them, and a stable verdict."""
    import ctypes as C
    from hemelb_b200 import geometry as G
    build_host_binaries()
    L = C.CDLL(os.path.join(BUILD, "libhost_lbm_real.so"))
    geom, Q = geometry("tree"), 19
    inlets, outlets = iolets_for(geom, "NASH", "NASH")
    rank = None if R == 1 else G.basic_decomposition(geom, R)
    steps, dt = 16, physical_dt(0.8)
    got = np.zeros(5 * R)
    os.environ["HLB_MOCK_LOG"] = str(tmp_path / "calls.log")
    try:
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        arrs = [np.ascontiguousarray(geom.coords, np.int32), np.ascontiguousarray(geom.bsite, np.int64),
                np.ascontiguousarray(geom.btype, np.uint8), np.ascontiguousarray(geom.biolet, np.int32),
                np.ascontiguousarray(geom.bdist, np.float32), np.ascontiguousarray(geom.bnavail, np.uint8),
                np.ascontiguousarray(geom.bnormal, np.float32)]
        bd = np.ascontiguousarray(geom.block_dims, np.int32)
        rk = None if rank is None else np.ascontiguousarray(rank, np.int32)
        inr, outr = np.ascontiguousarray(np.stack(inlets)), np.ascontiguousarray(np.stack(outlets))
        rc = L.hreal_run(R, p(bd), int(geom.block_size), C.c_int64(geom.n_sites), p(arrs[0]), C.c_int64(arrs[1].size),
                         *[p(a) for a in arrs[1:]], None if rk is None else p(rk), C.c_double(dt), C.c_double(DX),
                         len(inlets), p(inr), len(outlets), p(outr), C.c_int64(steps), None, None, 1, 1, p(got), 0)
        assert rc == 0
    finally:
        del os.environ["HLB_MOCK_LOG"]
    got = got.reshape(R, 5)
    for r in range(R):
        assert got[r, 4] == 1.0, "no densities on rank %d after %d steps" % (r, steps)
        # lb::Stable: the recording ABI reports |du| = 1e-3 against a tolerance of 1e-9 x 0.01 -- stable, not converged
        assert got[r, 0] == 1.0
        # (as in the reference, the tracker's range always contains the reference density 1)
        assert got[r, 1] == 1.0 - 0.001 * R and got[r, 2] == 1.0 + 0.002 * R and got[r, 3] == 0.003 * R, got[r]
        log = open(str(tmp_path / "calls.log") + ".rank%d" % r).read().splitlines()
        assert not [ln for ln in log if ln.startswith("get_cache") or ln.startswith("get_f")]
        # as in the reference, the root of the tree never looks at sites of its own (PostSendToParent is for nodes
        # that have a parent, PhasedBroadcastRegular.h:142-158): only the other ranks ask their engines
        asked = log.count("monitor") >= 1 and any(ln.startswith("stability ") for ln in log)
        assert asked == (r > 0)


@pytest.mark.skipif(not os.path.isdir("/root/reference/Code") and not os.path.exists(os.path.join(BUILD, "libhost_lbm_real.so")),
                    reason="reference checkout absent and no prebuilt tests/_build/libhost_lbm_real.so")
@pytest.mark.parametrize("variant,Q,wall,inlet", [(1, 15, "SBB", "NASH"), (2, 27, "BFL", "NASH"), (3, 19, "BFL", "LADD")])
def test_the_reference_lbm_with_other_lattices_and_link_rules(tmp_path, variant, Q, wall, inlet):
    """The reference's lb::LBM instantiated with D3Q15 + simple bounce-back, D3Q27 + BFL, and D3Q19 + BFL with a Ladd
    velocity inlet (the reference's InOutLetParabolicVelocity, warm-up ramp on): two ranks each; the engines are
    created with those policies and that rank's tables, and stepped in lb.hpp's order."""
    import ctypes as C
    from hemelb_b200 import geometry as G
    build_host_binaries()
    L = C.CDLL(os.path.join(BUILD, "libhost_lbm_real.so"))
    geom, R = geometry("tree"), 2
    inlets, outlets = iolets_for(geom, inlet, "NASH")
    if inlet == "LADD":
        inlets[0][13] = 6  # InOutLetParabolicVelocity::SetWarmup
    rank = G.basic_decomposition(geom, R)
    doms = build_domains(geom, Q, rank, R)
    steps, dt = 2, physical_dt(0.8)
    os.environ["HLB_MOCK_LOG"] = str(tmp_path / "calls.log")
    try:
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        arrs = [np.ascontiguousarray(geom.coords, np.int32), np.ascontiguousarray(geom.bsite, np.int64),
                np.ascontiguousarray(geom.btype, np.uint8), np.ascontiguousarray(geom.biolet, np.int32),
                np.ascontiguousarray(geom.bdist, np.float32), np.ascontiguousarray(geom.bnavail, np.uint8),
                np.ascontiguousarray(geom.bnormal, np.float32)]
        bd = np.ascontiguousarray(geom.block_dims, np.int32)
        rk = np.ascontiguousarray(rank, np.int32)
        inr, outr = np.ascontiguousarray(np.stack(inlets)), np.ascontiguousarray(np.stack(outlets))
        rc = L.hreal_run(R, p(bd), int(geom.block_size), C.c_int64(geom.n_sites), p(arrs[0]), C.c_int64(arrs[1].size),
                         *[p(a) for a in arrs[1:]], p(rk), C.c_double(dt), C.c_double(DX), len(inlets), p(inr),
                         len(outlets), p(outr), C.c_int64(steps), None, None, WALLS[wall], 0, None, variant)
        assert rc == 0
    finally:
        del os.environ["HLB_MOCK_LOG"]
    for r, dom in enumerate(doms):
        log = open(str(tmp_path / "calls.log") + ".rank%d" % r).read().splitlines()
        kv = dict(x.split("=") for x in log[0].split()[1:])
        assert (int(kv["lattice"]), int(kv["kernel"]), int(kv["wall"]), int(kv["inlet"]), int(kv["outlet"])) == \
            (Q, 0, WALLS[wall], IOLETS[inlet], 0)
        assert (int(kv["n_sites"]), int(kv["shared"])) == (dom.N, dom.totalSharedFs)
        assert [int(x) for x in kv["mid"].split(",")] == [int(x) for x in dom.mid]
        assert [int(x) for x in kv["edge"].split(",")] == [int(x) for x in dom.edge]
        build = log[:log.index("finalise") + 1]
        assert build[1] == "set_neighbour_indices 0 %d first=%d" % (dom.N, int(dom.neighbour_indices(0, 1)[0]))
        if inlet == "LADD":  # the warm-up length has no getter: the host classes read it off the ramp
            assert any(ln.startswith("set_iolets 0 %d kind0=1 " % len(inlets)) and ln.endswith("warmup0=6") for ln in build)
        got = [ln for ln in log if ln not in build and not ln.startswith("set_f ") and not ln.startswith("comm_")][:-1]
        want_calls = expected_calls(dom, steps, len(inlets), len(outlets), 0)
        assert len(got) == len(want_calls)
        for g_, w_ in zip(got, want_calls):
            if isinstance(w_, tuple):
                assert g_.startswith("set_step_scalars t=%d mask=%d" % (w_[1], w_[2]))
            else:
                assert g_ == w_, (r, g_, w_)
