"""GPU extraction / checkpoint path (hlb_xtr_*, hlb_gpu_load_distributions; through the C ABI and
the host mirror hemelb_b200/extraction.py) against the golden files the compiled reference wrote
and against the oracle (restated streamers + oracle/xtr.py): byte-identical files."""
import os
import struct
import threading

import numpy as np
import pytest

from oracle import xtr as X
from tests.cases import geometry, iolets_for
from tests.xtr_cases import (ALL_FIELDS, CASES, CHECKPOINT_FIELDS, DT, DX, ORIGIN, REF_PRESSURE, RHO, initial_f, make_sim,
                             rank_data, rank_of, tau, xfields)

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class ThreadComm:
    """R ranks = R threads of this process sharing one GPU (collectives through a barrier)."""

    class World:
        def __init__(self, n):
            self.n, self.barrier, self.slots = n, threading.Barrier(n), [None] * n

    def __init__(self, world, rank):
        self.w, self.rank, self.size = world, rank, world.n

    def _gather(self, v):
        self.w.slots[self.rank] = v
        self.w.barrier.wait()
        out = list(self.w.slots)
        self.w.barrier.wait()
        return out

    def allreduce_sum(self, v):
        return sum(self._gather(v))

    def scan_sum(self, v):
        return sum(self._gather(v)[:self.rank + 1])

    def broadcast(self, v, root=0):
        return self._gather(v)[root]

    def scatter(self, values, root=0):
        return self._gather(values)[root][self.rank]

    def barrier(self):
        self.w.barrier.wait()


def run_ranks(R, fn):
    errs = [None] * R
    world = ThreadComm.World(R)

    def body(r):
        try:
            fn(r, ThreadComm(world, r))
        except BaseException as e:  # noqa: BLE001
            errs[r] = e
            world.barrier.abort()

    th = [threading.Thread(target=body, args=(r,)) for r in range(R)]
    [t.start() for t in th]
    [t.join() for t in th]
    for e in errs:
        if e is not None and not isinstance(e, threading.BrokenBarrierError):
            raise e
    for e in errs:
        if e is not None:
            raise e


def make_gpus(case, reorder=True, device_domain=False):
    from hemelb_b200.domain import build_domains
    from hemelb_b200.lbm import GpuLBM
    geom = geometry(case["geom"])
    Q, R = case["Q"], case["R"]
    inlets, outlets = iolets_for(geom, "NASH", "NASH")
    gpus = []
    if device_domain:
        from hemelb_b200.devdomain import DeviceDomain
        for r in range(R):
            dd = DeviceDomain.from_geometry(geom, Q, rank_of(case), r, R)
            gpus.append(GpuLBM.from_device_domain(dd, "LBGK", case["wall"], "NASH", "NASH", tau=tau(), inlets=inlets,
                                                  outlets=outlets, reorder=reorder))
    else:
        doms = build_domains(geom, Q, rank_of(case), R)
        for r in range(R):
            gpus.append(GpuLBM(doms[r], "LBGK", case["wall"], "NASH", "NASH", tau=tau(), inlets=inlets, outlets=outlets,
                               reorder=reorder))
    return gpus


def step_all(gpus, n):
    """n steps of R engines on one GPU with the halo staged through the host."""
    R = len(gpus)
    for _ in range(n):
        for g in gpus:
            g._push_scalars()
            g.pre_send()
        if R > 1:
            sent = [g.get_halo(1) for g in gpus]
            for r, g in enumerate(gpus):
                d = g.domain
                recv = np.zeros(d.totalSharedFs)
                base = d.N * d.Q + 1
                for (p, cnt, first) in np.asarray(d.procs).reshape(-1, 3):
                    od = gpus[int(p)].domain
                    for (q, ocnt, ofirst) in np.asarray(od.procs).reshape(-1, 3):
                        if int(q) == r:
                            ob = od.N * od.Q + 1
                            recv[first - base:first - base + cnt] = sent[int(p)][ofirst - ob:ofirst - ob + ocnt]
                g.set_halo(recv, 0)
        for g in gpus:
            g.pre_receive()
            g.post_receive()
            g.swap_old_and_new()
            g.state.increment()


def spec_of(case, path, fields=None):
    from hemelb_b200.extraction import OutputField, PropertyOutputFile
    fl = [OutputField(n, s, t, o) for (n, s, t, o) in (fields or case["fields"])]
    return PropertyOutputFile(str(path), case["frequency"], case["selector"], case["params"], fl)


def units():
    from hemelb_b200.extraction import Units
    return Units(DT, DX, ORIGIN, RHO, REF_PRESSURE)


def start(case, gpus):
    for r, g in enumerate(gpus):
        d = g.domain
        f = np.zeros(d.N * d.Q + 1 + d.totalSharedFs)
        geomf = initial_f({"N": d.N, "totalSharedFs": d.totalSharedFs}, d.Q, r)
        f[:] = geomf
        g.set_f(f)
        g.set_cache_mask(255)
    step_all(gpus, case["steps"])


@pytest.mark.parametrize("reorder", (True, False))
@pytest.mark.parametrize("name", sorted(CASES))
def test_gpu_writes_the_reference_files(tmp_path, name, reorder):
    """Same inputs as tests/golden/make_golden_xtr.py -> the same .xtr and .off bytes."""
    from hemelb_b200.extraction import GpuLocalPropertyOutput
    case = CASES[name]
    gpus = make_gpus(case, reorder)
    start(case, gpus)
    path = tmp_path / (name + ".xtr")
    outs = [None] * case["R"]

    def open_(r, comm):
        outs[r] = GpuLocalPropertyOutput(gpus[r], spec_of(case, path), units(), comm, chunk_sites=1000)
    run_ranks(case["R"], open_)
    t = case["steps"]
    for more in case["writes"]:
        step_all(gpus, more)
        t += more
        run_ranks(case["R"], lambda r, comm: outs[r].write(t, 1000))
    [o.close() for o in outs]
    with open(os.path.join(GOLD, "xtr_%s.xtr" % name), "rb") as fh:
        want = fh.read()
    got = path.read_bytes()
    assert len(got) == len(want)
    assert got == want
    with open(os.path.join(GOLD, "xtr_%s.off" % name), "rb") as fh:
        assert (tmp_path / (name + ".off")).read_bytes() == fh.read()


@pytest.mark.parametrize("device_domain", (False, True))
@pytest.mark.parametrize("selector,params", [
    ("whole", ()), ("surface", ()),
    ("plane", (0.034 + 8e-4, 0.001 + 8e-4, 0.074 + 10e-4, 0.3, -0.2, 1.0, 3.2e-4)),
    ("plane", (0.034 + 8e-4, 0.001 + 8e-4, 0.074 + 10e-4, 0.0, 0.0, 1.0, 0.0)),
    ("line", (0.034 + 8e-4, 0.001 + 8e-4, 0.074 - 20e-4, 0.034 + 9e-4, 0.001 + 8.5e-4, 0.074 + 40e-4)),
    ("surfacepoint", None)])
def test_selectors_and_all_fields_match_oracle(selector, params, device_domain):
    """Every selector x every field x 3 ranks: header, site lists and record bytes of each rank
    equal the oracle's; zero reference pressure exercises the NaN / inf traction of non-wall sites."""
    from hemelb_b200.extraction import GpuLocalPropertyOutput, OutputField, PropertyOutputFile, Units
    case = dict(geom="cylinder", Q=19, R=3, wall="BFL", steps=3, fields=ALL_FIELDS, selector=selector, params=params,
                frequency=1, writes=())
    gpus = make_gpus(case, True, device_domain)
    start(case, gpus)
    sim, T = make_sim("oracle", case)
    if params is None:  # a point a fraction of a voxel away from some wall site of rank 1
        w = np.flatnonzero(T[1]["wallMask"])[5]
        params = tuple(np.array(ORIGIN) + DX * (T[1]["globalCoords"].reshape(-1, 3)[w] + np.array([0.3, -0.2, 0.4])))
    for ref_p in (REF_PRESSURE, 0.0):
        conv = X.UnitConverter(DT, DX, ORIGIN, RHO, ref_p)
        po = X.PropertyOutput(xfields(ALL_FIELDS), selector, params, conv, 19, rank_data(sim, T, 19))
        if selector not in ("whole",):
            assert 0 < po.global_count < sum(t["N"] for t in T)
        for r, g in enumerate(gpus):
            fl = [OutputField(n, s, t, o) for (n, s, t, o) in ALL_FIELDS]
            spec = PropertyOutputFile(os.devnull + ".xtr", 1, selector, params, fl)
            from hemelb_b200 import extraction as E
            x = E._create_handle(g, spec, Units(DT, DX, ORIGIN, RHO, ref_p))
            out = GpuLocalPropertyOutput.__new__(GpuLocalPropertyOutput)
            out.L, out.x, out.fd = E.lib(), x, None
            n, sl, hl = (E.C.c_uint64(), E.C.c_uint64(), E.C.c_uint64())
            E.check(out.L.hlb_xtr_sizes(x, E.C.byref(n), E.C.byref(sl), E.C.byref(hl)))
            out.local_site_count, out.site_len = int(n.value), int(sl.value)
            assert out.local_site_count == po.local_counts[r]
            assert out.site_len == po.site_len and int(hl.value) == len(po.header)
            buf = (E.C.c_char * int(hl.value))()
            E.check(out.L.hlb_xtr_header(x, E.C.c_uint64(po.global_count), buf, E.C.c_uint64(int(hl.value))))
            assert bytes(buf) == po.header
            want = X.rank_chunk(po.fields, po.masks[r], po.rank_data[r], conv, 19, r)
            assert out.encode() == want, (selector, r, ref_p)
            out.close()


def test_integer_and_double_file_types():
    from hemelb_b200.extraction import GpuLocalPropertyOutput
    fields = [("P", "pressure", "int32", (3.0,)), ("P2", "pressure", "double", (0.5,)), ("R", "mpirank", "uint64", ()),
              ("R2", "mpirank", "float", ()), ("V", "velocity", "double", ()), ("S", "shearrate", "int64", ()),
              ("W", "shearstress", "double", ()), ("T", "stresstensor", "double", ()), ("D", "distributions", "float", ())]
    case = dict(geom="four_cube", Q=15, R=1, wall="SBB", steps=3, fields=fields, selector="whole", params=(), frequency=1)
    gpus = make_gpus(case)
    start(case, gpus)
    sim, T = make_sim("oracle", case)
    conv = X.UnitConverter(DT, DX, ORIGIN, RHO, REF_PRESSURE)
    po = X.PropertyOutput(xfields(fields), "whole", (), conv, 15, rank_data(sim, T, 15))
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        out = GpuLocalPropertyOutput(gpus[0], spec_of(case, os.path.join(d, "a.xtr")), units())
        out.write(0, 10)
        out.close()
        with open(os.path.join(d, "a.xtr"), "rb") as fh:
            assert fh.read() == po.header + po.record(0)


def test_checkpoint_restart_parity(tmp_path):
    """Write a checkpoint from the GPU, restart a fresh engine from it, continue: identical to the
    uninterrupted run; the reference-written golden checkpoint loads to the same state; and the
    oracle's reader accepts the GPU's file."""
    from hemelb_b200.extraction import GpuLocalDistributionInput, GpuLocalPropertyOutput
    case = CASES["cylinder_checkpoint_r2"]
    R, Q = case["R"], case["Q"]
    gpus = make_gpus(case)
    start(case, gpus)  # time step 4 is next (3 done)
    path = tmp_path / "ckpt.xtr"
    outs = [None] * R

    def open_(r, comm):
        outs[r] = GpuLocalPropertyOutput(gpus[r], spec_of(case, path), units(), comm)
    run_ranks(R, open_)
    run_ranks(R, lambda r, comm: outs[r].write(3, 1000))
    f3 = [g.get_f()[:g.N * Q].copy() for g in gpus]
    step_all(gpus, 4)
    run_ranks(R, lambda r, comm: outs[r].write(7, 1000))
    [o.close() for o in outs]
    want_end = [g.get_f()[:g.N * Q].copy() for g in gpus]
    # the oracle reads the GPU-written file
    coords = [np.asarray(g.domain.globalCoords) for g in gpus]
    t, f = X.load_checkpoint(path.read_bytes(), (tmp_path / "ckpt.off").read_bytes(), Q, coords, 3)
    assert t == 3
    for r in range(R):
        assert np.array_equal(f[r].ravel(), f3[r])
    # restart fresh engines from time step 3 of the GPU file and from the reference's golden file
    for src in (path, os.path.join(GOLD, "xtr_cylinder_checkpoint_r2.xtr")):
        fresh = make_gpus(case)
        got_t = [None] * R

        def load(r, comm):
            got_t[r] = GpuLocalDistributionInput(src, None, comm).load_distribution(fresh[r], 3)
        run_ranks(R, load)
        assert got_t == [3] * R
        for r, g in enumerate(fresh):
            assert np.array_equal(g.get_f(0)[:g.N * Q], f3[r])
            assert np.array_equal(g.get_f(1)[:g.N * Q], f3[r])
            g.set_time(4)  # SimulationState after 3 completed steps (1-indexed)
        step_all(fresh, 4)
        for r, g in enumerate(fresh):
            assert np.array_equal(g.get_f()[:g.N * Q], want_end[r])
    # "use the last one" (no target) -> time step 7
    fresh = make_gpus(case)
    got_t = [None] * R
    run_ranks(R, lambda r, comm: got_t.__setitem__(r, GpuLocalDistributionInput(path, None, comm).load_distribution(fresh[r])))
    assert got_t == [7] * R
    for r, g in enumerate(fresh):
        assert np.array_equal(g.get_f()[:g.N * Q], want_end[r])


def test_error_behaviour(tmp_path):
    from hemelb_b200.capi import HlbError
    from hemelb_b200.extraction import GpuLocalDistributionInput, GpuLocalPropertyOutput
    case = dict(CASES["four_cube_all"])
    gpus = make_gpus(case)
    g = gpus[0]
    g.set_f(initial_f({"N": g.N, "totalSharedFs": 0}, 15, 0))
    # a field whose cache no step has produced yet
    out = GpuLocalPropertyOutput(g, spec_of(case, tmp_path / "a.xtr"), units())
    with pytest.raises(HlbError, match="cache"):
        out.write(0, 10)
    out.close()
    # the file exists already: MPI_MODE_EXCL
    with pytest.raises(HlbError, match="cannot create"):
        GpuLocalPropertyOutput(g, spec_of(case, tmp_path / "a.xtr"), units())
    # offsets array of a wrong length (LocalPropertyOutput.cc:151-160)
    with pytest.raises(HlbError, match="Invalid length of offsets array 2"):
        GpuLocalPropertyOutput(g, spec_of(case, tmp_path / "b.xtr", [("V", "velocity", "float", (1.0, 2.0))]), units())
    # checkpoints: not a distributions file; wrong lattice; shuffled sites; truncated
    start(case, gpus)
    out = GpuLocalPropertyOutput(g, spec_of(case, tmp_path / "p.xtr", [("Pressure", "pressure", "float", (0.0,))]), units())
    out.write(0, 10)
    out.close()
    with pytest.raises(HlbError, match="field header must be 32 B long, but is 28 B"):
        GpuLocalDistributionInput(tmp_path / "p.xtr").load_distribution(g)
    out = GpuLocalPropertyOutput(g, spec_of(case, tmp_path / "c.xtr", CHECKPOINT_FIELDS), units())
    out.write(0, 10)
    out.close()
    assert GpuLocalDistributionInput(tmp_path / "c.xtr").load_distribution(g, 0) == 0
    with pytest.raises(HlbError, match="not found"):
        GpuLocalDistributionInput(tmp_path / "c.xtr").load_distribution(g, 5)
    raw = bytearray((tmp_path / "c.xtr").read_bytes())
    rec = 12 + 8 * 15
    h = 60 + 32 + 8
    raw[h:h + rec], raw[h + rec:h + 2 * rec] = raw[h + rec:h + 2 * rec], raw[h:h + rec]
    (tmp_path / "d.xtr").write_bytes(bytes(raw))
    (tmp_path / "d.off").write_bytes((tmp_path / "c.off").read_bytes())
    with pytest.raises(HlbError, match="Site read at index 0"):
        GpuLocalDistributionInput(tmp_path / "d.xtr").load_distribution(g)
    case19 = dict(case, Q=19)
    g19 = make_gpus(case19)[0]
    with pytest.raises(HlbError, match="contains 15 distributions"):
        GpuLocalDistributionInput(tmp_path / "c.xtr").load_distribution(g19)


def test_single_timestep_files_mode(tmp_path):
    """One file per written step, named as the reference names them (LocalPropertyOutput.cc:77-93,
    250-258; checked against the compiled reference in tests/test_xtr_oracle.py)."""
    from hemelb_b200.extraction import GpuLocalPropertyOutput, OutputField, PropertyOutputFile
    fields = [("Pressure", "pressure", "float", (80.0,)), ("Velocity", "velocity", "float", ())]
    case = dict(geom="cylinder", Q=19, R=2, wall="BFL", steps=3, fields=fields, selector="surface", params=(), frequency=2)
    gpus = make_gpus(case)
    start(case, gpus)
    sim, T = make_sim("oracle", case)
    outs = [None] * 2

    def open_(r, comm):
        spec = PropertyOutputFile(str(tmp_path / "snap_%d.xtr"), 2, "surface", (),
                                  [OutputField(n, s, t, o) for (n, s, t, o) in fields], single_timestep_files=True)
        outs[r] = GpuLocalPropertyOutput(gpus[r], spec, units(), comm)
    run_ranks(2, open_)
    for t in (4, 5, 120):
        run_ranks(2, lambda r, comm: outs[r].write(t, 12345))
    [o.close() for o in outs]
    assert sorted(p.name for p in tmp_path.iterdir()) == ["snap_    4.xtr", "snap_  120.xtr", "snap_.off"]
    po = X.PropertyOutput(xfields(fields), "surface", (), X.UnitConverter(DT, DX, ORIGIN, RHO, REF_PRESSURE), 19,
                          rank_data(sim, T, 19))
    assert (tmp_path / "snap_    4.xtr").read_bytes() == po.header + po.record(4)
    assert (tmp_path / "snap_  120.xtr").read_bytes() == po.header + po.record(120)
    assert (tmp_path / "snap_.off").read_bytes() == po.offset_file()
