"""TEST INFRASTRUCTURE ONLY -- ctypes bindings of the CPU oracle (``libhemelb_oracle.so``, our
restatement) and of ``_ref/libhemelb_ref*.so`` (the unmodified reference headers compiled by
``oracle/Makefile``).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
cpu_baseline / ``--impl reference`` legs may import this package; ``hemelb_b200`` never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
KERNELS = {"LBGK": 0, "MRT": 1, "TRT": 2}
WALLS = {"SBB": 0, "BFL": 1, "GZS": 2}
IOLETS = {"NASH": 0, "LADD": 1}
CACHE_BITS = {"density": 1, "velocity": 2, "wall_shear_stress": 4, "von_mises": 8, "shear_rate": 16,
              "stress_tensor": 32, "traction": 64, "tangential_traction": 128}
TABLE_DTYPES = {
    "counts": np.int64, "neighbourIndices": np.int64, "wallMask": np.uint32, "ioletMask": np.uint32,
    "siteType": np.int32, "ioletId": np.int32, "distanceToWall": np.float64, "wallNormal": np.float64,
    "globalCoords": np.int64, "inputIndex": np.int64, "streamingIndices": np.int64,
}


def build(force: bool = False) -> None:
    """Compile the oracle (and oracle/_ref when /root/reference is present)."""
    if force:
        subprocess.run(["make", "-C", HERE, "clean"], check=True, capture_output=True)
    subprocess.run(["make", "-C", HERE, "all"], check=True, capture_output=True)


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _d(a):
    return _ptr(a, C.c_double)


_oracle = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        path = os.path.join(HERE, "libhemelb_oracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.hlbo_geometry_create.restype = C.c_void_p
        L.hlbo_sim_create.restype = C.c_void_p
        L.hlbo_domain_get.restype = C.c_int64
        L.hlbo_sim_get_cache.restype = C.c_int64
        L.hlbo_sim_f_size.restype = C.c_int64
        L.hlbo_sim_get_time.restype = C.c_uint64
        L.hlbo_tau.restype = C.c_double
        L.hlbo_cosine_density.restype = C.c_double
        _oracle = L
    return _oracle


_refs = {}


def ref_lib(sse3=False):
    """The compiled reference, or None when oracle/_ref was not built/shipped.  ``sse3`` True: the build with the
    reference's x86-64 vector path; "mrtgzs": the build whose GuoZhengShi.h sets the wall node's m_neq before the
    collision (oracle/Makefile), for the one test that needs MRT + GuoZhengShi to be defined behaviour."""
    if sse3 not in _refs:
        name = {False: "libhemelb_ref.so", True: "libhemelb_ref_sse3.so", "mrtgzs": "libhemelb_ref_mrtgzs.so"}[sse3]
        path = os.path.join(HERE, "_ref", name)
        if not os.path.exists(path):
            _refs[sse3] = None
        else:
            L = C.CDLL(path)
            L.href_sim_create.restype = C.c_void_p
            L.href_sim_get_cache.restype = C.c_int64
            L.href_sim_f_size.restype = C.c_int64
            L.href_tau.restype = C.c_double
            L.href_sim_get_tau.restype = C.c_double
            L.href_cosine_density.restype = C.c_double
            if hasattr(L, "href_xtr_open"):
                L.href_xtr_open.restype = C.c_void_p
                L.href_xtr_last_error.restype = C.c_char_p
                L.href_xtr_string_length.restype = C.c_uint64
                L.href_xtr_field_header_length.restype = C.c_uint64
            _refs[sse3] = L
    return _refs[sse3]


_refdom = []


def ref_domain_lib():
    """oracle/_ref/libhemelb_refdom.so -- the reference's unmodified geometry::Domain, LookupTree, DistributedStore
    and BasicDecomposition over emulated ranks (oracle/ref_domain_driver.cc) -- or None when it was not
    built / shipped."""
    if not _refdom:
        path = os.path.join(HERE, "_ref", "libhemelb_refdom.so")
        L = None
        if os.path.exists(path):
            L = C.CDLL(path)
            L.hrefdom_run.restype = C.c_void_p
            L.hrefdom_get.restype = C.c_int64
            L.hrefdom_get.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_void_p]
            L.hrefdom_destroy.argtypes = [C.c_void_p]
        _refdom.append(L)
    return _refdom[0]


_reflbm = []

REF_LBM_VARIANTS = {  # oracle/ref_lbm_driver.cc rank_body: (Q, kernel, wall, inlet, outlet)
    (19, "LBGK", "BFL", "NASH", "NASH"): 0, (15, "LBGK", "SBB", "NASH", "NASH"): 1, (27, "LBGK", "BFL", "NASH", "NASH"): 2,
    (19, "LBGK", "BFL", "LADD", "NASH"): 3, (19, "LBGK", "GZS", "NASH", "NASH"): 4, (19, "LBGK", "SBB", "NASH", "NASH"): 5,
    (19, "LBGK", "GZS", "LADD", "NASH"): 6, (15, "LBGK", "BFL", "NASH", "NASH"): 7, (27, "LBGK", "SBB", "NASH", "NASH"): 8,
    (19, "MRT", "BFL", "LADD", "LADD"): 9, (15, "MRT", "SBB", "LADD", "LADD"): 10, (19, "LBGK", "BFL", "LADD", "LADD"): 11}


def ref_lbm_lib(sse3: bool = False):
    """oracle/_ref/libhemelb_reflbm.so -- the reference's unmodified lb::LBM<Traits> with its own CPU streamers,
    FieldData, Domain, NeighbouringDataManager, BoundaryValues, initial condition and StepManager over emulated
    ranks (oracle/ref_lbm_driver.cc) -- or None when it was not built / shipped."""
    while len(_reflbm) < 2:
        path = os.path.join(HERE, "_ref", "libhemelb_reflbm_sse3.so" if _reflbm else "libhemelb_reflbm.so")
        _reflbm.append(C.CDLL(path) if os.path.exists(path) else None)
    return _reflbm[1 if sse3 else 0]


def ref_lbm_run(geom, Q, wall, inlet, inlets, outlets, dt, dx, steps, sites_per_rank, rank_of_site=None, nranks=1,
                f0=None, equilibrium=None, kernel="LBGK", outlet="NASH", sse3=False, timing=None):
    """Runs the reference's whole lb::LBM for ``steps`` time steps on ``nranks`` emulated ranks and returns
    (per-rank f_old after the last swap [N_r * Q each, the Domain's site order], rank 0's iolet densities per step
    [steps, n_inlets + n_outlets]).  ``sites_per_rank``: the local fluid site counts the caller expects (the driver
    checks them against the reference Domain's); ``f0``: per-rank initial distributions, or ``equilibrium`` =
    (rho, (mx, my, mz)) for lb::EquilibriumInitialCondition; ``sse3``: the build with the reference's x86-64
    default vector path; ``timing``: a list that gets rank 0's wall-clock seconds around the step loop appended."""
    L = ref_lbm_lib(sse3)
    if L is None:
        raise RuntimeError("oracle/_ref/libhemelb_reflbm.so not built")
    bd = np.ascontiguousarray(geom.block_dims, np.int32)
    arrs = [np.ascontiguousarray(geom.coords, np.int32), np.ascontiguousarray(geom.bsite, np.int64),
            np.ascontiguousarray(geom.btype, np.uint8), np.ascontiguousarray(geom.biolet, np.int32),
            np.ascontiguousarray(geom.bdist, np.float32), np.ascontiguousarray(geom.bnavail, np.uint8),
            np.ascontiguousarray(geom.bnormal, np.float32)]
    rk = None if rank_of_site is None else np.ascontiguousarray(rank_of_site, np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    off = np.zeros(nranks + 1, np.int64)
    off[1:] = np.cumsum(np.asarray(sites_per_rank, np.int64) * Q)
    fin = None
    if f0 is not None:
        fin = np.ascontiguousarray(np.concatenate([np.asarray(f, np.float64)[:int(n) * Q] for f, n in zip(f0, sites_per_rank)]))
        assert fin.size == off[-1]
    rho, mom = (1.0, np.zeros(3)) if equilibrium is None else (float(equilibrium[0]), np.ascontiguousarray(equilibrium[1], np.float64))
    ri, ni = _recs(list(inlets))
    ro, no = _recs(list(outlets))
    out = np.zeros(int(off[-1]))
    n_local = np.zeros(nranks, np.int64)
    dens = np.zeros((max(steps, 1), ni + no))
    secs = np.zeros(1)
    rc = L.hreflbm_run(nranks, p(bd), int(geom.block_size), C.c_int64(geom.n_sites), p(arrs[0]), C.c_int64(arrs[1].size),
                       *[p(a) for a in arrs[1:]], None if rk is None else p(rk), C.c_double(dt), C.c_double(dx),
                       ni, _d(ri), no, _d(ro), C.c_int64(steps), REF_LBM_VARIANTS[(Q, kernel, wall, inlet, outlet)],
                       0 if equilibrium is None else 1, C.c_double(rho), _d(mom), None if fin is None else _d(fin),
                       p(off), _d(out), p(n_local), _d(dens), _d(secs))
    if timing is not None:
        timing.append(float(secs[0]))
    if rc != 0:
        raise RuntimeError("hreflbm_run failed (%d)" % rc)
    assert [int(x) for x in n_local] == [int(x) for x in sites_per_rank], (n_local, sites_per_rank)
    return [out[off[r]:off[r + 1]].copy() for r in range(nranks)], dens


class RefDomains:
    """Per-rank tables from the reference's own ``geometry::Domain`` (R emulated ranks).  ``rank_of_site``
    None: the blocks go to ranks by the reference's ``BasicDecomposition`` (``block_rank`` = its answer per
    .gmy block, SITE_OR_BLOCK_SOLID = INT_MIN for solid blocks)."""

    def __init__(self, geom, Q, rank_of_site=None, nranks=1):
        L = ref_domain_lib()
        if L is None:
            raise RuntimeError("oracle/_ref/libhemelb_refdom.so not built")
        self.L, self.Q, self.R = L, Q, nranks
        bd = np.ascontiguousarray(geom.block_dims, np.int32)
        arrs = [np.ascontiguousarray(geom.coords, np.int32), np.ascontiguousarray(geom.bsite, np.int64),
                np.ascontiguousarray(geom.btype, np.uint8), np.ascontiguousarray(geom.biolet, np.int32),
                np.ascontiguousarray(geom.bdist, np.float32), np.ascontiguousarray(geom.bnavail, np.uint8),
                np.ascontiguousarray(geom.bnormal, np.float32)]
        rk = None if rank_of_site is None else np.ascontiguousarray(rank_of_site, np.int32)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        self.h = C.c_void_p(L.hrefdom_run(Q, nranks, p(bd), int(geom.block_size), C.c_int64(geom.n_sites), p(arrs[0]),
                                          C.c_int64(arrs[1].size), *[p(a) for a in arrs[1:]], None if rk is None else p(rk)))
        if not self.h:
            raise ValueError("more ranks than non-empty blocks: the reference's BasicDecomposition refuses")
        self.block_rank = self._get(0, "blockRank", np.int32)

    def _get(self, r, name, dt):
        n = self.L.hrefdom_get(self.h, r, name.encode(), None)
        a = np.zeros(n, dt)
        self.L.hrefdom_get(self.h, r, name.encode(), a.ctypes.data_as(C.c_void_p))
        return a

    def tables(self, r=0):
        out = {"N": int(self.L.hrefdom_get(self.h, r, b"N", None)),
               "totalSharedFs": int(self.L.hrefdom_get(self.h, r, b"totalSharedFs", None)), "Q": self.Q}
        for name, dt in TABLE_DTYPES.items():
            if name != "inputIndex":  # (the reference keeps global coordinates, not the .gmy index)
                out[name] = self._get(r, name, dt)
        out["procs"] = self._get(r, "procs", np.int64).reshape(-1, 3)
        out["mid"] = out["counts"][:6].copy()
        out["edge"] = out["counts"][6:].copy()
        return out

    def __del__(self):
        try:
            self.L.hrefdom_destroy(self.h)
        except Exception:
            pass


def iolet_record(kind=0, normal=(0, 0, 1), position=(0, 0, 0), radius=1.0, max_speed=0.0,
                 density_mean=1.0, density_amp=0.0, phase=0.0, period=1000.0, warmup=0.0,
                 min_density=1.0):
    """The 16-double iolet descriptor shared by the oracle, the reference driver and the C ABI."""
    return np.array([kind, *normal, *position, radius, max_speed, density_mean, density_amp, phase,
                     period, warmup, min_density, 0.0], np.float64)


def _recs(lst):
    if not lst:
        return np.zeros((1, 16), np.float64), 0
    return np.ascontiguousarray(np.stack(lst)), len(lst)


# ------------------------------------------------------------------------------- pointwise
def lattice(Q, lib=None):
    c = np.zeros((Q, 3), np.int32)
    w = np.zeros(Q)
    inv = np.zeros(Q, np.int32)
    if lib is None:
        oracle_lib().hlbo_lattice(Q, _ptr(c, C.c_int), _d(w), _ptr(inv, C.c_int))
    else:
        lib.href_lattice(Q, _ptr(c, C.c_int), _d(w), _ptr(inv, C.c_int))
    return c, w, inv


def collide(Q, kernel, tau, f):
    f = np.ascontiguousarray(f, np.float64)
    fpost, feq, fneq, rmu = np.zeros(Q), np.zeros(Q), np.zeros(Q), np.zeros(7)
    oracle_lib().hlbo_collide(Q, KERNELS[kernel], C.c_double(tau), _d(f), _d(fpost), _d(feq), _d(fneq), _d(rmu))
    return dict(fpost=fpost, feq=feq, fneq=fneq, rho=rmu[0], m=rmu[1:4].copy(), u=rmu[4:7].copy())


def ref_collide(lib, Q, kernel, dt, dx, rho, eta, f, rates=None):
    f = np.ascontiguousarray(f, np.float64)
    fpost, feq, fneq, rmu = np.zeros(Q), np.zeros(Q), np.zeros(Q), np.zeros(7)
    r = None if rates is None else _d(np.ascontiguousarray(rates, np.float64))
    rc = lib.href_collide(Q, KERNELS[kernel], C.c_double(dt), C.c_double(dx), C.c_double(rho),
                          C.c_double(eta), _d(f), _d(fpost), _d(feq), _d(fneq), _d(rmu), r)
    if rc:
        raise ValueError("combination not buildable from the reference")
    return dict(fpost=fpost, feq=feq, fneq=fneq, rho=rmu[0], m=rmu[1:4].copy(), u=rmu[4:7].copy())


def stress_functions(Q, rho, tau, fneq, normal, lib=None):
    out = np.zeros(18)
    fneq = np.ascontiguousarray(fneq, np.float64)
    normal = np.ascontiguousarray(normal, np.float64)
    fn = oracle_lib().hlbo_stress_functions if lib is None else lib.href_stress_functions
    fn(Q, C.c_double(rho), C.c_double(tau), _d(fneq), _d(normal), _d(out))
    return dict(von_mises=out[0], wall_shear_stress=out[1], shear_rate=out[2],
                stress_tensor=out[3:12].copy(), traction=out[12:15].copy(), tangential_traction=out[15:18].copy())


# ------------------------------------------------------------------------------- tables
class OracleDomains:
    """Per-rank ``geometry::Domain`` tables built by the oracle from a Geometry + site->rank map."""

    def __init__(self, geom, Q, rank_of_site=None, nranks=1):
        L = oracle_lib()
        self.Q, self.R = Q, nranks
        rank = np.zeros(geom.n_sites, np.int32) if rank_of_site is None else np.ascontiguousarray(rank_of_site, np.int32)
        bd = np.ascontiguousarray(geom.block_dims, np.int32)
        coords = np.ascontiguousarray(geom.coords, np.int32)
        self._keep = (bd, coords, rank)
        self.h = C.c_void_p(L.hlbo_geometry_create(
            Q, nranks, _ptr(bd, C.c_int), geom.block_size, C.c_int64(geom.n_sites), _ptr(coords, C.c_int32),
            _ptr(rank, C.c_int32), C.c_int64(geom.bsite.size),
            _ptr(np.ascontiguousarray(geom.bsite, np.int64), C.c_int64),
            _ptr(np.ascontiguousarray(geom.btype, np.uint8), C.c_uint8),
            _ptr(np.ascontiguousarray(geom.biolet, np.int32), C.c_int32),
            _ptr(np.ascontiguousarray(geom.bdist, np.float32), C.c_float),
            _ptr(np.ascontiguousarray(geom.bnavail, np.uint8), C.c_uint8),
            _ptr(np.ascontiguousarray(geom.bnormal, np.float32), C.c_float)))

    def tables(self, r=0):
        L = oracle_lib()
        out = {"N": int(L.hlbo_domain_get(self.h, r, b"N", None)),
               "totalSharedFs": int(L.hlbo_domain_get(self.h, r, b"totalSharedFs", None)), "Q": self.Q}
        for name, dt in TABLE_DTYPES.items():
            n = L.hlbo_domain_get(self.h, r, name.encode(), None)
            a = np.zeros(n, dt)
            L.hlbo_domain_get(self.h, r, name.encode(), a.ctypes.data_as(C.c_void_p))
            out[name] = a
        n = L.hlbo_domain_get(self.h, r, b"procs", None)
        p = np.zeros(3 * n, np.int64)
        L.hlbo_domain_get(self.h, r, b"procs", p.ctypes.data_as(C.c_void_p))
        out["procs"] = p.reshape(n, 3)
        out["mid"] = out["counts"][:6].copy()
        out["edge"] = out["counts"][6:].copy()
        return out

    def __del__(self):
        try:
            oracle_lib().hlbo_geometry_destroy(self.h)
        except Exception:
            pass


class _SimCommon:
    prefix = ""

    def _fn(self, name):
        return getattr(self.L, self.prefix + name)

    def f_size(self, r=0):
        return int(self._fn("sim_f_size")(self.h, r))

    def set_f(self, f, r=0, which=0):
        f = np.ascontiguousarray(f, np.float64)
        assert f.size == self.f_size(r)
        self._fn("sim_set_f")(self.h, r, which, _d(f))

    def get_f(self, r=0, which=0):
        f = np.zeros(self.f_size(r))
        self._fn("sim_get_f")(self.h, r, which, _d(f))
        return f

    def set_time(self, t):
        self._fn("sim_set_time")(self.h, C.c_uint64(t))

    def set_cache_mask(self, mask):
        self._fn("sim_set_cache_mask")(self.h, C.c_uint(mask))

    def get_cache(self, name, r=0):
        bit = CACHE_BITS[name]
        n = self._fn("sim_get_cache")(self.h, r, bit, None) if self.prefix == "hlbo_" else None
        if n is None:
            per = {1: 1, 2: 3, 4: 1, 8: 1, 16: 1, 32: 9, 64: 3, 128: 3}[bit]
            n = per * self.N[r]
        out = np.zeros(n)
        self._fn("sim_get_cache")(self.h, r, bit, _d(out))
        return out

    def stream_and_collide(self, slot, first, count, r=0):
        self._fn("sim_stream_and_collide")(self.h, r, slot, C.c_int64(first), C.c_int64(count))

    def post_step(self, slot, first, count, r=0):
        self._fn("sim_post_step")(self.h, r, slot, C.c_int64(first), C.c_int64(count))

    def step(self, n=1):
        self._fn("sim_step")(self.h, n)

    def step_mt(self, n=1):
        """one thread per emulated rank (reference driver only)"""
        self._fn("sim_step_mt")(self.h, n)


class OracleSim(_SimCommon):
    """The restated LBM phase loop over all emulated ranks of an OracleDomains."""
    prefix = "hlbo_"

    def __init__(self, domains, kernel="LBGK", wall="SBB", inlet="NASH", outlet="NASH", tau=0.8,
                 inlets=(), outlets=()):
        self.L = oracle_lib()
        self.domains = domains
        self.N = [int(self.L.hlbo_domain_get(domains.h, r, b"N", None)) for r in range(domains.R)]
        ri, ni = _recs(list(inlets))
        ro, no = _recs(list(outlets))
        self.h = C.c_void_p(self.L.hlbo_sim_create(domains.h, KERNELS[kernel], WALLS[wall], IOLETS[inlet],
                                                   IOLETS[outlet], C.c_double(tau), ni, _d(ri), no, _d(ro)))

    def set_equilibrium(self, rho=1.0, m=(0.0, 0.0, 0.0)):
        m = np.ascontiguousarray(m, np.float64)
        self.L.hlbo_sim_set_equilibrium(self.h, C.c_double(rho), _d(m))

    def __del__(self):
        try:
            self.L.hlbo_sim_destroy(self.h)
        except Exception:
            pass


class RefSim(_SimCommon):
    """The reference's own streamers/kernels (oracle/_ref) run over supplied tables."""
    prefix = "href_"

    def __init__(self, tables_per_rank, Q, kernel="LBGK", wall="SBB", inlet="NASH", outlet="NASH",
                 dt=1e-4, dx=1e-4, rho=1000.0, eta=0.004, inlets=(), outlets=(), sse3=False):
        self.L = ref_lib(sse3)
        if self.L is None:
            raise RuntimeError("oracle/_ref not built")
        ri, ni = _recs(list(inlets))
        ro, no = _recs(list(outlets))
        R = len(tables_per_rank)
        h = self.L.href_sim_create(Q, KERNELS[kernel], WALLS[wall], IOLETS[inlet], IOLETS[outlet],
                                   C.c_double(dt), C.c_double(dx), C.c_double(rho), C.c_double(eta), R,
                                   ni, _d(ri), no, _d(ro))
        if not h:
            raise ValueError("combination not buildable from the reference")
        self.h = C.c_void_p(h)
        self.N = []
        for r, t in enumerate(tables_per_rank):
            self.N.append(int(t["N"]))
            procs = np.ascontiguousarray(t["procs"], np.int64).reshape(-1)
            self.L.href_sim_set_domain(
                self.h, r, C.c_int64(t["N"]), _ptr(np.ascontiguousarray(t["counts"], np.int64), C.c_int64),
                _ptr(np.ascontiguousarray(t["neighbourIndices"], np.int64), C.c_int64),
                _ptr(np.ascontiguousarray(t["wallMask"], np.uint32), C.c_uint32),
                _ptr(np.ascontiguousarray(t["ioletMask"], np.uint32), C.c_uint32),
                _ptr(np.ascontiguousarray(t["siteType"], np.int32), C.c_int32),
                _ptr(np.ascontiguousarray(t["ioletId"], np.int32), C.c_int32),
                _d(np.ascontiguousarray(t["distanceToWall"], np.float64)),
                _d(np.ascontiguousarray(t["wallNormal"], np.float64)),
                _ptr(np.ascontiguousarray(t["globalCoords"], np.int64), C.c_int64),
                _ptr(np.ascontiguousarray(t["inputIndex"], np.int64), C.c_int64),
                C.c_int64(t["totalSharedFs"]), int(t["procs"].shape[0]),
                _ptr(procs if procs.size else np.zeros(3, np.int64), C.c_int64),
                _ptr(np.ascontiguousarray(t["streamingIndices"], np.int64) if t["totalSharedFs"] else np.zeros(1, np.int64), C.c_int64))
        self.L.href_sim_init(self.h)
        self.tau = float(self.L.href_sim_get_tau(self.h))

    def __del__(self):
        try:
            self.L.href_sim_destroy(self.h)
        except Exception:
            pass

    # ---- extraction / checkpoint through the reference's own sources (oracle/ref_xtr.h)
    def xtr_open(self, path, fields, selector="whole", sel_params=(), frequency=1, single_timestep_files=False,
                 dt=1e-4, dx=1e-4, origin=(0.0, 0.0, 0.0), fluid_density=1000.0, reference_pressure=0.0):
        """fields: oracle.xtr.Field list.  Constructs one LocalPropertyOutput per emulated rank (writes
        the header and the .off file); returns a session for xtr_write."""
        from . import xtr as X
        names = (C.c_char_p * len(fields))(*[f.name.encode() for f in fields])
        src = np.array([X.SOURCES.index(f.source) for f in fields], np.int32)
        tc = np.array([f.typecode for f in fields], np.int32)
        noff = np.array([len(f.offsets) for f in fields], np.int32)
        offs = np.array([o for f in fields for o in f.offsets] + [0.0], np.float64)
        sel = np.zeros(8, np.float32)
        sel[:len(sel_params)] = sel_params
        org = np.ascontiguousarray(origin, np.float64)
        kind = X.SELECTORS[selector]
        h = self.L.href_xtr_open(self.h, str(path).encode(), C.c_uint64(frequency), int(single_timestep_files), kind,
                                 _ptr(sel, C.c_float), len(fields), names, _ptr(src, C.c_int), _ptr(tc, C.c_int),
                                 _ptr(noff, C.c_int), _d(offs), C.c_double(dt), C.c_double(dx), _d(org),
                                 C.c_double(fluid_density), C.c_double(reference_pressure))
        if not h:
            raise RuntimeError(self.L.href_xtr_last_error().decode())
        return C.c_void_p(h)

    def xtr_write(self, session, timestep, total_steps=1000):
        if self.L.href_xtr_write(session, C.c_uint64(timestep), C.c_uint64(total_steps)):
            raise RuntimeError(self.L.href_xtr_last_error().decode())

    def xtr_close(self, session):
        self.L.href_xtr_close(session)

    def load_checkpoint(self, xtr_path, off_path=None, target=None):
        t = C.c_uint64(0)
        rc = self.L.href_load_checkpoint(self.h, str(xtr_path).encode(), (str(off_path) if off_path else "").encode(),
                                         C.c_int64(-1 if target is None else target), C.byref(t))
        if rc:
            raise RuntimeError(self.L.href_xtr_last_error().decode())
        return int(t.value)
