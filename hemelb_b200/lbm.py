"""Host-side mirror of the reference's ``lb::LBM<Traits>`` for the GPU engine.

Same phase names, same order of ranges and the same arguments as ``Code/lb/lb.hpp:162-314``; every
method forwards to the C ABI (``include/hemelb_b200.h``).  The header-only C++ policy classes in
``hemelb_b200/host/`` are what a HemeLB build links; this module is the same thing for Python
callers (the parity tests and ``bench.py``).
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import capi
from .capi import HlbConfig, check, lib, ptr
from .domain import RankDomain


class SimulationState:
    """``lb::SimulationState`` (Code/lb/SimulationState.cc:13-60): 1-indexed time step."""

    def __init__(self):
        self.time_step = 1

    def increment(self):
        self.time_step += 1

    def get_0_indexed_time_step(self):
        return self.time_step - 1


class BoundaryValues:
    """``lb::iolets::BoundaryValues`` as a host-side scalar provider (BoundaryValues.cc:108-165):
    ``get_boundary_density(i)`` evaluates the iolet at the 0-indexed time step."""

    def __init__(self, records, state: SimulationState):
        self.records = [np.asarray(r, np.float64) for r in records]
        self.state = state

    def __len__(self):
        return len(self.records)

    def get_boundary_density(self, i: int) -> float:
        r = self.records[i]
        if int(r[0]) == 1:  # InOutLetVelocity::GetDensity
            return 1.0
        mean, amp, phase, period, warmup, min_density = r[9], r[10], r[11], r[12], r[13], r[14]
        t = self.state.get_0_indexed_time_step()
        # InOutLetCosine::GetDensity, InOutLetCosine.cc:26-43
        w = 2.0 * 3.14159265358979323846264338327950288 / period
        target = mean + amp * math.cos(w * t + phase)
        if t >= warmup:
            return target
        fac = float(t) / float(warmup)
        return fac * target + (1.0 - fac) * min_density

    def densities(self) -> np.ndarray:
        return np.array([self.get_boundary_density(i) for i in range(len(self))], np.float64)


def prepare_boundary_objects(inlets, outlets):
    """``LBM::PrepareBoundaryObjects`` (lb.hpp:128-152): tell every iolet the minimum density."""
    def dmin(r):
        return 1.0 if int(r[0]) == 1 else r[9] - r[10]
    allr = list(inlets) + list(outlets)
    if not allr:
        return
    m = min(dmin(r) for r in allr)
    for r in allr:
        r[14] = m


def gzs_device_domain_needs(dd):
    """The GZS remote needs of a ``devdomain.DeviceDomain``: (need rows (n, 4): local site, direction, owner
    rank, key of the remote site; {owner rank: coordinates of the sites asked of it, each once, in ghost-row
    order}) -- the second is what the ranks all-gather (``GpuLBM.from_device_domain``)."""
    site, direction, owner, coords = dd.gzs_needs()
    if not site.size:
        return np.zeros((0, 4), np.int64), {}
    ext = dd.block_dims * dd.block_size
    key = (coords[:, 0] * ext[1] + coords[:, 1]) * ext[2] + coords[:, 2]  # its position in the lattice
    need = np.stack([site, direction.astype(np.int64), owner.astype(np.int64), key], 1)
    _, firsts = np.unique(need[:, 2:4], axis=0, return_index=True)
    firsts = np.sort(firsts)
    return need, {int(o): coords[firsts][owner[firsts] == o] for o in np.unique(owner)}


def gzs_ghost_rows(need):
    """Owner rank of every ghost row of a GZS need list ((n, 4): local site, direction, owner rank, owner
    site key): links with equal (owner rank, owner site) share a row; rows in order of first appearance
    (``hlb_gpu_set_gzs_remote``)."""
    need = np.asarray(need, np.int64).reshape(-1, 4)
    if not need.shape[0]:
        return np.zeros(0, np.int64)
    _, firsts = np.unique(need[:, 2:4], axis=0, return_index=True)
    return need[np.sort(firsts), 2]


class GpuLBM:
    """One rank's collide-and-stream engine on one B200."""

    def __init__(self, domain: RankDomain, kernel="LBGK", wall="SBB", inlet="NASH", outlet="NASH", tau=0.8,
                 inlets=(), outlets=(), device=0, chunk=1 << 21, reorder=True):
        self.L = lib()
        self.domain = domain
        self.Q = domain.Q
        self.N = domain.N
        self.S = domain.totalSharedFs
        self.state = SimulationState()
        self.inlet_values = BoundaryValues(inlets, self.state)
        self.outlet_values = BoundaryValues(outlets, self.state)
        self.cache_mask = 0
        cfg = HlbConfig()
        cfg.lattice = domain.Q
        cfg.kernel = capi.KERNELS[kernel]
        cfg.wall = capi.WALLS[wall]
        cfg.inlet = capi.IOLETS[inlet]
        cfg.outlet = capi.IOLETS[outlet]
        cfg.tau = tau
        cfg.device = device
        cfg.rank = domain.rank
        cfg.nranks = domain.nranks
        cfg.n_sites = domain.N
        for t in range(6):
            cfg.mid_count[t] = int(domain.mid[t])
            cfg.edge_count[t] = int(domain.edge[t])
        cfg.total_shared_fs = domain.totalSharedFs
        cfg.n_neighbours = int(domain.procs.shape[0])
        cfg.n_inlets = len(inlets)
        cfg.n_outlets = len(outlets)
        cfg.reorder = 1 if reorder else 0
        self.cfg = cfg
        h = C.c_void_p()
        check(self.L.hlb_gpu_create(C.byref(cfg), C.byref(h)))
        self.h = h
        L, N, Q = self.L, domain.N, domain.Q
        for s0 in range(0, N, chunk):
            n = min(chunk, N - s0)
            idx = np.ascontiguousarray(domain.neighbour_indices(s0, n), np.int64)
            check(L.hlb_gpu_set_neighbour_indices(h, C.c_int64(s0), C.c_int64(n), ptr(idx, C.c_int64)))
        # boundary-typed sites: two contiguous id ranges
        mid_total = int(domain.mid.sum())
        for first, n in ((int(domain.mid[0]), mid_total - int(domain.mid[0])),
                         (mid_total + int(domain.edge[0]), N - mid_total - int(domain.edge[0]))):
            if n <= 0:
                continue
            sl = slice(first, first + n)
            check(L.hlb_gpu_set_site_data(h, C.c_int64(first), C.c_int64(n),
                                          ptr(np.ascontiguousarray(domain.wallMask[sl]), C.c_uint32),
                                          ptr(np.ascontiguousarray(domain.ioletMask[sl]), C.c_uint32),
                                          ptr(np.ascontiguousarray(domain.ioletId[sl]), C.c_int32)))
            d = np.ascontiguousarray(domain.distance_to_wall(first, n))
            check(L.hlb_gpu_set_wall_distances(h, C.c_int64(first), C.c_int64(n), ptr(d, C.c_double)))
            nr = np.ascontiguousarray(domain.wall_normal(first, n))
            check(L.hlb_gpu_set_wall_normals(h, C.c_int64(first), C.c_int64(n), ptr(nr, C.c_double)))
            if not reorder:
                gc = np.ascontiguousarray(domain.globalCoords[sl], np.int64)
                check(L.hlb_gpu_set_site_coords(h, C.c_int64(first), C.c_int64(n), ptr(gc, C.c_int64)))
        if reorder:  # the renumbering needs every site's coordinates
            for s0 in range(0, N, chunk):
                n = min(chunk, N - s0)
                gc = np.ascontiguousarray(domain.globalCoords[s0:s0 + n], np.int64)
                check(L.hlb_gpu_set_site_coords(h, C.c_int64(s0), C.c_int64(n), ptr(gc, C.c_int64)))
        if cfg.n_neighbours:
            pr = domain.procs
            check(L.hlb_gpu_set_neighbours(h, ptr(np.ascontiguousarray(pr[:, 0], np.int32), C.c_int),
                                           ptr(np.ascontiguousarray(pr[:, 1], np.int64), C.c_int64),
                                           ptr(np.ascontiguousarray(pr[:, 2], np.int64), C.c_int64)))
            check(L.hlb_gpu_set_streaming_indices(h, ptr(np.ascontiguousarray(domain.streamingIndices, np.int64),
                                                         C.c_int64)))
        self._set_iolets(inlets, outlets)
        self.gzs_need = np.zeros((0, 4), np.int64)
        self.gzs_serve = np.zeros((0, 2), np.int64)
        if wall == "GZS" and domain.nranks > 1:
            need, serve = domain._builder.gzs_site_halo()
            self.gzs_need, self.gzs_serve = need[domain.rank], serve[domain.rank]
            nd, sv = self.gzs_need, self.gzs_serve
            if nd.shape[0]:
                check(L.hlb_gpu_set_gzs_remote(h, C.c_int64(nd.shape[0]),
                                               ptr(np.ascontiguousarray(nd[:, 0], np.int64), C.c_int64),
                                               ptr(np.ascontiguousarray(nd[:, 1], np.int32), C.c_int32),
                                               ptr(np.ascontiguousarray(nd[:, 2], np.int32), C.c_int32),
                                               ptr(np.ascontiguousarray(nd[:, 3], np.int64), C.c_int64)))
            if sv.shape[0]:
                check(L.hlb_gpu_set_gzs_serve(h, C.c_int64(sv.shape[0]),
                                              ptr(np.ascontiguousarray(sv[:, 0], np.int32), C.c_int32),
                                              ptr(np.ascontiguousarray(sv[:, 1], np.int64), C.c_int64)))
        self.gzs_row_owner = gzs_ghost_rows(self.gzs_need)
        check(L.hlb_gpu_finalise(h))

    def _set_iolets(self, inlets, outlets):
        for which, recs in ((0, inlets), (1, outlets)):
            if len(recs):
                r = np.ascontiguousarray(np.stack(recs), np.float64)
                check(self.L.hlb_gpu_set_iolets(self.h, which, len(recs), ptr(r, C.c_double)))

    @classmethod
    def from_device_domain(cls, dd, kernel="LBGK", wall="SBB", inlet="NASH", outlet="NASH", tau=0.8, inlets=(),
                           outlets=(), reorder=True, all_gather=None):
        """The engine for a ``devdomain.DeviceDomain``: the tables move device-to-device
        (``hlb_gpu_create_from_domain``); nothing of size N crosses the host.  GuoZhengShi walls on more
        than one rank need ``all_gather(obj) -> [obj of rank 0, ...]`` (collective) once, to tell the
        owners which sites the phase-0 site halo has to carry (NeighbouringDataManager::ShareNeeds)."""
        self = cls.__new__(cls)
        self.L = lib()
        self.domain = dd
        self.Q, self.N, self.S = dd.Q, dd.N, dd.totalSharedFs
        self.state = SimulationState()
        self.inlet_values = BoundaryValues(inlets, self.state)
        self.outlet_values = BoundaryValues(outlets, self.state)
        self.cache_mask = 0
        cfg = HlbConfig()
        cfg.kernel = capi.KERNELS[kernel]
        cfg.wall = capi.WALLS[wall]
        cfg.inlet = capi.IOLETS[inlet]
        cfg.outlet = capi.IOLETS[outlet]
        cfg.tau = tau
        cfg.n_inlets = len(inlets)
        cfg.n_outlets = len(outlets)
        cfg.reorder = 1 if reorder else 0
        self.cfg = cfg
        h = C.c_void_p()
        check(self.L.hlb_gpu_create_from_domain(dd.d, C.byref(cfg), C.byref(h)))
        self.h = h
        self._set_iolets(inlets, outlets)
        self.gzs_need = np.zeros((0, 4), np.int64)
        self.gzs_serve = np.zeros((0, 2), np.int64)
        if wall == "GZS" and dd.nranks > 1:
            if all_gather is None:
                raise ValueError("GuoZhengShi walls on several ranks: from_device_domain needs all_gather")
            need, asks = gzs_device_domain_needs(dd)
            # every rank learns which sites every other rank wants of it -- each once, in the order of the
            # requester's ghost rows
            wanted = all_gather(asks)
            sv_rank, sv_site = [], []
            for r, asks in enumerate(wanted):
                c = asks.get(dd.rank)
                if r == dd.rank or c is None or not len(c):
                    continue
                local = dd.lookup_sites(c)
                if (local < 0).any():
                    raise capi.HlbError("rank %d asks rank %d for a site it does not own" % (r, dd.rank))
                sv_rank.append(np.full(local.size, r, np.int32))
                sv_site.append(local)
            self.gzs_need = need
            if need.shape[0]:
                check(self.L.hlb_gpu_set_gzs_remote(h, C.c_int64(need.shape[0]),
                                                    ptr(np.ascontiguousarray(need[:, 0]), C.c_int64),
                                                    ptr(np.ascontiguousarray(need[:, 1], np.int32), C.c_int32),
                                                    ptr(np.ascontiguousarray(need[:, 2], np.int32), C.c_int32),
                                                    ptr(np.ascontiguousarray(need[:, 3]), C.c_int64)))
            if sv_rank:
                sr, ss = np.concatenate(sv_rank), np.concatenate(sv_site)
                self.gzs_serve = np.stack([sr.astype(np.int64), ss], 1)
                check(self.L.hlb_gpu_set_gzs_serve(h, C.c_int64(sr.size), ptr(np.ascontiguousarray(sr), C.c_int32),
                                                   ptr(np.ascontiguousarray(ss), C.c_int64)))
        self.gzs_row_owner = gzs_ghost_rows(self.gzs_need)
        check(self.L.hlb_gpu_finalise(h))
        return self

    # ---- multi-GPU ---------------------------------------------------------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        check(lib().hlb_gpu_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, unique_id: bytes):
        check(self.L.hlb_gpu_comm_init(self.h, C.c_char_p(unique_id)))

    # ---- FieldData ---------------------------------------------------------------------------
    def f_size(self):
        return self.N * self.Q + 1 + self.S

    def set_f(self, f, which=0):
        f = np.ascontiguousarray(f, np.float64)
        assert f.size == self.f_size()
        check(self.L.hlb_gpu_set_f(self.h, which, ptr(f, C.c_double)))

    def get_f(self, which=0):
        f = np.zeros(self.f_size())
        check(self.L.hlb_gpu_get_f(self.h, which, ptr(f, C.c_double)))
        return f

    def get_halo(self, which=1):
        out = np.zeros(max(self.S, 1))
        check(self.L.hlb_gpu_get_halo(self.h, which, ptr(out, C.c_double)))
        return out[:self.S]

    def set_halo(self, data, which=0):
        data = np.ascontiguousarray(data, np.float64)
        check(self.L.hlb_gpu_set_halo(self.h, which, ptr(data, C.c_double) if self.S else None))

    def set_equilibrium(self, rho=1.0, m=(0.0, 0.0, 0.0)):
        m = np.ascontiguousarray(m, np.float64)
        check(self.L.hlb_gpu_set_equilibrium(self.h, C.c_double(rho), ptr(m, C.c_double)))

    def swap_old_and_new(self):
        check(self.L.hlb_gpu_swap(self.h))

    # ---- streamer concept --------------------------------------------------------------------
    def _push_scalars(self):
        di = self.inlet_values.densities()
        do = self.outlet_values.densities()
        check(self.L.hlb_gpu_set_step_scalars(self.h, C.c_uint64(self.state.time_step),
                                              ptr(di, C.c_double) if di.size else None,
                                              ptr(do, C.c_double) if do.size else None, C.c_uint32(self.cache_mask)))

    def set_cache_mask(self, mask: int):
        self.cache_mask = mask

    def set_time(self, t: int):
        self.state.time_step = t

    def stream_and_collide(self, slot, first, count):
        self._push_scalars()
        check(self.L.hlb_gpu_stream_and_collide(self.h, slot, C.c_int64(first), C.c_int64(count)))

    def post_step(self, slot, first, count):
        check(self.L.hlb_gpu_post_step(self.h, slot, C.c_int64(first), C.c_int64(count)))

    # ---- phase 0: NeighbouringDataManager (GZS site halo) ----------------------------------------
    def exchange_site_halo(self):
        check(self.L.hlb_gpu_exchange_site_halo(self.h))

    def get_gzs_send(self):
        out = np.zeros(max(1, self.gzs_serve.shape[0] * self.Q))
        check(self.L.hlb_gpu_get_gzs_send(self.h, ptr(out, C.c_double)))
        return out[:self.gzs_serve.shape[0] * self.Q].reshape(-1, self.Q)

    def set_gzs_ghost(self, rows):
        """``rows``: (ghost rows, Q) -- one per entry of ``gzs_row_owner``."""
        rows = np.ascontiguousarray(rows, np.float64)
        check(self.L.hlb_gpu_set_gzs_ghost(self.h, ptr(rows, C.c_double) if rows.size else None))

    # ---- IteratedAction phases of LBM (lb.hpp:162-314) ------------------------------------------
    def request_comms(self):
        check(self.L.hlb_gpu_request_comms(self.h))

    def pre_send(self):
        self._push_scalars()
        d = self.domain
        off = int(d.mid.sum())
        for t in range(6):
            check(self.L.hlb_gpu_stream_and_collide(self.h, t, C.c_int64(off), C.c_int64(int(d.edge[t]))))
            off += int(d.edge[t])
        check(self.L.hlb_gpu_edge_done(self.h))

    def pre_receive(self):
        d = self.domain
        off = 0
        for t in range(6):
            check(self.L.hlb_gpu_stream_and_collide(self.h, t, C.c_int64(off), C.c_int64(int(d.mid[t]))))
            off += int(d.mid[t])

    def post_receive(self):
        d = self.domain
        check(self.L.hlb_gpu_copy_received(self.h))
        off = int(d.mid.sum())
        for t in range(6):
            check(self.L.hlb_gpu_post_step(self.h, t, C.c_int64(off), C.c_int64(int(d.edge[t]))))
            off += int(d.edge[t])
        off = 0
        for t in range(6):
            check(self.L.hlb_gpu_post_step(self.h, t, C.c_int64(off), C.c_int64(int(d.mid[t]))))
            off += int(d.mid[t])

    def end_iteration(self):
        pass

    def do_time_step(self):
        """One pass of StepManager's phase 1 for the LBM actor + SimulationMaster::DoTimeStep's tail
        (SimulationMaster.impl.h:169-220)."""
        self.exchange_site_halo()
        self.request_comms()
        self.pre_send()
        self.pre_receive()
        self.post_receive()
        self.end_iteration()
        self.swap_old_and_new()
        self.state.increment()

    # ---- whole steps inside the library -----------------------------------------------------------
    def step(self, n=1):
        check(self.L.hlb_gpu_set_step_scalars(self.h, C.c_uint64(self.state.time_step), self._dens(0), self._dens(1),
                                              C.c_uint32(self.cache_mask)))
        check(self.L.hlb_gpu_step(self.h, n))
        self.state.time_step += n

    def _dens(self, which):
        bv = self.outlet_values if which else self.inlet_values
        d = bv.densities()
        self._keep = getattr(self, "_keep", [None, None])
        self._keep[which] = d
        return ptr(d, C.c_double) if d.size else None

    def time_steps(self, n) -> float:
        ms = C.c_float()
        check(self.L.hlb_gpu_set_step_scalars(self.h, C.c_uint64(self.state.time_step), self._dens(0), self._dens(1),
                                              C.c_uint32(self.cache_mask)))
        check(self.L.hlb_gpu_time_steps(self.h, n, C.byref(ms)))
        self.state.time_step += n
        return float(ms.value)

    def time_steps_detail(self, n):
        """(total ms, bulk-kernel ms, bulk sites updated) for n whole steps, CUDA events."""
        ms, bms, bs = C.c_float(), C.c_float(), C.c_int64()
        check(self.L.hlb_gpu_set_step_scalars(self.h, C.c_uint64(self.state.time_step), self._dens(0), self._dens(1),
                                              C.c_uint32(self.cache_mask)))
        check(self.L.hlb_gpu_time_steps_detail(self.h, n, C.byref(ms), C.byref(bms), C.byref(bs)))
        self.state.time_step += n
        return float(ms.value), float(bms.value), int(bs.value)

    def sync(self):
        check(self.L.hlb_gpu_sync(self.h))

    def set_overlap(self, enabled: bool):
        """True: the product schedule (fused mid-domain kernel / second stream, as created); False: every
        range its own kernel back to back on one stream.  Results are identical."""
        check(self.L.hlb_gpu_set_overlap(self.h, 1 if enabled else 0))

    def monitor(self):
        out = np.zeros(4)
        check(self.L.hlb_gpu_monitor(self.h, ptr(out, C.c_double)))
        return dict(min_f=out[0], min_density=out[1], max_density=out[2], max_speed=out[3])

    def monitor_begin(self):
        """First half of :meth:`monitor`: enqueue the reduction and the 32-byte copy, return at once."""
        check(self.L.hlb_gpu_monitor_begin(self.h))

    def monitor_end(self):
        """Second half: wait for the values asked for by :meth:`monitor_begin`."""
        out = np.zeros(4)
        check(self.L.hlb_gpu_monitor_end(self.h, ptr(out, C.c_double)))
        return dict(min_f=out[0], min_density=out[1], max_density=out[2], max_speed=out[3])

    def monitor_global(self):
        """The same extrema over all ranks (one ncclAllReduce; collective)."""
        out = np.zeros(4)
        check(self.L.hlb_gpu_monitor_global(self.h, ptr(out, C.c_double)))
        return dict(min_f=out[0], min_density=out[1], max_density=out[2], max_speed=out[3])

    def stability(self, with_convergence=False):
        """``StabilityTester``'s site loop on the device, before the swap: (populations of f_new that fail
        ``value > 0``, largest |u_new - u_old| over the local sites)."""
        out = np.zeros(2)
        check(self.L.hlb_gpu_stability(self.h, 1 if with_convergence else 0, ptr(out, C.c_double)))
        return int(out[0]), float(out[1])

    def launch_count(self) -> int:
        n = C.c_int64()
        check(self.L.hlb_gpu_launch_count(self.h, C.byref(n)))
        return int(n.value)

    def target_runs(self):
        """(groups of 32 device sites whose streaming targets the whole-part launches read in run form,
        all groups): hlb_gpu_target_runs."""
        a, b = C.c_int64(), C.c_int64()
        check(self.L.hlb_gpu_target_runs(self.h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def get_cache(self, name):
        bit = capi.CACHES[name]
        out = np.zeros(capi.CACHE_WIDTH[bit] * self.N)
        check(self.L.hlb_gpu_get_cache(self.h, C.c_uint32(bit), ptr(out, C.c_double)))
        return out

    def get_neighbour_indices(self):
        out = np.zeros(self.N * self.Q, np.int64)
        check(self.L.hlb_gpu_get_neighbour_indices(self.h, ptr(out, C.c_int64)))
        return out

    def close(self):
        if getattr(self, "h", None):
            self.L.hlb_gpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
