"""Device-side ``geometry::Domain`` construction (``hlb_dom_*`` in ``include/hemelb_b200.h``).

``DeviceDomain`` is the GPU counterpart of ``domain.RankDomain``: the same tables
(``Code/geometry/Domain.cc:69-580``), built by kernels over a dense voxel window in HBM from either
an explicit ``.gmy``-level site list or an analytic capsule shape voxelised on the device, and
handed to the collide-and-stream engine without visiting the host.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import check, lib, ptr
from .geometry import Geometry, IoletPlane


class HlbDomConfig(C.Structure):
    _fields_ = [("lattice", C.c_int), ("block_size", C.c_int), ("block_dims", C.c_int64 * 3), ("rank", C.c_int),
                ("nranks", C.c_int), ("device", C.c_int)]


def capsule_array(A, B, R) -> np.ndarray:
    """(n, 7) doubles {a, b, radius} from end points and radii."""
    A, B, R = np.atleast_2d(A), np.atleast_2d(B), np.atleast_1d(R)
    return np.ascontiguousarray(np.concatenate([A, B, R[:, None]], 1), np.float64)


def iolet_array(iolets) -> np.ndarray:
    """(n, 9) doubles {kind, index, position, normal, radius} from ``geometry.IoletPlane``s."""
    out = np.zeros((len(iolets), 9))
    for k, io in enumerate(iolets):
        out[k] = [io.kind, io.index, *io.position, *io.normal, io.radius]
    return out


class DeviceDomain:
    """One rank's Domain tables, resident on one GPU."""

    def __init__(self, Q: int, block_size: int, block_dims, rank=0, nranks=1, device=0):
        self.L = lib()
        cfg = HlbDomConfig()
        cfg.lattice, cfg.block_size = Q, int(block_size)
        for k in range(3):
            cfg.block_dims[k] = int(block_dims[k])
        cfg.rank, cfg.nranks, cfg.device = rank, nranks, device
        self.cfg = cfg
        self.Q, self.rank, self.nranks, self.device = Q, rank, nranks, device
        self.block_size = int(block_size)
        self.block_dims = np.array([int(x) for x in block_dims], np.int64)
        d = C.c_void_p()
        check(self.L.hlb_dom_create(C.byref(cfg), C.byref(d)))
        self.d = d
        self.meta = {}
        self.N = 0

    # ---- sources -----------------------------------------------------------------------------
    @classmethod
    def from_geometry(cls, geom: Geometry, Q: int, rank_of_site=None, rank=0, nranks=1, device=0):
        self = cls(Q, geom.block_size, geom.block_dims, rank, nranks, device)
        coords = np.ascontiguousarray(geom.coords, np.int32)
        ros = None if rank_of_site is None else np.ascontiguousarray(rank_of_site, np.int32)
        nrec = int(geom.bsite.size)
        rs = np.ascontiguousarray(geom.bsite, np.int64)
        ty = np.ascontiguousarray(geom.btype, np.uint8)
        io = np.ascontiguousarray(geom.biolet, np.int32)
        di = np.ascontiguousarray(geom.bdist, np.float32)
        na = np.ascontiguousarray(geom.bnavail, np.uint8)
        no = np.ascontiguousarray(geom.bnormal, np.float32)
        check(self.L.hlb_dom_set_sites(
            self.d, C.c_int64(geom.n_sites), ptr(coords, C.c_int32), ptr(ros, C.c_int32) if ros is not None else None,
            C.c_int64(nrec), ptr(rs, C.c_int64), ptr(ty, C.c_uint8), ptr(io, C.c_int32), ptr(di, C.c_float),
            ptr(na, C.c_uint8), ptr(no, C.c_float)))
        self.meta = dict(geom.meta)
        return self.build()

    @classmethod
    def from_shape(cls, capsules, iolets, shape, Q: int, block_size=8, partition=None, rank=0, nranks=1, device=0,
                   build=True, roughness=None):
        """``capsules``: (n,7); ``iolets``: IoletPlanes; ``shape``: voxel extent of the lattice.
        ``partition``: None | ("slabs", axis, first_coord[nranks+1]) | ("blocks", rank_of_block).
        ``roughness``: None | (amplitude per capsule, (g,g,g) noise grid): wall roughness, see ``sac_shape``."""
        shape = np.asarray(shape, np.int64)
        bdims = (shape + block_size - 1) // block_size
        self = cls(Q, block_size, bdims, rank, nranks, device)
        caps = np.ascontiguousarray(capsules, np.float64).reshape(-1, 7)
        ios = np.ascontiguousarray(iolet_array(iolets), np.float64)
        check(self.L.hlb_dom_set_shape(self.d, caps.shape[0], ptr(caps, C.c_double), ios.shape[0],
                                       ptr(ios, C.c_double) if ios.size else None))
        self.meta = dict(inlets=[i for i in iolets if i.kind == 2], outlets=[i for i in iolets if i.kind == 3])
        if roughness is not None:
            amp = np.ascontiguousarray(roughness[0], np.float64)
            noise = np.ascontiguousarray(roughness[1], np.float64)
            ext = np.ascontiguousarray(shape, np.float64)
            assert amp.size == caps.shape[0] and noise.ndim == 3
            check(self.L.hlb_dom_set_roughness(self.d, ptr(amp, C.c_double), int(noise.shape[0]), ptr(noise, C.c_double),
                                               ptr(ext, C.c_double)))
        if partition is not None:
            self.set_partition(partition)
        return self.build() if build else self

    def set_partition(self, partition):
        if partition[0] == "slabs":
            first = np.ascontiguousarray(partition[2], np.int64)
            assert first.size == self.nranks + 1
            check(self.L.hlb_dom_set_partition_slabs(self.d, int(partition[1]), ptr(first, C.c_int64)))
        elif partition[0] == "blocks":
            rob = np.ascontiguousarray(partition[1], np.int32)
            assert rob.size == int(self.block_dims.prod())
            check(self.L.hlb_dom_set_partition_blocks(self.d, ptr(rob, C.c_int32)))
        else:
            raise ValueError(partition[0])

    def count_block_sites(self, lo=None, hi=None) -> np.ndarray:
        """Fluid sites per block over the block box [lo, hi) (default: the whole lattice)."""
        lo = np.zeros(3, np.int64) if lo is None else np.ascontiguousarray(lo, np.int64)
        hi = self.block_dims.copy() if hi is None else np.ascontiguousarray(hi, np.int64)
        out = np.zeros(tuple(int(x) for x in (hi - lo)), np.int32)
        check(self.L.hlb_dom_count_block_sites(self.d, ptr(lo, C.c_int64), ptr(hi, C.c_int64), ptr(out, C.c_int32)))
        return out

    def count_block_sites_typed(self, lo=None, hi=None):
        """(fluid sites, boundary-typed sites) per block over the block box [lo, hi)."""
        lo = np.zeros(3, np.int64) if lo is None else np.ascontiguousarray(lo, np.int64)
        hi = self.block_dims.copy() if hi is None else np.ascontiguousarray(hi, np.int64)
        shape = tuple(int(x) for x in (hi - lo))
        out, bnd = np.zeros(shape, np.int32), np.zeros(shape, np.int32)
        check(self.L.hlb_dom_count_block_sites_typed(self.d, ptr(lo, C.c_int64), ptr(hi, C.c_int64), ptr(out, C.c_int32),
                                                     ptr(bnd, C.c_int32)))
        return out, bnd

    def build(self):
        check(self.L.hlb_dom_build(self.d))
        n, S, nn = C.c_int64(), C.c_int64(), C.c_int()
        mid = np.zeros(6, np.int64)
        edge = np.zeros(6, np.int64)
        check(self.L.hlb_dom_get_counts(self.d, C.byref(n), ptr(mid, C.c_int64), ptr(edge, C.c_int64), C.byref(S),
                                        C.byref(nn)))
        self.N, self.totalSharedFs, self.mid, self.edge = int(n.value), int(S.value), mid, edge
        k = int(nn.value)
        r = np.zeros(max(k, 1), np.int32)
        c = np.zeros(max(k, 1), np.int64)
        f = np.zeros(max(k, 1), np.int64)
        check(self.L.hlb_dom_get_neighbours(self.d, ptr(r, C.c_int), ptr(c, C.c_int64), ptr(f, C.c_int64)))
        self.procs = np.stack([r[:k].astype(np.int64), c[:k], f[:k]], 1) if k else np.zeros((0, 3), np.int64)
        s = np.zeros(max(self.totalSharedFs, 1), np.int64)
        check(self.L.hlb_dom_get_streaming_indices(self.d, ptr(s, C.c_int64)))
        self.streamingIndices = s[:self.totalSharedFs]
        sec = C.c_double()
        check(self.L.hlb_dom_build_seconds(self.d, C.byref(sec)))
        self.build_seconds = float(sec.value)
        return self

    # ---- GuoZhengShi across ranks ---------------------------------------------------------------
    def gzs_needs(self):
        """(local site, direction, owner rank, neighbour coordinates (n,3)) of every GZS wall link that
        extrapolates from a site on another rank, ordered by owner rank, site, direction."""
        n = C.c_int64()
        check(self.L.hlb_dom_gzs_needs(self.d, C.c_int64(0), C.byref(n), None, None, None, None))
        k = int(n.value)
        site, direction = np.zeros(max(k, 1), np.int64), np.zeros(max(k, 1), np.int32)
        owner, coords = np.zeros(max(k, 1), np.int32), np.zeros((max(k, 1), 3), np.int64)
        if k:
            check(self.L.hlb_dom_gzs_needs(self.d, C.c_int64(k), C.byref(n), ptr(site, C.c_int64), ptr(direction, C.c_int32),
                                           ptr(owner, C.c_int32), ptr(coords, C.c_int64)))
        return site[:k], direction[:k], owner[:k], coords[:k]

    def lookup_sites(self, coords) -> np.ndarray:
        """Local site id of each global coordinate triple (-1: not a local fluid site)."""
        coords = np.ascontiguousarray(coords, np.int64).reshape(-1, 3)
        out = np.full(max(coords.shape[0], 1), -1, np.int64)
        if coords.shape[0]:
            check(self.L.hlb_dom_lookup_sites(self.d, C.c_int64(coords.shape[0]), ptr(coords, C.c_int64), ptr(out, C.c_int64)))
        return out[:coords.shape[0]]

    # ---- reference-form tables ------------------------------------------------------------------
    @property
    def NB(self):
        return int(self.N - self.mid[0] - self.edge[0])

    def boundary_sites(self) -> np.ndarray:
        mt = int(self.mid.sum())
        return np.concatenate([np.arange(int(self.mid[0]), mt), np.arange(mt + int(self.edge[0]), self.N)]).astype(np.int64)

    def neighbour_indices(self, first=0, n=None) -> np.ndarray:
        n = self.N - first if n is None else n
        out = np.zeros(max(n * self.Q, 1), np.int64)
        check(self.L.hlb_dom_get_neighbour_indices(self.d, C.c_int64(first), C.c_int64(n), ptr(out, C.c_int64)))
        return out[:n * self.Q]

    def global_coords(self, first=0, n=None) -> np.ndarray:
        n = self.N - first if n is None else n
        out = np.zeros((max(n, 1), 3), np.int64)
        check(self.L.hlb_dom_get_site_coords(self.d, C.c_int64(first), C.c_int64(n), ptr(out, C.c_int64)))
        return out[:n]

    def input_index(self) -> np.ndarray:
        out = np.zeros(max(self.N, 1), np.int64)
        check(self.L.hlb_dom_get_input_index(self.d, C.c_int64(0), C.c_int64(self.N), ptr(out, C.c_int64)))
        return out[:self.N]

    def boundary_tables(self) -> dict:
        nb, Q = max(self.NB, 1), self.Q
        w, i = np.zeros(nb, np.uint32), np.zeros(nb, np.uint32)
        ii = np.zeros(nb, np.int32)
        d, nr = np.zeros((nb, Q - 1)), np.zeros((nb, 3))
        check(self.L.hlb_dom_get_boundary_tables(self.d, ptr(w, C.c_uint32), ptr(i, C.c_uint32), ptr(ii, C.c_int32),
                                                 ptr(d, C.c_double), ptr(nr, C.c_double)))
        k = self.NB
        return dict(wallMask=w[:k], ioletMask=i[:k], ioletId=ii[:k], distanceToWall=d[:k], wallNormal=nr[:k])

    def tables(self) -> dict:
        """Everything, expanded to per-site arrays like ``RankDomain.tables()`` (bulk-typed sites
        carry no masks, distance -1 and normal +inf)."""
        N, Q = self.N, self.Q
        bt = self.boundary_tables()
        bs = self.boundary_sites()
        wall, iol = np.zeros(N, np.uint32), np.zeros(N, np.uint32)
        iid = np.full(N, -1, np.int32)
        dist, nrm = np.full((N, Q - 1), -1.0), np.full((N, 3), np.inf)
        wall[bs], iol[bs], iid[bs], dist[bs], nrm[bs] = (bt["wallMask"], bt["ioletMask"], bt["ioletId"],
                                                         bt["distanceToWall"], bt["wallNormal"])
        # geometry::SiteType of every site (FLUID 1, INLET 2, OUTLET 3) from its collision-type range
        stype = np.ones(N, np.int32)
        first = 0
        for counts in (self.mid, self.edge):
            for t in range(6):
                n = int(counts[t])
                if t in (2, 4):
                    stype[first:first + n] = 2
                elif t in (3, 5):
                    stype[first:first + n] = 3
                first += n
        return dict(N=N, totalSharedFs=self.totalSharedFs, counts=np.concatenate([self.mid, self.edge]),
                    mid=self.mid.copy(), edge=self.edge.copy(), neighbourIndices=self.neighbour_indices(),
                    Q=Q, wallMask=wall, ioletMask=iol, siteType=stype, ioletId=iid, distanceToWall=dist.reshape(-1),
                    inputIndex=np.arange(N, dtype=np.int64),  # (no input list behind an analytic shape)
                    wallNormal=nrm.reshape(-1), globalCoords=self.global_coords().reshape(-1),
                    streamingIndices=self.streamingIndices, procs=self.procs)

    def geometry(self) -> Geometry:
        """The device-voxelised sites and cut links as a ``Geometry`` (analytic source)."""
        n, nr = C.c_int64(), C.c_int64()
        check(self.L.hlb_dom_get_geometry_sizes(self.d, C.byref(n), C.byref(nr)))
        n, nr = int(n.value), int(nr.value)
        coords = np.zeros((max(n, 1), 3), np.int32)
        rs = np.zeros(max(nr, 1), np.int64)
        ty = np.zeros((max(nr, 1), 26), np.uint8)
        io = np.zeros((max(nr, 1), 26), np.int32)
        di = np.zeros((max(nr, 1), 26), np.float32)
        na = np.zeros(max(nr, 1), np.uint8)
        no = np.zeros((max(nr, 1), 3), np.float32)
        check(self.L.hlb_dom_get_geometry(self.d, ptr(coords, C.c_int32), ptr(rs, C.c_int64), ptr(ty, C.c_uint8),
                                          ptr(io, C.c_int32), ptr(di, C.c_float), ptr(na, C.c_uint8), ptr(no, C.c_float)))
        o = np.argsort(rs[:nr], kind="stable")
        g = Geometry(self.block_dims.astype(np.int32), self.block_size, coords[:n], rs[:nr][o], ty[:nr][o], io[:nr][o],
                     di[:nr][o], na[:nr][o], no[:nr][o], meta=dict(self.meta))
        return g.gmy_sort()

    def close(self):
        if getattr(self, "d", None):
            self.L.hlb_dom_destroy(self.d)
            self.d = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---------------------------------------------------------------------------------------------
# analytic shapes of the BASELINE configs
# ---------------------------------------------------------------------------------------------
def cylinder_shape(radius: float, length: int, margin: int = 2):
    """configs[1] as capsules: the cylinder of ``geometry.cylinder`` (axis z, inlet cap below z0,
    outlet cap above z1).  Returns (capsules, iolets, voxel shape)."""
    R = float(radius)
    n = int(np.ceil(2 * R)) + 2 * margin + 1
    c = (n - 1) / 2.0
    shape = (n, n, length + 2 * margin)
    z0, z1 = margin, margin + length - 1
    caps = capsule_array([[c, c, z0 - 2.0 * R - 16.0]], [[c, c, z1 + 2.0 * R + 16.0]], [R])
    iolets = [IoletPlane(2, 0, np.array([c, c, z0 - 0.5]), np.array([0.0, 0.0, 1.0]), R + 2),
              IoletPlane(3, 0, np.array([c, c, z1 + 0.5]), np.array([0.0, 0.0, -1.0]), R + 2)]
    return caps, iolets, shape


def sac_shape(radius: float, neck_radius: float, neck_length: float, roughness: float = 0.0, seed: int = 20261017,
              margin: int = 3):
    """configs[4] as capsules: the aneurysm-like sac of ``geometry.sac`` -- a sphere whose wall is displaced
    by seeded value noise, crossed by a neck cylinder along z that the iolet caps close.  Returns
    (capsules, iolets, voxel shape, roughness argument of ``DeviceDomain.from_shape``)."""
    rng = np.random.default_rng(seed)
    R = float(radius)
    n = int(np.ceil(2 * (R + roughness))) + 2 * margin + 1
    zlen = int(np.ceil(2 * R + 2 * neck_length)) + 2 * margin
    c = np.array([(n - 1) / 2.0, (n - 1) / 2.0, (zlen - 1) / 2.0])
    shape = (n, n, zlen)
    noise = rng.uniform(-1.0, 1.0, (8, 8, 8))
    far = 4.0 * zlen + 64.0
    caps = capsule_array([c, [c[0], c[1], c[2] - far]], [c, [c[0], c[1], c[2] + far]], [R, float(neck_radius)])
    zin, zout = margin - 0.5, zlen - margin - 0.5
    iolets = [IoletPlane(2, 0, np.array([c[0], c[1], zin]), np.array([0.0, 0.0, 1.0]), neck_radius + 2),
              IoletPlane(3, 0, np.array([c[0], c[1], zout]), np.array([0.0, 0.0, -1.0]), neck_radius + 2)]
    return caps, iolets, shape, ((float(roughness), 0.0), noise)


def tree_shape(generations: int, root_radius: float, root_length: float, seed: int = 20261017,
               half_angle_deg: float = 35.0, margin: int = 3):
    """configs[2]: the bifurcating Murray's-law tree of ``geometry.capsule_tree``."""
    from .geometry import tree_segments
    A, Bp, Rr, leaf, shape = tree_segments(generations, root_radius, root_length, seed, half_angle_deg, margin)
    AB = Bp - A
    L = np.sqrt((AB * AB).sum(1))
    d0 = AB[0] / L[0]
    iolets = [IoletPlane(2, 0, A[0] + d0 * 0.25, d0, Rr[0] + 2)]
    k_out = 0
    for k in range(A.shape[0]):
        if leaf[k]:
            dk = AB[k] / L[k]
            iolets.append(IoletPlane(3, k_out, Bp[k] - dk * 0.25, -dk, Rr[k] + 2))
            k_out += 1
    return capsule_array(A, Bp, Rr), iolets, shape


def weighted_decomposition_of_counts(counts: np.ndarray, boundary: np.ndarray, nranks: int, wall="BFL",
                                     architecture="B200", tolerance=0.03, initial="morton") -> np.ndarray:
    """``partition.weighted_kway`` on dense (bx,by,bz) arrays of fluid / boundary-typed sites per block:
    rank of every block in .gmy block order (-1 for empty blocks), for ``hlb_dom_set_partition_blocks``.
    ``initial``: "morton" | "rcb" | "inertial" | "best" (``partition.weighted_kway``)."""
    from .partition import REFERENCE_WEIGHTS, weighted_kway
    w = REFERENCE_WEIGHTS[architecture]
    ijk = np.argwhere(counts > 0)
    c, b = counts[counts > 0].astype(np.float64), boundary[counts > 0].astype(np.float64)
    loads = w["bulk"] * (c - b) + w[wall] * b
    part = weighted_kway(ijk, loads, nranks, tolerance, initial=initial)
    out = np.full(counts.shape, -1, np.int32)
    out[counts > 0] = part
    return out.ravel()


def basic_decomposition_of_counts(counts: np.ndarray, nranks: int) -> np.ndarray:
    """``BasicDecomposition`` on a dense (bx,by,bz) array of fluid sites per block (what
    ``DeviceDomain.count_block_sites`` returns): rank of every block in .gmy block order, -1 for
    empty blocks."""
    from .geometry import basic_decomposition_blocks
    bd = counts.shape
    flat = counts.reshape(-1)
    uniq = np.nonzero(flat)[0]
    ijk = np.stack([uniq // (bd[1] * bd[2]), (uniq // bd[2]) % bd[1], uniq % bd[2]], 1)
    out = np.full(flat.size, -1, np.int32)
    out[uniq] = basic_decomposition_blocks(ijk, flat[uniq], nranks)
    return out
