"""Geometry inputs for the collide-and-stream path: the ``.gmy`` site/link model, a reader/writer
for the file format, synthetic generators, and the reference's basic block decomposition.

This is host-side *input* handling (what ``geometry::GeometryReader`` hands to ``geometry::Domain``
in the reference); nothing here runs per time step.

Reference: ``doc/dev/file-formats/geometry.md`` (format), ``Code/io/formats/geometry.h:120-156``
(26-link order), ``Code/geometry/GeometryReader.cc:556-652`` (``ParseSite``),
``Code/geometry/decomposition/BasicDecomposition.cc:21-96``.
"""
from __future__ import annotations

import struct
import zlib
from dataclasses import dataclass, field

import numpy as np

# Code/io/formats/geometry.h:120-156 -- the 3D Moore neighbourhood in file order
NEIGHBOURHOOD = np.array(
    [(i, j, k) for i in (-1, 0, 1) for j in (-1, 0, 1) for k in (-1, 0, 1) if (i, j, k) != (0, 0, 0)],
    dtype=np.int32,
)
CUT_NONE, CUT_WALL, CUT_INLET, CUT_OUTLET = 0, 1, 2, 3
HLB_MAGIC = 0x686C6221  # Code/io/formats/formats.h
GMY_MAGIC = 0x676D7904  # Code/io/formats/geometry.h
GMY_VERSION = 4


@dataclass
class Geometry:
    """Fluid sites of a voxelised domain plus the cut-link records of its boundary sites.

    ``coords`` are global voxel coordinates of the fluid sites in ``.gmy`` order (blocks x-major /
    z-fastest, sites within a block likewise).  Only sites with at least one cut link carry a
    record (``bsite`` indexes ``coords``); link arrays are in the file's 26-neighbour order.
    """

    block_dims: np.ndarray  # (3,) blocks per axis
    block_size: int
    coords: np.ndarray  # (N,3) int32
    bsite: np.ndarray  # (Nb,) int64
    btype: np.ndarray  # (Nb,26) uint8
    biolet: np.ndarray  # (Nb,26) int32
    bdist: np.ndarray  # (Nb,26) float32
    bnavail: np.ndarray  # (Nb,) uint8
    bnormal: np.ndarray  # (Nb,3) float32
    meta: dict = field(default_factory=dict)

    @property
    def n_sites(self) -> int:
        return int(self.coords.shape[0])

    def gmy_sort(self) -> "Geometry":
        """Put sites into .gmy order (needed after a generator emitted them in another order)."""
        B = self.block_size
        c = self.coords.astype(np.int64)
        b = c // B
        s = c % B
        bidx = (b[:, 0] * self.block_dims[1] + b[:, 1]) * self.block_dims[2] + b[:, 2]
        sidx = (s[:, 0] * B + s[:, 1]) * B + s[:, 2]
        order = np.argsort(bidx * (B**3) + sidx, kind="stable")
        inv = np.empty_like(order)
        inv[order] = np.arange(order.size)
        self.coords = np.ascontiguousarray(self.coords[order])
        self.bsite = inv[self.bsite]
        o2 = np.argsort(self.bsite, kind="stable")
        for name in ("bsite", "btype", "biolet", "bdist", "bnavail", "bnormal"):
            setattr(self, name, np.ascontiguousarray(getattr(self, name)[o2]))
        return self


# ---------------------------------------------------------------------------------------------
# .gmy reader / writer
# ---------------------------------------------------------------------------------------------
def write_gmy(geom: Geometry, path: str) -> None:
    """Write ``geom`` in HemeLB's geometry format (XDR, zlib per block)."""
    B = geom.block_size
    bd = [int(x) for x in geom.block_dims]
    nblocks = bd[0] * bd[1] * bd[2]
    c = geom.coords.astype(np.int64)
    bidx = ((c[:, 0] // B) * bd[1] + (c[:, 1] // B)) * bd[2] + (c[:, 2] // B)
    sidx = ((c[:, 0] % B) * B + (c[:, 1] % B)) * B + (c[:, 2] % B)
    if np.any(np.diff(bidx * B**3 + sidx) <= 0):
        raise ValueError("geometry is not in .gmy order; call gmy_sort() first")
    brec = np.full(geom.n_sites, -1, dtype=np.int64)
    brec[geom.bsite] = np.arange(geom.bsite.size)
    starts = np.searchsorted(bidx, np.arange(nblocks + 1))
    header = []
    payload = []
    for b in range(nblocks):
        lo, hi = int(starts[b]), int(starts[b + 1])
        if lo == hi:
            header.append((0, 0, 0))
            continue
        fluid_at = {int(sidx[i]): i for i in range(lo, hi)}
        out = bytearray()
        for s in range(B**3):
            i = fluid_at.get(s)
            if i is None:
                out += struct.pack(">I", 0)
                continue
            out += struct.pack(">I", 1)
            r = int(brec[i])
            for l in range(26):
                t = int(geom.btype[r, l]) if r >= 0 else 0
                out += struct.pack(">I", t)
                if t == CUT_WALL:
                    out += struct.pack(">f", float(geom.bdist[r, l]))
                elif t != CUT_NONE:
                    out += struct.pack(">If", int(geom.biolet[r, l]), float(geom.bdist[r, l]))
            if r >= 0 and geom.bnavail[r]:
                out += struct.pack(">I", 1) + struct.pack(">fff", *[float(x) for x in geom.bnormal[r]])
            else:
                out += struct.pack(">I", 0)
        comp = zlib.compress(bytes(out))
        header.append((hi - lo, len(comp), len(out)))
        payload.append(comp)
    with open(path, "wb") as fh:
        fh.write(struct.pack(">IIIIIIII", HLB_MAGIC, GMY_MAGIC, GMY_VERSION, bd[0], bd[1], bd[2], B, 0))
        for h in header:
            fh.write(struct.pack(">III", *h))
        for p in payload:
            fh.write(p)


def read_gmy(path: str) -> Geometry:
    """Parse a ``.gmy`` file (all 26 links kept; the lattice picks its subset later)."""
    with open(path, "rb") as fh:
        raw = fh.read()
    hlb, gmy, version, bx, by, bz, B, pad = struct.unpack_from(">IIIIIIII", raw, 0)
    if hlb != HLB_MAGIC or gmy != GMY_MAGIC:
        raise ValueError("not a HemeLB geometry file")
    nblocks = bx * by * bz
    off = 32
    header = [struct.unpack_from(">III", raw, off + 12 * b) for b in range(nblocks)]
    off += 12 * nblocks
    coords, bsite, btype, biolet, bdist, bnavail, bnormal = [], [], [], [], [], [], []
    for b, (nfluid, clen, ulen) in enumerate(header):
        if nfluid == 0:
            continue
        data = zlib.decompress(raw[off : off + clen])
        off += clen
        if len(data) != ulen:
            raise ValueError("block %d: uncompressed length mismatch" % b)
        bi, bj, bk = b // (by * bz), (b // bz) % by, b % bz
        p = 0
        for s in range(B**3):
            (isfluid,) = struct.unpack_from(">I", data, p)
            p += 4
            if not isfluid:
                continue
            si, sj, sk = s // (B * B), (s // B) % B, s % B
            types = np.zeros(26, np.uint8)
            ids = np.full(26, -1, np.int32)
            dists = np.full(26, -1.0, np.float32)
            for l in range(26):
                (t,) = struct.unpack_from(">I", data, p)
                p += 4
                types[l] = t
                if t == CUT_WALL:
                    (dists[l],) = struct.unpack_from(">f", data, p)
                    p += 4
                elif t != CUT_NONE:
                    ids[l], dists[l] = struct.unpack_from(">If", data, p)
                    p += 8
            (navail,) = struct.unpack_from(">I", data, p)
            p += 4
            normal = (0.0, 0.0, 0.0)
            if navail:
                normal = struct.unpack_from(">fff", data, p)
                p += 12
            if types.any() or navail:
                bsite.append(len(coords))
                btype.append(types)
                biolet.append(ids)
                bdist.append(dists)
                bnavail.append(navail)
                bnormal.append(normal)
            coords.append((bi * B + si, bj * B + sj, bk * B + sk))
    nb = len(bsite)
    return Geometry(
        block_dims=np.array([bx, by, bz], np.int32),
        block_size=int(B),
        coords=np.array(coords, np.int32).reshape(-1, 3),
        bsite=np.array(bsite, np.int64),
        btype=np.array(btype, np.uint8).reshape(nb, 26),
        biolet=np.array(biolet, np.int32).reshape(nb, 26),
        bdist=np.array(bdist, np.float32).reshape(nb, 26),
        bnavail=np.array(bnavail, np.uint8),
        bnormal=np.array(bnormal, np.float32).reshape(nb, 3),
        meta={"source": path},
    )


# ---------------------------------------------------------------------------------------------
# implicit-surface voxeliser and the synthetic geometries of BASELINE.json
# ---------------------------------------------------------------------------------------------
@dataclass
class IoletPlane:
    """A flat inlet/outlet cap: sites with (p - position).normal <= 0 within ``radius`` of the
    centre are outside the domain; ``normal`` points into the fluid."""

    kind: int  # CUT_INLET or CUT_OUTLET
    index: int  # index into the inlet (or outlet) array of the simulation config
    position: np.ndarray
    normal: np.ndarray
    radius: float


def voxelise(shape, phi, iolets, block_size=8, normal_fn=None, chunk=None, mask=None, phi_local=None, _site_hook=None) -> Geometry:
    """Voxelise the region ``phi(p) < 0`` clipped by the iolet planes.

    ``phi`` maps an (M,3) float64 array of positions to signed distance-like values (negative =
    fluid).  Cut distances are the fraction of the lattice vector to the first crossing, found by
    bisection on ``phi`` (walls) or analytically (planes), then rounded to float32 as a ``.gmy``
    file stores them.  ``normal_fn`` gives wall normals at positions (default: normalised
    finite-difference gradient of ``phi``).
    """
    shape = np.asarray(shape, np.int64)
    B = block_size
    bdims = (shape + B - 1) // B

    def clipped(p):
        out = np.zeros(p.shape[0], bool)
        which = np.full(p.shape[0], -1, np.int32)
        for k, io in enumerate(iolets):
            d = p - io.position
            h = d @ io.normal
            r2 = (d * d).sum(1) - h * h
            m = (h <= 0) & (r2 <= (io.radius * 1.5) ** 2) & (h > -4.0 - io.radius)
            which[m & ~out] = k
            out |= m
        return out, which

    def is_fluid(p):
        c, _ = clipped(p)
        return (phi(p) < 0) & ~c

    # fluid mask, slab by slab to bound memory (or a precomputed candidate mask refined by the clips)
    coords = []
    nx = int(shape[0])
    if mask is not None:
        c = np.argwhere(mask).astype(np.float64)
        cl, _ = clipped(c)
        coords.append(c[~cl].astype(np.int32))
        nx = 0
    step = chunk or max(1, int(4e6 // max(1, int(shape[1] * shape[2]))))
    yy, zz = np.meshgrid(np.arange(shape[1]), np.arange(shape[2]), indexing="ij")
    for x0 in range(0, nx, step):
        xs = np.arange(x0, min(nx, x0 + step))
        p = np.stack(
            [np.repeat(xs, yy.size), np.tile(yy.ravel(), xs.size), np.tile(zz.ravel(), xs.size)], 1
        ).astype(np.float64)
        m = is_fluid(p)
        coords.append(p[m].astype(np.int32))
    coords = np.concatenate(coords, 0)
    N = coords.shape[0]
    grid = np.zeros(tuple(int(s) + 2 for s in shape), bool)  # 1-voxel solid rim
    grid[coords[:, 0] + 1, coords[:, 1] + 1, coords[:, 2] + 1] = True

    # boundary sites = any of the 26 neighbours is not fluid
    nb_fluid = np.empty((N, 26), bool)
    for l, c in enumerate(NEIGHBOURHOOD):
        nb_fluid[:, l] = grid[coords[:, 0] + 1 + c[0], coords[:, 1] + 1 + c[1], coords[:, 2] + 1 + c[2]]
    bsite = np.nonzero(~nb_fluid.all(1))[0].astype(np.int64)
    Nb = bsite.size
    btype = np.zeros((Nb, 26), np.uint8)
    biolet = np.full((Nb, 26), -1, np.int32)
    bdist = np.full((Nb, 26), -1.0, np.float32)
    p0 = coords[bsite].astype(np.float64)
    if _site_hook is not None:
        _site_hook["sites"] = p0
    if phi_local is None:
        def phi_local(p, which):
            return phi(p)
    for l, c in enumerate(NEIGHBOURHOOD):
        cut = ~nb_fluid[bsite, l]
        if not cut.any():
            continue
        idx = np.nonzero(cut)[0]
        a = p0[idx]
        cvec = c.astype(np.float64)
        b = a + cvec
        # wall crossing by bisection (phi(a) < 0 always)
        t_wall = np.full(idx.size, np.inf)
        outside = phi_local(b, idx) >= 0
        if outside.any():
            lo = np.zeros(outside.sum())
            hi = np.ones(outside.sum())
            aa = a[outside]
            io = idx[outside]
            for _ in range(30):
                mid = 0.5 * (lo + hi)
                inside = phi_local(aa + mid[:, None] * cvec, io) < 0
                lo = np.where(inside, mid, lo)
                hi = np.where(inside, hi, mid)
            t_wall[outside] = hi
        # plane crossing
        t_io = np.full(idx.size, np.inf)
        k_io = np.full(idx.size, -1, np.int32)
        cl, which = clipped(b)
        for k, io in enumerate(iolets):
            m = cl & (which == k)
            if not m.any():
                continue
            h0 = (a[m] - io.position) @ io.normal
            dh = cvec @ io.normal
            t = h0 / (-dh)
            t_io[m] = t
            k_io[m] = k
        is_io = t_io < t_wall
        t = np.where(is_io, t_io, t_wall)
        kinds = np.array([io.kind for io in iolets] + [CUT_WALL], np.uint8)
        ids = np.array([io.index for io in iolets] + [-1], np.int32)
        btype[idx, l] = np.where(is_io, kinds[k_io], CUT_WALL)
        biolet[idx, l] = np.where(is_io, ids[k_io], -1)
        bdist[idx, l] = np.clip(t, 1e-6, 1.0).astype(np.float32)
    bnavail = (btype == CUT_WALL).any(1).astype(np.uint8)
    if normal_fn is None:
        def normal_fn(p):
            g = np.empty_like(p)
            ar = np.arange(p.shape[0])
            for k in range(3):
                e = np.zeros(3)
                e[k] = 0.25
                g[:, k] = phi_local(p + e, ar) - phi_local(p - e, ar)
            n = np.linalg.norm(g, axis=1)
            n[n == 0] = 1.0
            return g / n[:, None]
    bnormal = np.zeros((Nb, 3), np.float32)
    if Nb:
        nn = normal_fn(p0).astype(np.float32)
        bnormal[bnavail.astype(bool)] = nn[bnavail.astype(bool)]
    g = Geometry(bdims.astype(np.int32), B, coords, bsite, btype, biolet, bdist, bnavail, bnormal)
    return g.gmy_sort()


def cylinder(radius: float, length: int, block_size: int = 8, margin: int = 2) -> Geometry:
    """configs[1]: a straight cylinder along z, inlet cap at z-min, outlet cap at z-max."""
    R = float(radius)
    n = int(np.ceil(2 * R)) + 2 * margin + 1
    cx = cy = (n - 1) / 2.0
    shape = (n, n, length + 2 * margin)
    z0, z1 = margin, margin + length - 1

    def phi(p):
        return np.hypot(p[:, 0] - cx, p[:, 1] - cy) - R

    def normal_fn(p):
        d = np.stack([p[:, 0] - cx, p[:, 1] - cy, np.zeros(p.shape[0])], 1)
        r = np.linalg.norm(d, axis=1)
        r[r == 0] = 1.0
        return d / r[:, None]

    iolets = [
        IoletPlane(CUT_INLET, 0, np.array([cx, cy, z0 - 0.5]), np.array([0.0, 0.0, 1.0]), R + 2),
        IoletPlane(CUT_OUTLET, 0, np.array([cx, cy, z1 + 0.5]), np.array([0.0, 0.0, -1.0]), R + 2),
    ]
    g = voxelise(shape, phi, iolets, block_size, normal_fn)
    g.meta.update(kind="cylinder", radius=R, length=length, axis=(cx, cy), z0=z0, z1=z1,
                  inlets=[iolets[0]], outlets=[iolets[1]])
    return g


def cylinder_slab(radius: float, length: int, nranks: int, rank: int, block_size: int = 8, margin: int = 2):
    """Rank ``rank``'s share of ``cylinder(radius, length)`` cut into ``nranks`` equal z-slabs, plus one
    halo slice either side: (sub-geometry in global coordinates, site -> rank).  Enough to build
    this rank's Domain tables exactly (the pair lists only involve the adjacent slices)."""
    per = length // nranks
    lo = rank * per
    hi = length if rank == nranks - 1 else (rank + 1) * per
    a, b = max(0, lo - 1), min(length, hi + 1)
    g = cylinder_extruded(radius, length, block_size, margin, z_range=(a, b))
    zi = g.coords[:, 2].astype(np.int64) - margin
    ranks = np.minimum(zi // per, nranks - 1).astype(np.int32)
    return g, ranks


def cylinder_extruded(radius: float, length: int, block_size: int = 8, margin: int = 2, z_range=None) -> Geometry:
    """The same geometry as ``cylinder`` built by extruding the three distinct z-slices (inlet cap,
    interior, outlet cap) of a short template -- O(N) and fast enough for the 1e8-site benchmark."""
    if length < 3:
        return cylinder(radius, length, block_size, margin)
    t = cylinder(radius, 3, block_size, margin)
    z0 = margin
    tz = t.coords[:, 2]
    brec = np.full(t.n_sites, -1, np.int64)
    brec[t.bsite] = np.arange(t.bsite.size)
    parts_c, parts_b = [], []
    sl = {k: np.nonzero(tz == z0 + k)[0] for k in range(3)}
    xy_mid = t.coords[sl[1]][:, :2]
    za, zb = (0, length) if z_range is None else z_range  # slice indices kept, [za, zb)
    m_lo, m_hi = max(za, 1), min(zb, length - 1)
    nmid = max(0, m_hi - m_lo)
    # interior slices replicate the template's middle slice
    zs = np.arange(z0 + m_lo, z0 + m_lo + nmid, dtype=np.int32)
    cm = np.empty((nmid, xy_mid.shape[0], 3), np.int32)
    cm[:, :, :2] = xy_mid[None]
    cm[:, :, 2] = zs[:, None]
    c0 = t.coords[sl[0]].copy()
    c2 = t.coords[sl[2]].copy()
    c2[:, 2] = z0 + length - 1
    if za > 0:
        c0 = c0[:0]
    if zb < length:
        c2 = c2[:0]
    coords = np.concatenate([c0, cm.reshape(-1, 3), c2], 0)
    # boundary records
    b0 = brec[sl[0]] if za == 0 else brec[sl[0]][:0]
    b1 = brec[sl[1]]
    b2 = brec[sl[2]] if zb == length else brec[sl[2]][:0]
    n0, n1 = c0.shape[0], sl[1].size
    has1 = np.nonzero(b1 >= 0)[0]
    bsite = np.concatenate([
        np.nonzero(b0 >= 0)[0],
        (n0 + (np.arange(nmid, dtype=np.int64)[:, None] * n1 + has1[None, :])).ravel(),
        n0 + nmid * n1 + np.nonzero(b2 >= 0)[0]]).astype(np.int64)
    rec = np.concatenate([b0[b0 >= 0], np.tile(b1[has1], nmid), b2[b2 >= 0]])
    shape_z = length + 2 * margin
    bdims = t.block_dims.copy()
    bdims[2] = (shape_z + block_size - 1) // block_size
    g = Geometry(bdims, block_size, coords, bsite, t.btype[rec], t.biolet[rec], t.bdist[rec], t.bnavail[rec],
                 t.bnormal[rec])
    inl = t.meta["inlets"][0]
    outl = t.meta["outlets"][0]
    outl = IoletPlane(outl.kind, outl.index, np.array([outl.position[0], outl.position[1], z0 + length - 0.5]),
                      outl.normal, outl.radius)
    g.meta.update(kind="cylinder", radius=float(radius), length=length, axis=t.meta["axis"], z0=z0,
                  z1=z0 + length - 1, inlets=[inl], outlets=[outl])
    return g.gmy_sort()


def tree_segments(generations: int, root_radius: float, root_length: float, seed: int = 20261017,
                  half_angle_deg: float = 35.0, margin: int = 3):
    """Capsule end points / radii of the bifurcating tree (Murray's law, r_child = r / 2^(1/3)),
    shifted into a lattice with ``margin`` voxels of clearance: (A, B, R, is_leaf, voxel shape)."""
    rng = np.random.default_rng(seed)
    segs = []  # (a, b, r, is_leaf)

    def grow(a, direction, r, length, gen, ref):
        b = a + direction * length
        leaf = gen == generations - 1
        segs.append((a, b, r, leaf))
        if leaf:
            return
        # branch plane: rotate a reference perpendicular by a seeded angle
        perp = np.cross(direction, ref)
        if np.linalg.norm(perp) < 1e-6:
            perp = np.cross(direction, np.array([1.0, 0.0, 0.0]))
        perp /= np.linalg.norm(perp)
        ang = rng.uniform(0, np.pi)
        perp = perp * np.cos(ang) + np.cross(direction, perp) * np.sin(ang)
        th = np.deg2rad(half_angle_deg)
        for sgn in (+1, -1):
            d2 = direction * np.cos(th) + sgn * perp * np.sin(th)
            grow(b, d2 / np.linalg.norm(d2), r * 2 ** (-1.0 / 3.0), length * 0.8, gen + 1, perp)

    grow(np.zeros(3), np.array([0.0, 0.0, 1.0]), root_radius, root_length, 0, np.array([0.0, 1.0, 0.0]))
    pts = np.array([s[0] for s in segs] + [s[1] for s in segs])
    rmax = root_radius
    lo = pts.min(0) - rmax - margin
    hi = pts.max(0) + rmax + margin
    shift = -lo
    shape = np.ceil(hi - lo).astype(np.int64) + 1
    A = np.array([s[0] + shift for s in segs])
    Bp = np.array([s[1] + shift for s in segs])
    Rr = np.array([s[2] for s in segs])
    leaf = np.array([s[3] for s in segs], bool)
    return A, Bp, Rr, leaf, shape


def capsule_tree(generations: int, root_radius: float, root_length: float, seed: int = 20261017,
                 half_angle_deg: float = 35.0, block_size: int = 8, margin: int = 3):
    """configs[2]: a bifurcating tree of cylinders obeying Murray's law (r_child = r / 2^(1/3)),
    one inlet at the root and one outlet per leaf branch."""
    A, Bp, Rr, leaf, shape = tree_segments(generations, root_radius, root_length, seed, half_angle_deg, margin)
    segs = [(A[k], Bp[k], Rr[k], bool(leaf[k])) for k in range(A.shape[0])]
    AB = Bp - A
    L2 = (AB * AB).sum(1)

    def seg_dist(p, k):
        d = p - A[k]
        t = np.clip((d @ AB[k]) / L2[k], 0.0, 1.0)
        return np.linalg.norm(d - t[:, None] * AB[k], axis=1) - Rr[k]

    def phi(p):
        out = np.full(p.shape[0], np.inf)
        for k in range(A.shape[0]):
            np.minimum(out, seg_dist(p, k), out=out)
        return out

    # rasterise each capsule inside its own bounding box
    mask = np.zeros(tuple(int(x) for x in shape), bool)
    boxes = []
    for k in range(A.shape[0]):
        lo_k = np.maximum(np.floor(np.minimum(A[k], Bp[k]) - Rr[k] - 2).astype(int), 0)
        hi_k = np.minimum(np.ceil(np.maximum(A[k], Bp[k]) + Rr[k] + 2).astype(int) + 1, shape)
        boxes.append((lo_k, hi_k))
        gx, gy, gz = np.meshgrid(np.arange(lo_k[0], hi_k[0]), np.arange(lo_k[1], hi_k[1]), np.arange(lo_k[2], hi_k[2]),
                                 indexing="ij")
        pts = np.stack([gx.ravel(), gy.ravel(), gz.ravel()], 1).astype(np.float64)
        inside = seg_dist(pts, k) < 0
        mask[lo_k[0]:hi_k[0], lo_k[1]:hi_k[1], lo_k[2]:hi_k[2]] |= inside.reshape(gx.shape)

    cand_cache = {}

    def phi_local(p, which):
        # min over the capsules whose box holds the boundary site (set up lazily per call pattern)
        out = np.full(p.shape[0], np.inf)
        site = cand_cache["sites"][which]
        for k, (lo_k, hi_k) in enumerate(boxes):
            m = ((site >= lo_k - 1) & (site < hi_k + 1)).all(1)
            if m.any():
                out[m] = np.minimum(out[m], seg_dist(p[m], k))
        return out

    inlets, outlets = [], []
    d0 = AB[0] / np.sqrt(L2[0])
    inlets.append(IoletPlane(CUT_INLET, 0, A[0] + d0 * 0.25, d0, Rr[0] + 2))
    for k, s in enumerate(segs):
        if s[3]:
            dk = AB[k] / np.sqrt(L2[k])
            outlets.append(IoletPlane(CUT_OUTLET, len(outlets), Bp[k] - dk * 0.25, -dk, Rr[k] + 2))
    g = voxelise(shape, phi, inlets + outlets, block_size, mask=mask, phi_local=phi_local, _site_hook=cand_cache)
    g.meta.update(kind="tree", generations=generations, inlets=inlets, outlets=outlets, segments=len(segs))
    return g


def sac(radius: float, neck_radius: float, neck_length: float, roughness: float = 0.0,
        seed: int = 20261017, block_size: int = 8, margin: int = 3) -> Geometry:
    """configs[4]: an aneurysm-like sphere with two opposite necks along z and seeded value-noise
    wall roughness (high wall-site fraction)."""
    rng = np.random.default_rng(seed)
    R = float(radius)
    n = int(np.ceil(2 * (R + roughness))) + 2 * margin + 1
    c = np.array([(n - 1) / 2.0] * 2 + [0.0])
    zlen = int(np.ceil(2 * R + 2 * neck_length)) + 2 * margin
    c[2] = (zlen - 1) / 2.0
    shape = (n, n, zlen)
    ng = 8
    noise = rng.uniform(-1.0, 1.0, (ng, ng, ng))

    def value_noise(p):
        q = (p / np.array(shape, float)) * (ng - 1)
        i0 = np.clip(np.floor(q).astype(int), 0, ng - 2)
        f = q - i0
        acc = np.zeros(p.shape[0])
        for dx in (0, 1):
            for dy in (0, 1):
                for dz in (0, 1):
                    w = (f[:, 0] if dx else 1 - f[:, 0]) * (f[:, 1] if dy else 1 - f[:, 1]) * (
                        f[:, 2] if dz else 1 - f[:, 2])
                    acc += w * noise[i0[:, 0] + dx, i0[:, 1] + dy, i0[:, 2] + dz]
        return acc

    def phi(p):
        d = p - c
        sphere = np.linalg.norm(d, axis=1) - R
        if roughness:
            sphere = sphere - roughness * value_noise(p)
        neck = np.hypot(d[:, 0], d[:, 1]) - neck_radius
        return np.minimum(sphere, neck)

    zin = margin - 0.5
    zout = zlen - margin - 0.5
    inlets = [IoletPlane(CUT_INLET, 0, np.array([c[0], c[1], zin]), np.array([0, 0, 1.0]), neck_radius + 2)]
    outlets = [IoletPlane(CUT_OUTLET, 0, np.array([c[0], c[1], zout]), np.array([0, 0, -1.0]), neck_radius + 2)]
    g = voxelise(shape, phi, inlets + outlets, block_size)
    g.meta.update(kind="sac", inlets=inlets, outlets=outlets)
    return g


def four_cube() -> Geometry:
    """configs[0]: the reference's ``four_cube.gmy`` test geometry re-created from its recipe
    (``Scripts/SimpleGeometryGenerationScripts/four_cube.py``, ``tests/helpers/
    FourCubeLatticeData.cc:53-170``): a 4^3 fluid cube centred in one 6^3 block, inlet below
    z-min, outlet above z-max, walls on the x/y faces, every cut distance 0.5."""
    B = 6
    coords, bsite, btype, biolet, bdist, bnavail, bnormal = [], [], [], [], [], [], []
    for i in range(1, 5):
        for j in range(1, 5):
            for k in range(1, 5):
                types = np.zeros(26, np.uint8)
                ids = np.full(26, -1, np.int32)
                dists = np.full(26, -1.0, np.float32)
                normal = np.zeros(3, np.float32)
                for l, c in enumerate(NEIGHBOURHOOD):
                    ni, nj, nk = i + c[0], j + c[1], k + c[2]
                    if 1 <= ni <= 4 and 1 <= nj <= 4 and 1 <= nk <= 4:
                        continue
                    # links crossing both a wall face and an iolet face count as iolet
                    if nk < 1:
                        types[l], ids[l] = CUT_INLET, 0
                    elif nk > 4:
                        types[l], ids[l] = CUT_OUTLET, 0
                    else:
                        types[l] = CUT_WALL
                    dists[l] = 0.5
                if i == 1:
                    normal[:] = (-1, 0, 0)
                if i == 4:
                    normal[:] = (1, 0, 0)
                if j == 1:
                    normal[:] = (0, -1, 0)
                if j == 4:
                    normal[:] = (0, 1, 0)
                if types.any():
                    bsite.append(len(coords))
                    btype.append(types)
                    biolet.append(ids)
                    bdist.append(dists)
                    iswall = (types == CUT_WALL).any()
                    bnavail.append(1 if iswall else 0)
                    bnormal.append(normal if iswall else np.zeros(3, np.float32))
                coords.append((i, j, k))
    nb = len(bsite)
    return Geometry(np.array([1, 1, 1], np.int32), B, np.array(coords, np.int32), np.array(bsite, np.int64),
                    np.array(btype, np.uint8).reshape(nb, 26), np.array(biolet, np.int32).reshape(nb, 26),
                    np.array(bdist, np.float32).reshape(nb, 26), np.array(bnavail, np.uint8),
                    np.array(bnormal, np.float32).reshape(nb, 3), meta={"kind": "four_cube"})


# ---------------------------------------------------------------------------------------------
# decomposition
# ---------------------------------------------------------------------------------------------
def _spread(v: np.ndarray) -> np.ndarray:
    v = v.astype(np.uint64)
    r = np.zeros_like(v)
    for b in range(21):
        r |= ((v >> np.uint64(b)) & np.uint64(1)) << np.uint64(3 * b)
    return r


def morton(ijk: np.ndarray) -> np.ndarray:
    """Octree id of block coordinates, x most significant (``LookupTree.h:92-97``)."""
    return (_spread(ijk[:, 0]) << np.uint64(2)) ^ (_spread(ijk[:, 1]) << np.uint64(1)) ^ _spread(ijk[:, 2])


def basic_decomposition_blocks(ijk: np.ndarray, counts: np.ndarray, nranks: int) -> np.ndarray:
    """Rank of each non-empty block by the reference's ``BasicDecomposition``: recursive bisection
    of the Morton-ordered cumulative fluid-site counts (``BasicDecomposition.cc:21-96``)."""
    order = np.argsort(morton(ijk), kind="stable")
    counts_m = counts[order]
    if counts_m.size < nranks:
        raise ValueError("More ranks than blocks")
    cum = np.concatenate([[0], np.cumsum(counts_m.astype(np.int64))]).astype(np.uint64)
    rank_for_block = np.zeros(counts_m.size, np.int32)

    def assign(cb, ce, rb, re, n):
        if n < 2:
            return
        lo, hi = cum[cb], cum[ce]
        delta = hi - lo
        n_lo = n // 2
        mid = np.float32(lo) + np.float32(delta) * np.float32(n_lo) / np.float32(n)
        m = cb + int(np.searchsorted(cum[cb:ce].astype(np.float32), mid, side="left"))
        if (np.float32(cum[m]) - mid) / np.float32(delta) > np.float32(0.5):
            m -= 1
        dn = m - cb
        rm = rb + dn
        assign(cb, m, rb, rm, n_lo)
        assign(m, ce, rm, re, n - n_lo)
        rank_for_block[rm:re] += n_lo

    assign(0, cum.size - 1, 0, rank_for_block.size, nranks)
    block_rank = np.empty(counts_m.size, np.int32)
    block_rank[order] = rank_for_block
    return block_rank


def basic_decomposition(geom: Geometry, nranks: int) -> np.ndarray:
    """Site -> rank by the reference's ``BasicDecomposition`` (whole blocks)."""
    B = geom.block_size
    bc = (geom.coords // B).astype(np.int64)
    bd = geom.block_dims.astype(np.int64)
    gmy_idx = (bc[:, 0] * bd[1] + bc[:, 1]) * bd[2] + bc[:, 2]
    uniq, counts = np.unique(gmy_idx, return_counts=True)
    ijk = np.stack([uniq // (bd[1] * bd[2]), (uniq // bd[2]) % bd[1], uniq % bd[2]], 1)
    block_rank = basic_decomposition_blocks(ijk, counts, nranks)
    return block_rank[np.searchsorted(uniq, gmy_idx)].astype(np.int32)


def slab_decomposition(geom: Geometry, nranks: int, axis: int = 2) -> np.ndarray:
    """Site-level equal-count slabs along ``axis`` (a stand-in for a ParMETIS site partition: cuts
    through blocks, so ranks share blocks as they do after the reference's optimisation step)."""
    x = geom.coords[:, axis].astype(np.int64)
    order = np.argsort(x, kind="stable")
    rank = np.empty(geom.n_sites, np.int32)
    rank[order] = (np.arange(geom.n_sites) * nranks // geom.n_sites).astype(np.int32)
    return rank
