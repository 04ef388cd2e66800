"""Per-rank ``geometry::Domain`` tables from a Geometry and a site -> rank map.

A vectorised (numpy) construction of exactly the tables the reference builds in
``Code/geometry/Domain.cc:69-580``: the local site order (Morton-ordered blocks, z-fastest sites,
stably bucketed mid-domain[type 0..5] then domain-edge[type 0..5]), ``neighbourIndices``, the
``neighbouringProcs`` list, the halo send slots and ``streamingIndicesForReceivedDistributions``.
The tests compare every table bit-for-bit with the oracle's literal restatement.

Large domains never materialise the N*Q int64 ``neighbourIndices`` on the host: it is produced in
site chunks (``RankDomain.neighbour_indices(first, n)``) and streamed to the device.
"""
from __future__ import annotations

import numpy as np

from .geometry import NEIGHBOURHOOD, CUT_WALL, CUT_INLET, CUT_OUTLET, Geometry, morton

C27 = np.array(
    [(0, 0, 0), (1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1), (1, 1, 0), (-1, -1, 0),
     (1, -1, 0), (-1, 1, 0), (1, 0, 1), (-1, 0, -1), (1, 0, -1), (-1, 0, 1), (0, 1, 1), (0, -1, -1), (0, 1, -1),
     (0, -1, 1), (1, 1, 1), (-1, -1, -1), (1, 1, -1), (-1, -1, 1), (1, -1, 1), (-1, 1, -1), (1, -1, -1),
     (-1, 1, 1)], np.int64)
C15 = np.array(
    [(0, 0, 0), (1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1), (1, 1, 1), (-1, -1, -1),
     (1, 1, -1), (-1, -1, 1), (1, -1, 1), (-1, 1, -1), (1, -1, -1), (-1, 1, 1)], np.int64)


def lattice_vectors(Q: int) -> np.ndarray:
    """Velocity set (D3Q15.h:17-35, D3Q19.h:18-36, D3Q27.h:18-44)."""
    if Q == 15:
        return C15
    if Q in (19, 27):
        return C27[:Q]
    raise ValueError("lattice must be D3Q15, D3Q19 or D3Q27")


def inverse_directions(Q: int) -> np.ndarray:
    inv = np.arange(Q)
    inv[1::2] += 1
    inv[2::2] -= 1
    return inv


def gmy_link_of_direction(Q: int) -> np.ndarray:
    """Index into the .gmy 26-neighbourhood of each lattice direction 1..Q-1
    (GeometryReader.cc:614-626)."""
    c = lattice_vectors(Q)
    out = np.zeros(Q, np.int64)
    for d in range(1, Q):
        out[d] = int(np.nonzero((NEIGHBOURHOOD == c[d]).all(1))[0][0])
    return out


class _Lookup:
    """global voxel coordinates -> input site index (or -1), over the sites' bounding box padded by
    one voxel (so every lattice neighbour of a site falls inside the window)."""

    def __init__(self, geom: Geometry):
        c = geom.coords.astype(np.int64)
        if geom.n_sites:
            self.origin = c.min(0) - 1
            self.dims = c.max(0) - c.min(0) + 3
        else:
            self.origin = np.zeros(3, np.int64)
            self.dims = np.ones(3, np.int64)
        self.full = geom.block_dims.astype(np.int64) * geom.block_size
        nvox = int(self.dims.prod())
        self.dense = None
        q = c - self.origin
        self.key_of_site = (q[:, 0] * self.dims[1] + q[:, 1]) * self.dims[2] + q[:, 2]
        self.rimmed = True
        if nvox <= 3_000_000_000 and nvox <= 64 * max(geom.n_sites, 1) + 10_000_000:
            dt = np.int32 if geom.n_sites < 2**31 else np.int64
            self.dense = np.full(nvox, -1, dt)
            self.dense[self.key_of_site] = np.arange(geom.n_sites, dtype=dt)
        else:
            self.order = np.argsort(self.key_of_site, kind="stable")
            self.keys = self.key_of_site[self.order]

    def key_offset(self, c) -> int:
        return int((c[0] * self.dims[1] + c[1]) * self.dims[2] + c[2])

    def __call__(self, p: np.ndarray) -> np.ndarray:
        q = p - self.origin
        inside = ((q >= 0) & (q < self.dims)).all(1)
        q = np.where(inside[:, None], q, 0)
        key = (q[:, 0] * self.dims[1] + q[:, 1]) * self.dims[2] + q[:, 2]
        if self.dense is not None:
            out = self.dense[key].astype(np.int64)
        else:
            pos = np.searchsorted(self.keys, key)
            pos = np.minimum(pos, self.keys.size - 1)
            out = np.where(self.keys[pos] == key, self.order[pos], -1)
        return np.where(inside, out, -1)


class RankDomain:
    """The slice of one rank's ``geometry::Domain`` that the hot path reads."""

    def __init__(self):
        self.Q = 0
        self.rank = 0
        self.nranks = 1
        self.N = 0
        self.mid = np.zeros(6, np.int64)
        self.edge = np.zeros(6, np.int64)
        self.inputIndex = None  # (N,) local site -> input site
        self.globalCoords = None  # (N,3) int64
        self.totalSharedFs = 0
        self.procs = np.zeros((0, 3), np.int64)  # rank, SharedDistributionCount, FirstSharedDistribution
        self.streamingIndices = np.zeros(0, np.int64)
        # boundary-typed sites only (local ids boundary_first ranges), reference-form values
        self.wallMask = None
        self.ioletMask = None
        self.siteType = None
        self.ioletId = None
        self._b_local = None  # local ids that carry a boundary record
        self._b_dist = None  # (nb, Q-1) float32
        self._b_normal = None  # (nb,3) float32, inf where unavailable
        self._send_key = np.zeros(0, np.int64)  # local*Q + d, sorted
        self._send_slot = np.zeros(0, np.int64)
        self._builder = None

    # ---- chunked reference-form tables -------------------------------------------------------
    def neighbour_indices(self, first: int = 0, n: int | None = None) -> np.ndarray:
        """``neighbourIndices[first*Q : (first+n)*Q]`` (Domain.cc:425-505, 548-580)."""
        n = self.N - first if n is None else n
        b = self._builder
        Q = self.Q
        sites = self.inputIndex[first:first + n]
        out = np.empty((n, Q), np.int64)
        out[:, 0] = (np.arange(first, first + n, dtype=np.int64)) * Q
        if b.lookup.dense is not None:
            # one gather per direction from a grid of this rank's local ids (-1 solid, -2 a site
            # of another rank); only the few cross-rank links need the send-slot search
            key0 = b.lookup.key_of_site[sites]
            grid = b.local_grid(self.rank)
            ids = np.arange(first, first + n, dtype=np.int64)
            for d in range(1, Q):
                v = grid[key0 + b.lookup.key_offset(b.c[d])].astype(np.int64)
                col = np.where(v >= 0, v * Q + d, self.N * Q)
                rem = np.nonzero(v == -2)[0]
                if rem.size:
                    col[rem] = self._send_slot[np.searchsorted(self._send_key, ids[rem] * Q + d)]
                out[:, d] = col
            return out.reshape(-1)
        base = b.geom.coords[sites].astype(np.int64)
        for d in range(1, Q):
            nb = b.lookup(base + b.c[d])
            col = np.full(n, self.N * Q, np.int64)  # rubbish site
            fluid = nb >= 0
            local = fluid & (b.rank_of_site[np.maximum(nb, 0)] == self.rank)
            col[local] = b.local_of_input[nb[local]] * Q + d
            remote = fluid & ~local
            if remote.any():
                key = (np.arange(first, first + n, dtype=np.int64)[remote]) * Q + d
                pos = np.searchsorted(self._send_key, key)
                col[remote] = self._send_slot[pos]
            out[:, d] = col
        return out.reshape(-1)

    def distance_to_wall(self, first: int = 0, n: int | None = None) -> np.ndarray:
        n = self.N - first if n is None else n
        out = np.full((n, self.Q - 1), -1.0, np.float64)
        lo = np.searchsorted(self._b_local, first)
        hi = np.searchsorted(self._b_local, first + n)
        out[self._b_local[lo:hi] - first] = self._b_dist[lo:hi].astype(np.float64)
        return out.reshape(-1)

    def wall_normal(self, first: int = 0, n: int | None = None) -> np.ndarray:
        n = self.N - first if n is None else n
        out = np.full((n, 3), np.inf, np.float64)
        lo = np.searchsorted(self._b_local, first)
        hi = np.searchsorted(self._b_local, first + n)
        out[self._b_local[lo:hi] - first] = self._b_normal[lo:hi].astype(np.float64)
        return out.reshape(-1)

    def tables(self) -> dict:
        """Everything in one dict, fully materialised (small domains / tests)."""
        counts = np.concatenate([self.mid, self.edge]).astype(np.int64)
        return dict(Q=self.Q, N=self.N, totalSharedFs=self.totalSharedFs, counts=counts, mid=self.mid.copy(),
                    edge=self.edge.copy(), neighbourIndices=self.neighbour_indices(),
                    wallMask=self.wallMask, ioletMask=self.ioletMask, siteType=self.siteType,
                    ioletId=self.ioletId, distanceToWall=self.distance_to_wall(), wallNormal=self.wall_normal(),
                    globalCoords=self.globalCoords.reshape(-1), inputIndex=self.inputIndex,
                    streamingIndices=self.streamingIndices, procs=self.procs)


class DomainBuilder:
    def __init__(self, geom: Geometry, Q: int, rank_of_site=None, nranks: int = 1, chunk: int = 1 << 22):
        self.geom = geom
        self.Q = Q
        self.R = nranks
        self.c = lattice_vectors(Q)
        self.inv = inverse_directions(Q)
        N = geom.n_sites
        self.rank_of_site = (np.zeros(N, np.int32) if rank_of_site is None
                             else np.ascontiguousarray(rank_of_site, np.int32))
        self.lookup = _Lookup(geom)
        B = geom.block_size
        coords = geom.coords.astype(np.int64)

        # traversal position of every site: Morton order of its block, then z-fastest in block
        bc = coords // B
        sc = coords % B
        bd = geom.block_dims.astype(np.int64)
        gkey = ((bc[:, 0] * bd[1] + bc[:, 1]) * bd[2] + bc[:, 2]) * (B ** 3) + (sc[:, 0] * B + sc[:, 1]) * B + sc[:, 2]
        if N and np.all(np.diff(gkey) > 0):
            # .gmy order: only the blocks need re-ordering (O(N), no big sort)
            blk = gkey // (B ** 3)
            starts = np.concatenate([[0], np.nonzero(np.diff(blk))[0] + 1])
            ub = blk[starts]
            cnts = np.diff(np.concatenate([starts, [N]]))
            ijk = np.stack([ub // (bd[1] * bd[2]), (ub // bd[2]) % bd[1], ub % bd[2]], 1)
            mo = np.argsort(morton(ijk), kind="stable")
            off = np.empty(ub.size, np.int64)
            off[mo] = np.concatenate([[0], np.cumsum(cnts[mo])[:-1]])
            trav = np.repeat(off - starts, cnts) + np.arange(N, dtype=np.int64)
        else:
            mkey = morton(bc)
            order = np.lexsort(((sc[:, 0] * B + sc[:, 1]) * B + sc[:, 2], mkey))
            trav = np.empty(N, np.int64)
            trav[order] = np.arange(N)
        del bc, sc, gkey
        self.trav = trav
        order_all = np.empty(N, np.int64)
        order_all[trav] = np.arange(N, dtype=np.int64)

        # collision type from the lattice's subset of cut links (SiteDataBare.cc:23-140)
        lk = gmy_link_of_direction(Q)[1:]
        bt = geom.btype[:, lk]
        bits = (np.uint32(1) << np.arange(Q - 1, dtype=np.uint32))
        wall_b = ((bt == CUT_WALL) * bits).sum(1).astype(np.uint32)
        iol_b = (((bt == CUT_INLET) | (bt == CUT_OUTLET)) * bits).sum(1).astype(np.uint32)
        had_in = (bt == CUT_INLET).any(1)
        had_out = (bt == CUT_OUTLET).any(1)
        type_b = np.where(had_in, 2, np.where(had_out, 3, 1)).astype(np.int32)
        # ioletId = id of the LAST iolet link in direction order
        io_any = (bt == CUT_INLET) | (bt == CUT_OUTLET)
        last = (Q - 2) - np.argmax(io_any[:, ::-1], axis=1)
        ioid_b = np.where(io_any.any(1), geom.biolet[:, lk][np.arange(bt.shape[0]), last], -1).astype(np.int32)
        coll_b = np.where(wall_b == 0, np.where(type_b == 1, 0, np.where(type_b == 2, 2, 3)),
                          np.where(type_b == 1, 1, np.where(type_b == 2, 4, 5)))
        coll = np.zeros(N, np.int8)
        coll[geom.bsite] = coll_b
        brec = np.full(N, -1, np.int64)
        brec[geom.bsite] = np.arange(geom.bsite.size)

        # remote links: (site, direction, neighbour site) with the neighbour on another rank
        rs, rd, rn = [], [], []
        is_edge = np.zeros(N, bool)
        if nranks > 1 and self.lookup.dense is not None:
            rgrid = np.full(self.lookup.dense.size, -1, np.int16)
            rgrid[self.lookup.key_of_site] = self.rank_of_site.astype(np.int16)
            for s0 in range(0, N, chunk):
                s1 = min(N, s0 + chunk)
                key0 = self.lookup.key_of_site[s0:s1]
                myrank = self.rank_of_site[s0:s1].astype(np.int16)
                for d in range(1, Q):
                    off = self.lookup.key_offset(self.c[d])
                    nr = rgrid[key0 + off]
                    idx = np.nonzero((nr >= 0) & (nr != myrank))[0]
                    if idx.size:
                        rs.append(idx + s0)
                        rd.append(np.full(idx.size, d, np.int64))
                        rn.append(self.lookup.dense[key0[idx] + off].astype(np.int64))
                        is_edge[idx + s0] = True
            del rgrid
        elif nranks > 1:
            for s0 in range(0, N, chunk):
                s1 = min(N, s0 + chunk)
                base = coords[s0:s1]
                myrank = self.rank_of_site[s0:s1]
                for d in range(1, Q):
                    nb = self.lookup(base + self.c[d])
                    rem = (nb >= 0)
                    rem[rem] = self.rank_of_site[nb[rem]] != myrank[rem]
                    if rem.any():
                        idx = np.nonzero(rem)[0]
                        rs.append(idx + s0)
                        rd.append(np.full(idx.size, d, np.int64))
                        rn.append(nb[idx])
                        is_edge[idx + s0] = True
        rs = np.concatenate(rs) if rs else np.zeros(0, np.int64)
        rd = np.concatenate(rd) if rd else np.zeros(0, np.int64)
        rn = np.concatenate(rn) if rn else np.zeros(0, np.int64)

        # local numbering: stable bucket sort of the traversal order by (edge, type)
        self.local_of_input = np.empty(N, np.int64)
        self.domains = []
        bucket = is_edge.astype(np.int64) * 6 + coll
        for r in range(nranks):
            D = RankDomain()
            D._builder = self
            D.Q, D.rank, D.nranks = Q, r, nranks
            mine = order_all if nranks == 1 else order_all[self.rank_of_site[order_all] == r]  # traversal order
            o = np.argsort(bucket[mine].astype(np.int8), kind="stable")
            local = mine[o]
            D.N = int(local.size)
            D.inputIndex = local.astype(np.int64)
            self.local_of_input[local] = np.arange(D.N)
            cnt = np.bincount(bucket[mine], minlength=12).astype(np.int64)
            D.mid, D.edge = cnt[:6], cnt[6:]
            D.globalCoords = coords[local]
            br = brec[local]
            hasb = br >= 0
            D.wallMask = np.zeros(D.N, np.uint32)
            D.ioletMask = np.zeros(D.N, np.uint32)
            D.siteType = np.ones(D.N, np.int32)
            D.ioletId = np.full(D.N, -1, np.int32)
            D.wallMask[hasb] = wall_b[br[hasb]]
            D.ioletMask[hasb] = iol_b[br[hasb]]
            D.siteType[hasb] = type_b[br[hasb]]
            D.ioletId[hasb] = ioid_b[br[hasb]]
            D._b_local = np.nonzero(hasb)[0].astype(np.int64)
            bb = br[hasb]
            dist = geom.bdist[bb][:, lk].astype(np.float32)
            dist = np.where(geom.btype[bb][:, lk] != 0, dist, np.float32(-1.0))
            D._b_dist = dist
            nrm = geom.bnormal[bb].astype(np.float32)
            D._b_normal = np.where(geom.bnavail[bb][:, None].astype(bool), nrm, np.float32(np.inf))
            self.domains.append(D)

        # halo tables (Domain.cc:247-285, 404-419, 507-580)
        if rs.size:
            r_rank = self.rank_of_site[rs]
            n_rank = self.rank_of_site[rn]
            # sender-side list order: traversal order of the site, then direction
            lo = np.lexsort((rd, trav[rs]))
            rs, rd, rn, r_rank, n_rank = rs[lo], rd[lo], rn[lo], r_rank[lo], n_rank[lo]
            for r in range(nranks):
                D = self.domains[r]
                m = r_rank == r
                if not m.any():
                    continue
                srs, srd, srn, snr = rs[m], rd[m], rn[m], n_rank[m]
                D.totalSharedFs = int(srs.size)
                # neighbouringProcs: first-encounter order over edge sites in traversal order
                uniq, firstpos = np.unique(snr, return_index=True)
                procs = uniq[np.argsort(firstpos)]
                counts = np.array([(snr == p).sum() for p in procs], np.int64)
                firsts = D.N * Q + 1 + np.concatenate([[0], np.cumsum(counts)[:-1]])
                D.procs = np.stack([procs.astype(np.int64), counts, firsts], 1)
                keys, slots, stream = [], [], []
                f_count = D.N * Q
                for p, cnt in zip(procs, counts):
                    if p > r:  # this rank's own list is authoritative
                        sel = snr == p
                        site = srs[sel]
                        l = srd[sel]
                    else:  # the lower rank's list, flipped: (site + c_l, inverse l)
                        sel = (r_rank == p) & (n_rank == r)
                        site = rn[sel]
                        l = self.inv[rd[sel]]
                    contig = self.local_of_input[site]
                    slot = f_count + 1 + np.arange(cnt, dtype=np.int64)
                    f_count += int(cnt)
                    keys.append(contig * Q + l)
                    slots.append(slot)
                    stream.append(contig * Q + self.inv[l])
                keys = np.concatenate(keys)
                slots = np.concatenate(slots)
                o = np.argsort(keys, kind="stable")
                D._send_key, D._send_slot = keys[o], slots[o]
                D.streamingIndices = np.concatenate(stream).astype(np.int64)


def _gzs_site_halo(self):
    """The GZS site halo of every rank (GuoZhengShi.h:38-99 + NeighbouringDataManager::ShareNeeds):
    ``need[r]`` = one row per wall link: (local site, direction towards the neighbour, owner rank,
    owner's local site); ``serve[r]`` = (requester rank, local site), one row per (requester, site).
    Both grouped by the other rank ascending and, inside a group, ordered by the requesting site's
    global (x, y, z) then direction -- an order both sides can compute from coordinates alone; a
    site several links ask for is served once, where the requester first names it."""
    if getattr(self, "_gzs", None) is not None:
        return self._gzs
    Q = self.Q
    coords = self.geom.coords.astype(np.int64)
    dims = self.geom.block_dims.astype(np.int64) * self.geom.block_size
    rec = []  # requester rank, owner rank, coord key, direction, requester input site, owner input site
    for D in self.domains:
        ws = np.nonzero(D.wallMask != 0)[0]
        if ws.size == 0:
            continue
        inp = D.inputIndex[ws]
        wm, im = D.wallMask[ws], D.ioletMask[ws]
        for d in range(1, Q):
            opp = int(self.inv[d])
            m = ((wm >> np.uint32(d - 1)) & 1).astype(bool)
            m &= ~((wm >> np.uint32(opp - 1)) & 1).astype(bool)
            m &= ~((im >> np.uint32(opp - 1)) & 1).astype(bool)
            if not m.any():
                continue
            src = inp[m]
            nb = self.lookup(coords[src] + self.c[opp])
            ok = nb >= 0
            ok[ok] = self.rank_of_site[nb[ok]] != D.rank
            if not ok.any():
                continue
            src, nb = src[ok], nb[ok]
            c = coords[src]
            key = (c[:, 0] * dims[1] + c[:, 1]) * dims[2] + c[:, 2]
            rec.append(np.stack([np.full(src.size, D.rank), self.rank_of_site[nb].astype(np.int64), key,
                                 np.full(src.size, opp), src, nb], 1))
    need = {D.rank: np.zeros((0, 4), np.int64) for D in self.domains}
    serve = {D.rank: np.zeros((0, 2), np.int64) for D in self.domains}
    if rec:
        rec = np.concatenate(rec, 0).astype(np.int64)
        for D in self.domains:
            mine = rec[rec[:, 0] == D.rank]
            o = np.lexsort((mine[:, 3], mine[:, 2], mine[:, 1]))
            mine = mine[o]
            need[D.rank] = np.stack([self.local_of_input[mine[:, 4]], mine[:, 3], mine[:, 1],
                                     self.local_of_input[mine[:, 5]]], 1)
            theirs = rec[rec[:, 1] == D.rank]
            o = np.lexsort((theirs[:, 3], theirs[:, 2], theirs[:, 0]))
            theirs = theirs[o]
            # one row per (requester, site), in the order the requester first names it: links that
            # extrapolate from the same site share a ghost row (NeighbouringDataManager.cc:27-39)
            _, firsts = np.unique(theirs[:, [0, 5]], axis=0, return_index=True)
            theirs = theirs[np.sort(firsts)]
            serve[D.rank] = np.stack([theirs[:, 0], self.local_of_input[theirs[:, 5]]], 1)
    self._gzs = (need, serve)
    return self._gzs


DomainBuilder.gzs_site_halo = _gzs_site_halo


def _local_grid(self, rank=0):
    """voxel -> local site id on ``rank``; -1 solid, -2 fluid on another rank."""
    cache = self.__dict__.setdefault("_lgrids", {})
    if rank not in cache:
        cache.clear()  # one at a time: these are big
        g = np.full(self.lookup.dense.size, -1, np.int32)
        own = self.rank_of_site == rank
        vals = np.where(own, self.local_of_input, -2).astype(np.int32)
        g[self.lookup.key_of_site] = vals
        cache[rank] = g
    return cache[rank]


DomainBuilder.local_grid = _local_grid


def build_domains(geom: Geometry, Q: int, rank_of_site=None, nranks: int = 1) -> list:
    """One RankDomain per rank, tables identical to the reference's ``geometry::Domain``."""
    return DomainBuilder(geom, Q, rank_of_site, nranks).domains
