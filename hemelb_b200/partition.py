"""METIS-free weighted k-way partition of the fluid-site blocks.

The reference refines its ``BasicDecomposition`` with ParMETIS (``OptimisedDecomposition.cc:138-154``:
``ParMETIS_V3_PartKway`` over the site graph, vertex weights by collision type from
``DecompositionWeights.h.in:25-62``).  ParMETIS is not in this image and no reference test pins a
partition, so parity is unpinned by design: any assignment is a valid input of the Domain builder
(``hlb_dom_set_partition_blocks`` / ``rank_of_site``), the tables are bit-exact *given* it.

What is built here is the same objective without METIS, at block granularity (the unit the
reference's own BasicDecomposition and the device Domain builder partition by):

1. vertex weight of a block = sum over its sites of the weight of their collision type -- the
   reference's table, or one measured on B200 (ns per site update of each range kernel);
2. initial partition = the reference's recursive bisection of the Morton-ordered blocks
   (``BasicDecomposition.cc:21-96``) on the *weighted* cumulative load instead of site counts;
3. k-way boundary refinement (Fiduccia-Mattheyses style, greedy, deterministic): blocks on a part
   boundary move to a face-adjacent part when that lowers the edge cut (face-adjacent block pairs
   in different parts, weighted by the smaller block's load^(2/3) -- a proxy for the links cut)
   without pushing any part above (1 + tolerance) x the mean load; then under-/over-loaded parts
   trade boundary blocks to tighten the balance.

Plain numpy; a few million blocks (a 1e9-site tree has ~3e6 non-empty 8^3 blocks) take seconds.
"""
from __future__ import annotations

import numpy as np

from .geometry import basic_decomposition_blocks, morton

# DecompositionWeights.h.in:25-62 -- {bulk, wall[WALL], iolet[BC]} for HEMELB_COMPUTE_ARCHITECTURE
REFERENCE_WEIGHTS = {
    "AMDBULLDOZER": dict(bulk=4, SBB=5, BFL=8, GZS=28, NASH=16, LADD=16),
    "INTELSANDYBRIDGE": dict(bulk=4, SBB=5, BFL=8, GZS=28, NASH=16, LADD=16),
    "NEUTRAL": dict(bulk=1, SBB=1, BFL=1, GZS=1, NASH=1, LADD=1),
    "ISBFILEVELOCITYINLET": dict(bulk=4, SBB=5, BFL=8, GZS=28, NASH=16, LADD=48),
    # measured on one B200 (profiles/r01_tree_1gpu_launches.csv, r01_cfg3_launches.csv): ns per site
    # update of the mid-fluid kernel = 1 unit of 4; wall sites incl. their PostStep / per-link kernel
    "B200": dict(bulk=4, SBB=8, BFL=14, GZS=60, NASH=20, LADD=24),
}


def site_weights(wall: str, inlet: str, outlet: str, architecture: str = "B200") -> np.ndarray:
    """``hemelbSiteWeights[6]``: weight per collision type {bulk, wall, inlet, outlet, inlet+wall,
    outlet+wall} (``DecompositionWeights.h.in:56-61`` -- the combined types take the iolet's weight)."""
    w = REFERENCE_WEIGHTS[architecture]
    return np.array([w["bulk"], w[wall], w[inlet], w[outlet], w[inlet], w[outlet]], np.float64)


def block_loads(block_of_site: np.ndarray, site_type: np.ndarray, weights: np.ndarray, n_blocks: int) -> np.ndarray:
    """Vertex weight of each block (``OptimisedDecomposition::PopulateVertexWeightData`` summed per block)."""
    return np.bincount(block_of_site, weights=weights[site_type], minlength=n_blocks)


def face_adjacency(ijk: np.ndarray):
    """Pairs (a, b), a < b, of non-empty blocks that share a face."""
    ijk = np.asarray(ijk, np.int64)
    lo = ijk.min(0)
    ext = ijk.max(0) - lo + 3
    key = ((ijk[:, 0] - lo[0] + 1) * ext[1] + (ijk[:, 1] - lo[1] + 1)) * ext[2] + (ijk[:, 2] - lo[2] + 1)
    order = np.argsort(key)
    skey = key[order]
    pairs = []
    for step in (ext[1] * ext[2], ext[2], 1):
        want = key + step
        pos = np.searchsorted(skey, want)
        pos[pos >= skey.size] = skey.size - 1
        hit = skey[pos] == want
        pairs.append(np.stack([np.nonzero(hit)[0], order[pos[hit]]], 1))
    p = np.concatenate(pairs, 0)
    return np.sort(p, 1)


def weighted_bisection(ijk: np.ndarray, loads: np.ndarray, nranks: int) -> np.ndarray:
    """Step 2: BasicDecomposition's recursive bisection on weighted loads (integerised so that the
    reference's float32 arithmetic is followed literally)."""
    scale = 1.0 if loads.max() <= 0 else min(1.0, float(2 ** 40) / float(loads.sum()))
    return basic_decomposition_blocks(np.asarray(ijk, np.int64), np.maximum(1, np.round(loads * scale)).astype(np.int64), nranks)


def edge_cut(pairs: np.ndarray, edge_w: np.ndarray, part: np.ndarray) -> float:
    return float(edge_w[part[pairs[:, 0]] != part[pairs[:, 1]]].sum())


def imbalance(loads: np.ndarray, part: np.ndarray, nranks: int) -> float:
    """max part load / mean part load."""
    pl = np.bincount(part, weights=loads, minlength=nranks)
    return float(pl.max() / pl.mean())


def refine(ijk, loads, part, nranks, tolerance=0.03, passes=8):
    """Step 3.  Returns the refined part array (a copy)."""
    part = np.asarray(part, np.int32).copy()
    pairs = face_adjacency(ijk)
    if pairs.size == 0:
        return part
    edge_w = np.minimum(loads[pairs[:, 0]], loads[pairs[:, 1]]) ** (2.0 / 3.0)
    n = part.size
    # CSR adjacency
    src = np.concatenate([pairs[:, 0], pairs[:, 1]])
    dst = np.concatenate([pairs[:, 1], pairs[:, 0]])
    ew = np.concatenate([edge_w, edge_w])
    o = np.argsort(src, kind="stable")
    src, dst, ew = src[o], dst[o], ew[o]
    start = np.searchsorted(src, np.arange(n + 1))
    mean = loads.sum() / nranks
    cap = (1.0 + tolerance) * mean
    pl = np.bincount(part, weights=loads, minlength=nranks)
    for _ in range(passes):
        moved = 0
        boundary = np.unique(src[part[src] != part[dst]])
        # heaviest-gain-first would need a priority queue; a deterministic sweep in Morton order is
        # enough here (parts are contiguous curve segments to begin with)
        boundary = boundary[np.argsort(morton(np.asarray(ijk, np.int64)[boundary]), kind="stable")]
        for b in boundary:
            a0, a1 = start[b], start[b + 1]
            nb, w = dst[a0:a1], ew[a0:a1]
            own = part[b]
            conn = np.bincount(part[nb], weights=w, minlength=nranks)
            conn_own = conn[own]
            conn[own] = -1.0
            t = int(np.argmax(conn))
            gain = conn[t] - conn_own
            if conn[t] <= 0:
                continue
            fits = pl[t] + loads[b] <= cap
            relieves = pl[own] > cap and pl[t] + loads[b] < pl[own]
            if (gain > 0 and fits) or (relieves and gain >= -0.25 * conn_own):
                part[b] = t
                pl[own] -= loads[b]
                pl[t] += loads[b]
                moved += 1
        if not moved:
            break
    return part


def weighted_kway(ijk, loads, nranks, tolerance=0.03, refine_passes=8):
    """Block -> rank.  ``ijk``: (n, 3) coordinates of the non-empty blocks; ``loads``: their vertex weights."""
    ijk = np.asarray(ijk, np.int64)
    loads = np.asarray(loads, np.float64)
    if ijk.shape[0] < nranks:
        raise ValueError("More ranks than blocks")
    part = weighted_bisection(ijk, loads, nranks)
    if refine_passes > 0 and nranks > 1:
        part = refine(ijk, loads, part, nranks, tolerance, refine_passes)
    return part.astype(np.int32)


def quality(ijk, loads, part, nranks):
    pairs = face_adjacency(ijk)
    edge_w = np.minimum(loads[pairs[:, 0]], loads[pairs[:, 1]]) ** (2.0 / 3.0) if pairs.size else np.zeros(0)
    return dict(imbalance=imbalance(loads, part, nranks), edge_cut=edge_cut(pairs, edge_w, part) if pairs.size else 0.0,
                parts=int(np.unique(part).size))


def partition_geometry(geom, site_type, wall="BFL", inlet="NASH", outlet="NASH", nranks=2, architecture="B200",
                       tolerance=0.03):
    """Site -> rank for a host ``Geometry`` (whole blocks), with the quality figures of the weighted
    k-way partition and of the reference's BasicDecomposition on the same weights."""
    B = geom.block_size
    bc = (geom.coords // B).astype(np.int64)
    bd = geom.block_dims.astype(np.int64)
    gmy = (bc[:, 0] * bd[1] + bc[:, 1]) * bd[2] + bc[:, 2]
    uniq, inv = np.unique(gmy, return_inverse=True)
    ijk = np.stack([uniq // (bd[1] * bd[2]), (uniq // bd[2]) % bd[1], uniq % bd[2]], 1)
    loads = block_loads(inv, np.asarray(site_type), site_weights(wall, inlet, outlet, architecture), uniq.size)
    counts = np.bincount(inv, minlength=uniq.size)
    part = weighted_kway(ijk, loads, nranks, tolerance)
    basic = basic_decomposition_blocks(ijk, counts, nranks)
    return part[inv].astype(np.int32), dict(weighted=quality(ijk, loads, part, nranks),
                                            basic=quality(ijk, loads, basic, nranks))
