"""The METIS-free weighted k-way block partitioner (hemelb_b200/partition.py): balance
and cut properties against the reference's BasicDecomposition on the same vertex weights, and that
its output is a valid input of the Domain builder (emulated multi-rank run == single-rank run)."""
import numpy as np
import pytest

import oracle as O
from hemelb_b200 import geometry as G
from hemelb_b200 import partition as P
from hemelb_b200.domain import build_domains
from tests.cases import anisotropic_f, geometry, iolets_for


def collision_types(geom, Q=19):
    """Collision type 0..5 of every input site (Domain.cc:186-207), from the single-rank tables."""
    dom = build_domains(geom, Q)[0]
    t = dom.tables()
    wall = np.asarray(t["wallMask"]) != 0
    st = np.asarray(t["siteType"])
    local = np.where(st == 2, np.where(wall, 4, 2), np.where(st == 3, np.where(wall, 5, 3), np.where(wall, 1, 0)))
    out = np.empty(geom.n_sites, np.int64)
    out[np.asarray(t["inputIndex"])] = local
    return out


def test_reference_weight_table():
    """DecompositionWeights.h.in:25-62 for the default architecture and BFL / GZS walls."""
    assert P.site_weights("BFL", "NASH", "NASH", "AMDBULLDOZER").tolist() == [4, 8, 16, 16, 16, 16]
    assert P.site_weights("GZS", "LADD", "NASH", "ISBFILEVELOCITYINLET").tolist() == [4, 28, 48, 16, 48, 16]
    assert P.site_weights("SBB", "NASH", "NASH", "NEUTRAL").tolist() == [1] * 6


def test_face_adjacency():
    ijk = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [2, 2, 2], [1, 1, 0]])
    got = {tuple(p) for p in P.face_adjacency(ijk).tolist()}
    assert got == {(0, 1), (0, 2), (0, 3), (1, 5), (2, 5)}


@pytest.mark.parametrize("geom_name,nranks", [("tree", 4), ("tree", 7), ("sac", 3), ("cylinder_long", 5)])
def test_weighted_kway_properties(geom_name, nranks):
    geom = geometry(geom_name)
    types = collision_types(geom)
    for wall, arch in (("BFL", "B200"), ("GZS", "AMDBULLDOZER")):
        rank, q = P.partition_geometry(geom, types, wall, "NASH", "NASH", nranks, arch, tolerance=0.05)
        assert rank.shape == (geom.n_sites,) and rank.min() == 0 and rank.max() == nranks - 1
        assert q["weighted"]["parts"] == nranks
        # whole blocks
        B = geom.block_size
        key = (geom.coords // B).astype(np.int64) @ np.array([1 << 40, 1 << 20, 1])
        for k in np.unique(key)[:50]:
            assert np.unique(rank[key == k]).size == 1
        # never worse balanced than the reference's count-based bisection on the same weights,
        # within one block of the tolerance
        loads = np.bincount(np.unique(key, return_inverse=True)[1],
                            weights=P.site_weights(wall, "NASH", "NASH", arch)[types])
        slack = loads.max() / (loads.sum() / nranks)
        assert q["weighted"]["imbalance"] <= max(q["basic"]["imbalance"], 1.05 + slack) + 1e-12
        assert q["weighted"]["imbalance"] <= 1.05 + slack
        # deterministic
        rank2, _ = P.partition_geometry(geom, types, wall, "NASH", "NASH", nranks, arch, tolerance=0.05)
        assert np.array_equal(rank, rank2)


def test_refinement_does_not_increase_the_cut():
    geom = geometry("tree")
    types = collision_types(geom)
    B = geom.block_size
    bc = (geom.coords // B).astype(np.int64)
    bd = geom.block_dims.astype(np.int64)
    gmy = (bc[:, 0] * bd[1] + bc[:, 1]) * bd[2] + bc[:, 2]
    uniq, inv = np.unique(gmy, return_inverse=True)
    ijk = np.stack([uniq // (bd[1] * bd[2]), (uniq // bd[2]) % bd[1], uniq % bd[2]], 1)
    loads = P.block_loads(inv, types, P.site_weights("BFL", "NASH", "NASH"), uniq.size)
    for nranks in (2, 4, 8):
        first = P.weighted_bisection(ijk, loads, nranks)
        done = P.refine(ijk, loads, first, nranks, tolerance=0.05)
        q0, q1 = P.quality(ijk, loads, first, nranks), P.quality(ijk, loads, done, nranks)
        assert q1["edge_cut"] <= q0["edge_cut"] + 1e-9 or q1["imbalance"] < q0["imbalance"]
        assert q1["parts"] == nranks


def test_site_graph_is_the_parmetis_graph_of_the_reference():
    """OptimisedDecomposition.cc:311-379 on four_cube (a 4^3 box of fluid): one adjacency per lattice
    direction with a fluid site at the other end, in direction order; symmetric."""
    geom = G.four_cube()
    for Q, vec in ((15, P.lattice_vectors(15)), (19, P.lattice_vectors(19)), (27, P.lattice_vectors(27))):
        xadj, adjncy = P.site_graph(geom, Q)
        want = sum(int(np.prod(4 - np.abs(vec[l]))) for l in range(1, Q))  # pairs at offset c_l inside the box
        assert xadj[0] == 0 and xadj[-1] == adjncy.size == want
        c = geom.coords.astype(np.int64)
        site_at = {tuple(x): i for i, x in enumerate(c.tolist())}
        for i in (0, 21, 42, 63):
            expect = [site_at[tuple(c[i] + vec[l])] for l in range(1, Q) if tuple(c[i] + vec[l]) in site_at]
            assert adjncy[xadj[i]:xadj[i + 1]].tolist() == expect
        src = np.repeat(np.arange(geom.n_sites), np.diff(xadj))
        assert {(a, b) for a, b in zip(src.tolist(), adjncy.tolist())} == {(b, a) for a, b in zip(src.tolist(), adjncy.tolist())}
    # the reference numbers vertices by octree block, then by site id in the block
    geom = geometry("tree")
    order = P.reference_vertex_order(geom)
    assert np.array_equal(np.sort(order), np.arange(geom.n_sites))
    B = geom.block_size
    m = G.morton(geom.coords[order].astype(np.int64) // B).astype(np.int64)
    assert (np.diff(m) >= 0).all()
    s = geom.coords[order].astype(np.int64) % B
    sid = (s[:, 0] * B + s[:, 1]) * B + s[:, 2]
    assert (np.diff(sid)[np.diff(m) == 0] > 0).all()
    # solid neighbours and the lattice's edge are no vertices
    xadj, adjncy = P.site_graph(geom, 19)
    assert adjncy.min() >= 0 and adjncy.max() < geom.n_sites and (np.diff(xadj) <= 18).all()


@pytest.mark.parametrize("geom_name,nranks", [("tree", 2), ("tree", 4), ("tree", 8), ("sac", 3), ("sac", 8),
                                              ("cylinder_long", 5)])
def test_site_granular_refinement(geom_name, nranks):
    """Step 4 (the ParMETIS stage, OptimisedDecomposition.cc:100-154): balance within ubvec (or one
    site of the mean on these small cases), no rank left empty, deterministic, and from a balanced
    start the number of cut links never grows."""
    geom = geometry(geom_name)
    types = collision_types(geom)
    for wall, arch in (("BFL", "B200"), ("GZS", "AMDBULLDOZER")):
        rank, q = P.partition_sites(geom, types, 19, wall, "NASH", "NASH", nranks, arch, ubvec=1.001)
        assert rank.dtype == np.int32 and rank.shape == (geom.n_sites,)
        assert q["sites"]["parts"] == nranks
        vw = P.site_weights(wall, "NASH", "NASH", arch)[types]
        mean = vw.sum() / nranks
        assert q["sites"]["imbalance"] <= max(1.001, 1 + vw.max() / mean) + 1e-12
        assert q["sites"]["imbalance"] <= q["blocks"]["imbalance"] + 1e-12
        rank2, _ = P.partition_sites(geom, types, 19, wall, "NASH", "NASH", nranks, arch, ubvec=1.001)
        assert np.array_equal(rank, rank2)
        xadj, adjncy = P.site_graph(geom, 19)
        again = P.refine_sites(xadj, adjncy, vw, rank, nranks, 1.001)
        assert P.site_cut(xadj, adjncy, again) <= q["sites"]["edge_cut"]
        assert P.site_quality(xadj, adjncy, vw, again, nranks)["imbalance"] <= max(1.001, 1 + vw.max() / mean) + 1e-12


def test_site_refinement_shrinks_the_halo():
    """The cut links are the halo: totalSharedFs of the built tables (FieldData.cc:27-39 sizes) equals
    the directed cut of the site graph, and the site stage does not leave it above the block stage on
    the tree when both are balanced."""
    geom, Q, R = geometry("tree"), 19, 4
    types = collision_types(geom)
    rank, q = P.partition_sites(geom, types, Q, nranks=R)
    xadj, adjncy = P.site_graph(geom, Q)
    doms = build_domains(geom, Q, rank, R)
    assert sum(int(d.tables()["totalSharedFs"]) for d in doms) == 2 * P.site_cut(xadj, adjncy, rank)
    assert q["sites"]["edge_cut"] <= q["blocks"]["edge_cut"]


@pytest.mark.parametrize("stage", ["blocks", "sites"])
def test_partition_is_a_valid_domain_input(stage):
    """Tables built for the partition drive a 4-rank emulated run that equals the single-rank run."""
    geom = geometry("tree")
    Q = 19
    if stage == "blocks":
        rank, _ = P.partition_geometry(geom, collision_types(geom), "BFL", "NASH", "NASH", 4)
    else:
        rank, _ = P.partition_sites(geom, collision_types(geom), Q, "BFL", "NASH", "NASH", 4)
        B = geom.block_size
        key = (geom.coords // B).astype(np.int64) @ np.array([1 << 40, 1 << 20, 1])
        assert any(np.unique(rank[key == k]).size > 1 for k in np.unique(key))  # cuts through blocks
    inlets, outlets = iolets_for(geom, "NASH", "NASH")
    one = O.OracleSim(O.OracleDomains(geom, Q), "LBGK", "BFL", tau=0.8, inlets=inlets, outlets=outlets)
    doms = O.OracleDomains(geom, Q, rank, 4)
    many = O.OracleSim(doms, "LBGK", "BFL", tau=0.8, inlets=inlets, outlets=outlets)
    t1 = O.OracleDomains(geom, Q).tables(0)
    f_input = anisotropic_f(geom.n_sites, Q, 0)[:geom.n_sites * Q].reshape(-1, Q)  # indexed by input site
    f0 = np.zeros(one.f_size(0))
    f0[:t1["N"] * Q] = f_input[np.asarray(t1["inputIndex"])].ravel()
    one.set_f(f0)
    for r in range(4):
        t = doms.tables(r)
        f = np.zeros(many.f_size(r))
        f[:t["N"] * Q] = f_input[np.asarray(t["inputIndex"])].ravel()
        many.set_f(f, r)
    one.step(5)
    many.step(5)
    got = np.zeros_like(f_input)
    for r in range(4):
        t = doms.tables(r)
        got[np.asarray(t["inputIndex"])] = many.get_f(r)[:t["N"] * Q].reshape(-1, Q)
    want = np.zeros_like(f_input)
    want[np.asarray(t1["inputIndex"])] = one.get_f(0)[:t1["N"] * Q].reshape(-1, Q)
    assert np.array_equal(got, want)


def test_coordinate_bisection():
    """Exact weighted-median splits across the longest extent; floor(n/2) : n - floor(n/2) ranks."""
    rng = np.random.default_rng(20261017)
    pts = rng.integers(0, 50, size=(4000, 3)) * np.array([4, 1, 1])  # longest along x
    w = rng.integers(1, 5, size=4000).astype(np.float64)
    for n in (1, 2, 3, 5, 8):
        part = P.coordinate_bisection(pts, w, n)
        assert part.min() == 0 and part.max() == n - 1
        loads = np.bincount(part, weights=w, minlength=n)
        assert loads.max() / loads.mean() <= 1 + n * w.max() / w.sum() * n + 1e-12
        assert np.array_equal(part, P.coordinate_bisection(pts, w, n))
    two = P.coordinate_bisection(pts, w, 2)
    assert pts[two == 0, 0].max() <= pts[two == 1, 0].min()  # one cut across x
    # inertial: a slanted rod is cut across its own axis, whatever the coordinate axes are
    s_ = np.arange(3000)
    rod = np.stack([s_, s_ // 2, s_ // 3], 1) + rng.integers(-3, 4, size=(3000, 3))
    ib = P.coordinate_bisection(rod, np.ones(3000), 4, inertial=True)
    along = rod @ np.array([1.0, 0.5, 1 / 3.0])
    for k in range(3):
        assert np.percentile(along[ib == k], 99) < np.percentile(along[ib == k + 1], 1) + 30
    assert np.bincount(ib).tolist() == [750] * 4
    assert np.array_equal(ib, P.coordinate_bisection(rod, np.ones(3000), 4, inertial=True))
    # more ranks than distinct coordinates still leaves no rank empty
    line = np.stack([np.arange(6), np.zeros(6, int), np.zeros(6, int)], 1)
    assert np.array_equal(np.sort(P.coordinate_bisection(line, np.array([100., 1, 1, 1, 1, 1]), 6)), np.arange(6))


@pytest.mark.parametrize("geom_name,nranks", [("tree", 4), ("cylinder_long", 4), ("sac", 3)])
def test_best_start_is_never_worse_and_valid(geom_name, nranks):
    """"best" keeps the smaller cut of the two starts (block stage and site stage), and the tables
    built for its partitions are the oracle's (any site -> rank map is a valid Domain input)."""
    from tests.test_domain_tables import _same_tables
    geom = geometry(geom_name)
    types = collision_types(geom)
    res = {ini: P.partition_sites(geom, types, 19, nranks=nranks, initial=ini) for ini in P.STARTS + ("best",)}
    vw = P.site_weights("BFL", "NASH", "NASH")[types]
    bound = max(1.001, 1 + vw.max() / (vw.sum() / nranks)) + 1e-12
    ok = [k for k in P.STARTS if res[k][1]["sites"]["imbalance"] <= bound]
    assert res["best"][1]["sites"]["edge_cut"] == min(res[k][1]["sites"]["edge_cut"] for k in ok)
    assert res["best"][1]["initial"] in ok
    _same_tables(geom, 19, res["best"][0], nranks)
    blocks_best, qb = P.partition_geometry(geom, types, nranks=nranks, initial="best", tolerance=0.05)
    for ini in P.STARTS:
        _, q = P.partition_geometry(geom, types, nranks=nranks, initial=ini, tolerance=0.05)
        slack = 0.05 + (vw.max() * geom.block_size ** 3) / (vw.sum() / nranks)
        if q["weighted"]["imbalance"] <= 1 + slack:
            assert qb["weighted"]["edge_cut"] <= q["weighted"]["edge_cut"] + 1e-9
    _same_tables(geom, 19, blocks_best, nranks)


def test_coordinate_start_cuts_tubes_across():
    """On the long cylinder the coordinate start gives z-slabs: two neighbours at most, and far fewer
    cut links than the Morton start."""
    geom = geometry("cylinder_long")
    types = collision_types(geom)
    rank, q = P.partition_sites(geom, types, 19, nranks=4, initial="rcb")
    _, qm = P.partition_sites(geom, types, 19, nranks=4, initial="morton")
    assert q["sites"]["edge_cut"] < 0.6 * qm["sites"]["edge_cut"]
    doms = build_domains(geom, 19, rank, 4)
    assert max(d.procs.shape[0] for d in doms) <= 2


@pytest.mark.parametrize("geom_name", ["tree", "sac", "cylinder_long"])
def test_library_partitioner_makes_the_same_moves(geom_name):
    """hlb_part_bisect / hlb_part_refine_kway (csrc/partition.cu, what a HemeLB build calls in place of
    ParMETIS_V3_PartKway) against the numpy statement of the algorithm: identical partitions from four
    different starts, two weight tables, 2..8 parts; the reported edge cut is the graph's."""
    geom = geometry(geom_name)
    types = collision_types(geom)
    xadj, adjncy = P.site_graph(geom, 19)
    for wall, arch in (("BFL", "B200"), ("GZS", "AMDBULLDOZER")):
        vw = P.site_weights(wall, "NASH", "NASH", arch)[types]
        for R in (2, 3, 5, 8):
            rcb = P.coordinate_bisection(geom.coords, vw, R)
            assert np.array_equal(rcb, P.coordinate_bisection_native(geom.coords, vw, R))
            rib = P.coordinate_bisection_native(geom.coords, vw, R, inertial=True)
            assert np.bincount(rib, minlength=R).min() > 0
            if geom_name != "cylinder_long":  # (a round tube's two short principal axes are degenerate)
                assert np.array_equal(rib, P.coordinate_bisection(geom.coords, vw, R, inertial=True))
            blocks, _ = P.partition_geometry(geom, types, wall, nranks=R, architecture=arch)
            for first in (blocks, rcb, rib, G.basic_decomposition(geom, R)):
                want = P.refine_sites(xadj, adjncy, vw, first, R)
                got, cut = P.refine_sites_native(xadj, adjncy, vw, first, R)
                assert np.array_equal(got, want)
                assert cut == P.site_cut(xadj, adjncy, want)
    start = "rcb" if geom_name == "cylinder_long" else "best"
    a, qa = P.partition_sites(geom, types, 19, nranks=4, native=True, initial=start)
    b, qb = P.partition_sites(geom, types, 19, nranks=4, native=False, initial=start)
    assert np.array_equal(a, b) and qa == qb


def test_library_partitioner_argument_checks():
    from hemelb_b200.capi import lib
    xadj = np.array([0, 1, 2], np.int64)
    adj = np.array([1, 0], np.int64)
    w = np.ones(2)
    with pytest.raises(RuntimeError, match="initial part outside"):
        P.refine_sites_native(xadj, adj, w, np.array([0, 5], np.int32), 2)
    part, cut = P.refine_sites_native(xadj, adj, w, np.array([0, 1], np.int32), 2)
    assert part.tolist() == [0, 1] and cut == 1
    part, cut = P.refine_sites_native(xadj, adj, w, np.array([0, 0], np.int32), 1)
    assert part.tolist() == [0, 0] and cut == 0
    assert P.coordinate_bisection_native(np.zeros((0, 3), np.int64), np.zeros(0), 3).size == 0
    assert lib().hlb_part_bisect(2, None, None, 2, 0, None) != 0
