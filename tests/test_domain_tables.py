"""Index tables: the product's vectorised Domain builder against the oracle's literal restatement
of Code/geometry/Domain.cc -- every table bit-exact -- plus the .gmy reader/writer."""
import os

import numpy as np
import pytest

import oracle as O
from hemelb_b200 import geometry as G
from hemelb_b200.domain import build_domains
from tests.cases import geometry

REF_RES = "/root/reference/Code/tests/resources"


def _same_tables(geom, Q, rank, R):
    od = O.OracleDomains(geom, Q, rank, R)
    mine = build_domains(geom, Q, rank, R)
    for r in range(R):
        a, b = od.tables(r), mine[r].tables()
        for k in a:
            if k == "Q":
                continue
            va, vb = np.asarray(a[k]), np.asarray(b[k])
            assert va.shape == vb.shape and np.array_equal(va, vb), (Q, R, r, k)


@pytest.mark.parametrize("name", ["four_cube", "cylinder", "tree", "sac"])
@pytest.mark.parametrize("Q", (15, 19, 27))
def test_tables_bit_exact(name, Q):
    geom = geometry(name)
    _same_tables(geom, Q, None, 1)
    _same_tables(geom, Q, G.slab_decomposition(geom, 2, axis=2), 2)
    _same_tables(geom, Q, G.slab_decomposition(geom, 3, axis=0), 3)
    if name != "four_cube":
        _same_tables(geom, Q, G.basic_decomposition(geom, 5), 5)


def test_four_cube_site_counts():
    """64 fluid = 8 bulk / 24 wall / 4 inlet / 4 outlet / 12 inlet-wall / 12
    outlet-wall (StreamerTests.cc:235 asserts the 24)."""
    d = build_domains(geometry("four_cube"), 15)[0]
    assert list(d.mid) == [8, 24, 4, 4, 12, 12] and d.edge.sum() == 0 and d.N == 64
    t = d.tables()
    assert (t["distanceToWall"][t["distanceToWall"] >= 0] == 0.5).all()


def test_table_invariants_multi_rank():
    """Halo slot pairing (Domain.cc:530-576): slice k of rank a towards b has the same length as
    b's towards a; every send slot is written by exactly one (site, direction); every received
    distribution lands on a distinct local population."""
    geom, Q, R = geometry("tree"), 19, 4
    doms = build_domains(geom, Q, G.basic_decomposition(geom, R), R)
    for a, d in enumerate(doms):
        idx = d.neighbour_indices()
        send = idx[idx > d.N * Q]
        assert np.array_equal(np.sort(send), np.arange(d.N * Q + 1, d.N * Q + 1 + d.totalSharedFs))
        assert np.unique(d.streamingIndices).size == d.totalSharedFs
        for (p, cnt, first) in d.procs:
            back = doms[p].procs
            j = np.nonzero(back[:, 0] == a)[0]
            assert j.size == 1 and back[j[0], 1] == cnt
        # edge sites are exactly those with a send slot
        has_send = (idx.reshape(d.N, Q) > d.N * Q).any(1)
        assert np.array_equal(np.nonzero(has_send)[0], np.arange(d.mid.sum(), d.N))


def test_chunked_neighbour_indices_equal_whole():
    d = build_domains(geometry("cylinder"), 19)[0]
    whole = d.neighbour_indices()
    parts = np.concatenate([d.neighbour_indices(s, min(500, d.N - s)) for s in range(0, d.N, 500)])
    assert np.array_equal(whole, parts)


def test_extruded_cylinder_equals_voxelised():
    a, b = G.cylinder(7.4, 37), G.cylinder_extruded(7.4, 37)
    for n in ("coords", "bsite", "btype", "biolet", "bdist", "bnavail", "bnormal", "block_dims"):
        assert np.array_equal(getattr(a, n), getattr(b, n)), n


def test_gmy_round_trip(tmp_path):
    for name in ("four_cube", "cylinder"):
        g = geometry(name)
        p = str(tmp_path / (name + ".gmy"))
        G.write_gmy(g, p)
        h = G.read_gmy(p)
        for n in ("coords", "bsite", "btype", "biolet", "bdist", "bnavail", "bnormal", "block_dims"):
            assert np.array_equal(getattr(g, n), getattr(h, n)), (name, n)


def test_basic_decomposition_is_balanced_and_blockwise():
    geom = geometry("tree")
    for R in (2, 4, 8):
        rank = G.basic_decomposition(geom, R)
        counts = np.bincount(rank, minlength=R)
        assert counts.min() > 0 and counts.max() < 2.0 * geom.n_sites / R
        blocks = (geom.coords // geom.block_size).astype(np.int64)
        key = (blocks[:, 0] * 1000 + blocks[:, 1]) * 1000 + blocks[:, 2]
        for k in np.unique(key):
            assert np.unique(rank[key == k]).size == 1


@pytest.mark.skipif(not os.path.exists(REF_RES), reason="reference checkout absent")
def test_reference_four_cube_gmy_is_reproduced_byte_for_byte(tmp_path):
    """GeometryReaderTests.cc:41-76 (four_cube.gmy == the in-memory four-cube fixture): our reader
    decodes the reference's file to our generator's geometry and our writer reproduces the file."""
    ref = G.read_gmy(os.path.join(REF_RES, "four_cube.gmy"))
    gen = G.four_cube()
    for n in ("coords", "bsite", "btype", "biolet", "bdist", "bnavail", "bnormal", "block_dims"):
        assert np.array_equal(getattr(ref, n), getattr(gen, n)), n
    p = str(tmp_path / "fc.gmy")
    G.write_gmy(gen, p)
    assert open(p, "rb").read() == open(os.path.join(REF_RES, "four_cube.gmy"), "rb").read()


@pytest.mark.skipif(not os.path.exists(REF_RES), reason="reference checkout absent")
def test_reference_large_cylinder_octree_counts():
    """LookupTreeTests.cc:213-306: large_cylinder.gmy has 20 non-empty leaf blocks; 5576 fluid
    sites."""
    g = G.read_gmy(os.path.join(REF_RES, "large_cylinder.gmy"))
    assert g.n_sites == 5576
    blocks = np.unique((g.coords // g.block_size), axis=0)
    assert blocks.shape[0] == 20
    _same_tables(g, 15, None, 1)
    _same_tables(g, 15, G.basic_decomposition(g, 4), 4)
