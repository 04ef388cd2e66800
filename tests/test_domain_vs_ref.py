"""The index tables of the hot path against the reference's OWN geometry::Domain: Code/geometry/Domain.cc,
LookupTree.cc (octree + DistributedStore over one-sided windows), decomposition/BasicDecomposition.cc and the
net:: classes they call, compiled unmodified into oracle/_ref/libhemelb_refdom.so and run over R emulated
ranks (oracle/ref_domain_driver.cc, oracle/fake_mpi.cc).  Every table bit for bit: site order and
per-type counts, neighbourIndices, masks / types / iolet ids, cut distances, wall normals, global
coordinates, neighbouring processors with their shared-distribution slices, streaming indices of
received distributions; and the block -> rank map of BasicDecomposition.

The golden tables under tests/golden/domain_tables_*.npz were written from the same library by
tests/golden/make_golden_domain_tables.py, so that a box without the reference still compares the
oracle and the product's builders with reference output."""
import os

import numpy as np
import pytest

import oracle as O
from hemelb_b200 import geometry as G
from hemelb_b200.domain import build_domains
from tests.cases import geometry
from tests.ref_inputs import load

HERE = os.path.dirname(os.path.abspath(__file__))
needs_ref = pytest.mark.skipif(O.ref_domain_lib() is None, reason="oracle/_ref/libhemelb_refdom.so not built")
KEYS = ["counts", "neighbourIndices", "wallMask", "ioletMask", "siteType", "ioletId", "distanceToWall", "wallNormal",
        "globalCoords", "streamingIndices", "procs"]


def block_rank_of_sites(geom, block_rank):
    B = geom.block_size
    c = geom.coords.astype(np.int64) // B
    return block_rank[(c[:, 0] * geom.block_dims[1] + c[:, 1]) * geom.block_dims[2] + c[:, 2]]


def same(a, b, what):
    assert a["N"] == b["N"] and a["totalSharedFs"] == b["totalSharedFs"], what
    for k in KEYS:
        va, vb = np.asarray(a[k]), np.asarray(b[k])
        assert va.shape == vb.shape and np.array_equal(va, vb), (what, k)


def decomposition(geom, R, kind):
    if R == 1:
        return None
    if kind == "slab":
        return G.slab_decomposition(geom, R)
    if kind == "basic":
        return G.basic_decomposition(geom, R)
    # site-granular and ragged: what an optimised (ParMETIS) decomposition hands to Domain
    rng = np.random.default_rng(7)
    rank = G.slab_decomposition(geom, R).copy()
    flip = rng.random(rank.size) < 0.05
    rank[flip] = rng.integers(0, R, int(flip.sum()))
    return rank


@needs_ref
@pytest.mark.parametrize("name", ["four_cube", "cylinder", "tree", "sac"])
@pytest.mark.parametrize("Q", (15, 19, 27))
@pytest.mark.parametrize("R,kind", [(1, None), (2, "slab"), (3, "slab"), (4, "basic"), (5, "ragged"), (8, "basic"), (16, "ragged")])
def test_oracle_and_product_tables_equal_the_reference_domain(name, Q, R, kind):
    geom = geometry(name)
    if name == "four_cube" and R > 1:
        with pytest.raises(ValueError):  # one block: BasicDecomposition.cc:66-67 throws
            O.RefDomains(geom, Q, None, R)
        return
    rank = decomposition(geom, R, kind)
    try:
        ref = O.RefDomains(geom, Q, rank, R)
    except ValueError:
        pytest.skip("fewer non-empty blocks than ranks: the reference refuses (BasicDecomposition.cc:66-67)")
    orc = O.OracleDomains(geom, Q, rank, R)
    mine = build_domains(geom, Q, rank, R)
    for r in range(R):
        t = ref.tables(r)
        same(t, orc.tables(r), "oracle %s Q%d rank %d/%d" % (name, Q, r, R))
        same(t, mine[r].tables(), "hemelb_b200.domain %s Q%d rank %d/%d" % (name, Q, r, R))


@needs_ref
@pytest.mark.parametrize("name", ["cylinder", "tree", "sac", "cylinder_long"])
@pytest.mark.parametrize("R", (2, 3, 5, 8))
def test_basic_decomposition_equals_the_reference(name, R):
    """BasicDecomposition::Decompose over the reference's own octree (build_block_tree), called with the
    fluid-site count of every block: the block -> rank map that geometry.basic_decomposition restates."""
    geom = geometry(name)
    ref = O.RefDomains(geom, 19, None, R)
    assert np.array_equal(block_rank_of_sites(geom, ref.block_rank), G.basic_decomposition(geom, R))
    solid = ref.block_rank == -(1 << 31)  # SITE_OR_BLOCK_SOLID (Code/constants.h:48)
    assert ref.block_rank[~solid].min() == 0 and ref.block_rank[~solid].max() == R - 1


@needs_ref
@pytest.mark.parametrize("name", ["four_cube", "large_cylinder", "fedosov1c", "cyl_l100_r5"])
def test_reference_fixtures_through_the_reference_domain(name):
    """The reference's own .gmy fixtures, three ranks by its own BasicDecomposition."""
    geom = load(name)[0]
    R = 1 if name == "four_cube" else 3
    ref = O.RefDomains(geom, 19, None, R)
    rank = None if R == 1 else block_rank_of_sites(geom, ref.block_rank)
    orc = O.OracleDomains(geom, 19, rank, R)
    mine = build_domains(geom, 19, rank, R)
    for r in range(R):
        same(ref.tables(r), orc.tables(r), "oracle %s rank %d" % (name, r))
        same(ref.tables(r), mine[r].tables(), "product %s rank %d" % (name, r))


GOLDEN = [("four_cube", 15, 1, None), ("cylinder", 19, 3, "slab"), ("tree", 19, 4, "basic"), ("sac", 27, 2, "slab")]


@pytest.mark.parametrize("name,Q,R,kind", GOLDEN)
def test_golden_reference_tables(name, Q, R, kind):
    """Committed output of the reference's Domain (no reference needed to run this)."""
    geom = geometry(name)
    rank = decomposition(geom, R, kind)
    gold = np.load(os.path.join(HERE, "golden", "domain_tables_%s_q%d_r%d.npz" % (name, Q, R)))
    orc = O.OracleDomains(geom, Q, rank, R)
    mine = build_domains(geom, Q, rank, R)
    for r in range(R):
        g = {k: gold["r%d_%s" % (r, k)] for k in KEYS}
        g["N"], g["totalSharedFs"] = int(gold["r%d_N" % r]), int(gold["r%d_totalSharedFs" % r])
        same(g, orc.tables(r), "oracle vs golden %s rank %d" % (name, r))
        same(g, mine[r].tables(), "product vs golden %s rank %d" % (name, r))
