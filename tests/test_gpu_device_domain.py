"""Device-side geometry::Domain construction (hlb_dom_*): every table bit-identical to the host
builder (itself pinned bit-for-bit to the oracle's literal restatement of Code/geometry/Domain.cc
in tests/test_domain_tables.py) and to the oracle directly; the device voxeliser against the host
one; and the engine fed device-to-device against the oracle."""
import numpy as np
import pytest

import oracle as O
from hemelb_b200 import geometry as G
from hemelb_b200.devdomain import (DeviceDomain, basic_decomposition_of_counts, cylinder_shape, tree_shape)
from hemelb_b200.domain import build_domains
from hemelb_b200.lbm import GpuLBM
from tests.cases import anisotropic_f, geometry, iolets_for

pytestmark = pytest.mark.gpu

KEYS = ("N", "totalSharedFs", "counts", "neighbourIndices", "wallMask", "ioletMask", "ioletId", "distanceToWall",
        "globalCoords", "streamingIndices", "procs")


def _same(dev: DeviceDomain, host, what):
    a, b = dev.tables(), host.tables()
    for k in KEYS:
        va, vb = np.asarray(a[k]), np.asarray(b[k])
        assert va.shape == vb.shape and np.array_equal(va, vb), (what, k)
    # wall normals: the engine only reads them for boundary-typed sites
    bs = dev.boundary_sites()
    na, nb = a["wallNormal"].reshape(-1, 3)[bs], b["wallNormal"].reshape(-1, 3)[bs]
    assert np.array_equal(na, nb), (what, "wallNormal")


@pytest.mark.parametrize("name", ["four_cube", "cylinder", "tree", "sac"])
@pytest.mark.parametrize("Q", (15, 19, 27))
def test_explicit_source_tables_bit_exact(name, Q):
    geom = geometry(name)
    cases = [(None, 1), (G.slab_decomposition(geom, 2, axis=2), 2), (G.slab_decomposition(geom, 3, axis=0), 3)]
    if name != "four_cube":
        cases.append((G.basic_decomposition(geom, 5), 5))
    if name in ("tree", "sac") and Q == 19:
        # a ParMETIS-like site partition (cuts through blocks): hemelb_b200/partition.py step 4
        from hemelb_b200.partition import partition_sites
        from tests.test_partition import collision_types
        cases.append((partition_sites(geom, collision_types(geom, Q), Q, nranks=4, initial="morton")[0], 4))
    for ros, R in cases:
        host = build_domains(geom, Q, ros, R)
        for r in range(R):
            dev = DeviceDomain.from_geometry(geom, Q, ros, r, R)
            _same(dev, host[r], (name, Q, R, r))
            assert np.array_equal(dev.input_index(), host[r].inputIndex)
            dev.close()


def test_explicit_source_against_oracle_directly():
    geom, Q, R = geometry("tree"), 19, 4
    ros = G.basic_decomposition(geom, R)
    od = O.OracleDomains(geom, Q, ros, R)
    for r in range(R):
        dev = DeviceDomain.from_geometry(geom, Q, ros, r, R)
        a, b = od.tables(r), dev.tables()
        for k in ("neighbourIndices", "streamingIndices", "globalCoords", "wallMask", "ioletMask", "ioletId",
                  "distanceToWall"):
            assert np.array_equal(np.asarray(a[k]).reshape(-1), np.asarray(b[k]).reshape(-1)), (r, k)
        assert np.array_equal(np.asarray(a["counts"]), b["counts"])
        dev.close()


def test_explicit_source_halo_only_upload():
    """A rank only needs its own sites and their lattice neighbours (what cylinder_slab ships)."""
    R, Q = 3, 19
    full = G.cylinder_extruded(6.3, 60)
    per = 60 // R
    ros_full = np.minimum((full.coords[:, 2].astype(np.int64) - 2) // per, R - 1).astype(np.int32)
    host = build_domains(full, Q, ros_full, R)
    for r in range(R):
        sub, ros = G.cylinder_slab(6.3, 60, R, r)
        dev = DeviceDomain.from_geometry(sub, Q, ros, r, R)
        _same(dev, host[r], ("slab", r))
        dev.close()


def _shape(name):
    if name == "cylinder":
        return cylinder_shape(5.3, 20)
    return tree_shape(3, 5.0, 16.0)


@pytest.mark.parametrize("name", ["cylinder", "tree"])
def test_device_voxeliser_matches_host_voxeliser(name):
    caps, iolets, shape = _shape(name)
    dev = DeviceDomain.from_shape(caps, iolets, shape, 19)
    g = dev.geometry()
    h = geometry(name)
    assert np.array_equal(g.block_dims, h.block_dims)
    assert np.array_equal(g.coords, h.coords)  # same fluid sites, .gmy order
    assert np.array_equal(g.bsite, h.bsite)
    assert np.array_equal(g.btype, h.btype) and np.array_equal(g.biolet, h.biolet)
    assert np.array_equal(g.bnavail, h.bnavail)
    # cut distances: both bisect to 2^-30, float32-rounded; the host's phi sums in another order
    assert np.abs(g.bdist.astype(np.float64) - h.bdist.astype(np.float64)).max() <= 2e-6
    # normals: exact radial direction on the device; the host tree uses a finite difference of phi
    # (so they differ where two capsules meet: compare away from the junction kinks)
    dn = np.abs(g.bnormal.astype(np.float64) - h.bnormal.astype(np.float64)).max(1)
    assert dn.max() <= 1e-6 if name == "cylinder" else np.quantile(dn, 0.8) <= 2e-2


def test_device_voxeliser_rough_sac_matches_host_voxeliser():
    """configs[4]: the rough-walled sac (sphere displaced by seeded value noise + neck) voxelised on the
    device against geometry.sac -- same fluid sites, same cut links, cut distances and the
    finite-difference wall normals to float32 rounding."""
    from hemelb_b200.devdomain import sac_shape
    caps, iolets, shape, rough = sac_shape(8, 3, 4, roughness=1.5)
    dev = DeviceDomain.from_shape(caps, iolets, shape, 27, roughness=rough)
    g = dev.geometry()
    h = geometry("sac")
    assert np.array_equal(g.block_dims, h.block_dims)
    assert np.array_equal(g.coords, h.coords)
    assert np.array_equal(g.bsite, h.bsite)
    assert np.array_equal(g.btype, h.btype) and np.array_equal(g.biolet, h.biolet)
    assert np.array_equal(g.bnavail, h.bnavail)
    assert np.abs(g.bdist.astype(np.float64) - h.bdist.astype(np.float64)).max() <= 2e-6
    assert np.abs(g.bnormal.astype(np.float64) - h.bnormal.astype(np.float64)).max() <= 1e-5
    # and the tables the engine consumes, against the host builder on the host-voxelised geometry
    want = build_domains(h, 27)[0].tables()
    got = dev.tables()
    for key in ("N", "totalSharedFs"):
        assert got[key] == want[key]
    for key in ("counts", "neighbourIndices", "wallMask", "ioletMask", "ioletId", "globalCoords"):
        assert np.array_equal(got[key], want[key]), key


@pytest.mark.parametrize("name", ["cylinder", "tree"])
@pytest.mark.parametrize("Q", (15, 19, 27))
def test_analytic_source_tables_bit_exact(name, Q):
    """Tables built straight from the shape == host builder on the downloaded voxelisation."""
    caps, iolets, shape = _shape(name)
    dev = DeviceDomain.from_shape(caps, iolets, shape, Q)
    g = dev.geometry()
    _same(dev, build_domains(g, Q)[0], (name, Q, "single"))
    # slabs along z through blocks
    R = 3
    z = g.coords[:, 2].astype(np.int64)
    cuts = [int(np.quantile(z, k / R)) for k in range(1, R)]
    first = np.array([np.iinfo(np.int64).min // 2] + cuts + [np.iinfo(np.int64).max // 2], np.int64)
    ros = (np.searchsorted(first, z, side="right") - 1).astype(np.int32)
    host = build_domains(g, Q, ros, R)
    for r in range(R):
        d = DeviceDomain.from_shape(caps, iolets, shape, Q, partition=("slabs", 2, first), rank=r, nranks=R)
        _same(d, host[r], (name, Q, "slabs", r))
        d.close()
    # whole blocks by BasicDecomposition of the device's own per-block counts
    R = 4
    counts = dev.count_block_sites()
    assert counts.sum() == g.n_sites
    rob = basic_decomposition_of_counts(counts, R)
    ros = G.basic_decomposition(g, R)
    bd = g.block_dims.astype(np.int64)
    bc = (g.coords // g.block_size).astype(np.int64)
    assert np.array_equal(rob[(bc[:, 0] * bd[1] + bc[:, 1]) * bd[2] + bc[:, 2]], ros)
    host = build_domains(g, Q, ros, R)
    for r in range(R):
        d = DeviceDomain.from_shape(caps, iolets, shape, Q, partition=("blocks", rob), rank=r, nranks=R)
        _same(d, host[r], (name, Q, "blocks", r))
        d.close()
    dev.close()


@pytest.mark.parametrize("source", ["explicit", "analytic"])
@pytest.mark.parametrize("Q,kernel,wall,inlet,outlet", [
    (19, "LBGK", "BFL", "NASH", "NASH"), (19, "MRT", "GZS", "LADD", "NASH"), (27, "TRT", "BFL", "NASH", "NASH"),
    (15, "LBGK", "SBB", "LADD", "LADD")])
def test_engine_from_device_domain_matches_oracle(source, Q, kernel, wall, inlet, outlet):
    if source == "explicit":
        geom = geometry("tree")
        dev = DeviceDomain.from_geometry(geom, Q)
    else:
        caps, iolets, shape = tree_shape(3, 5.0, 16.0)
        dev = DeviceDomain.from_shape(caps, iolets, shape, Q)
        geom = dev.geometry()
    inlets, outlets = iolets_for(geom, inlet, outlet)
    gpu = GpuLBM.from_device_domain(dev, kernel, wall, inlet, outlet, tau=0.8, inlets=inlets, outlets=outlets)
    sim = O.OracleSim(O.OracleDomains(geom, Q), kernel, wall, inlet, outlet, tau=0.8, inlets=inlets, outlets=outlets)
    # the engine's table read-back (reference form, through the internal renumbering)
    assert np.array_equal(gpu.get_neighbour_indices(), build_domains(geom, Q)[0].neighbour_indices())
    f0 = anisotropic_f(dev.N, Q, 0)
    gpu.set_f(f0)
    sim.set_f(f0)
    gpu.set_cache_mask(255)
    sim.set_cache_mask(255)
    gpu.step(10)
    sim.step(10)
    assert np.array_equal(gpu.get_f()[:dev.N * Q], sim.get_f()[:dev.N * Q])
    for name in ("density", "velocity", "wall_shear_stress", "traction"):
        assert np.array_equal(gpu.get_cache(name), sim.get_cache(name)), name


def test_two_rank_engine_from_device_domain_host_staged_halo():
    """Two ranks built on the device from the shape, halo moved through the host: identical to the
    oracle's two emulated ranks."""
    Q, R = 19, 2
    caps, iolets, shape = cylinder_shape(5.3, 20)
    whole = DeviceDomain.from_shape(caps, iolets, shape, Q)
    geom = whole.geometry()
    cut = 2 + 10
    first = np.array([-2**60, cut, 2**60], np.int64)
    ros = (geom.coords[:, 2] >= cut).astype(np.int32)
    inlets, outlets = iolets_for(geom, "NASH", "NASH")
    doms = [DeviceDomain.from_shape(caps, iolets, shape, Q, partition=("slabs", 2, first), rank=r, nranks=R)
            for r in range(R)]
    gpus = [GpuLBM.from_device_domain(d, "LBGK", "BFL", tau=0.8, inlets=inlets, outlets=outlets) for d in doms]
    sim = O.OracleSim(O.OracleDomains(geom, Q, ros, R), "LBGK", "BFL", tau=0.8, inlets=inlets, outlets=outlets)
    for r, (d, g) in enumerate(zip(doms, gpus)):
        f0 = anisotropic_f(d.N, Q, d.totalSharedFs, site_offset=1000 * r)
        g.set_f(f0)
        sim.set_f(f0, r)
    for _ in range(6):
        for g in gpus:
            g.request_comms()
            g.pre_send()
            g.pre_receive()
        halos = [g.get_halo(1) for g in gpus]
        # slice k of rank a towards b pairs with b's slice towards a (same offsets both sides)
        for a, (d, g) in enumerate(zip(doms, gpus)):
            recv = np.zeros(d.totalSharedFs)
            for (p, cnt, fst) in d.procs:
                o = int(fst) - (d.N * Q + 1)
                back = doms[p].procs
                j = int(np.nonzero(back[:, 0] == a)[0][0])
                po = int(back[j, 2]) - (doms[p].N * Q + 1)
                recv[o:o + cnt] = halos[p][po:po + cnt]
            g.set_halo(recv, 0)
        for g in gpus:
            g.post_receive()
            g.swap_old_and_new()
            g.state.increment()
        sim.step(1)
    for r, (d, g) in enumerate(zip(doms, gpus)):
        assert np.array_equal(g.get_f()[:d.N * Q], sim.get_f(r)[:d.N * Q]), r


@pytest.mark.parametrize("kernel,inlet", [("LBGK", "LADD"), ("MRT", "NASH")])
def test_gzs_site_halo_from_device_domains(kernel, inlet):
    """GuoZhengShi walls over three ranks built on the device: the remote needs come from
    hlb_dom_gzs_needs, the owners' serve lists from hlb_dom_lookup_sites of the coordinates asked for, the
    site halo (whole f_old rows) and the distribution halo move through the host: identical to the
    oracle's three emulated ranks."""
    Q, R = 19, 3
    caps, iolets, shape = cylinder_shape(4.2, 44)
    geom = DeviceDomain.from_shape(caps, iolets, shape, Q).geometry()
    cuts = [2 + 15, 2 + 30]
    first = np.array([-2**60] + cuts + [2**60], np.int64)
    ros = np.searchsorted(np.array(cuts), geom.coords[:, 2], side="right").astype(np.int32)
    inlets, outlets = iolets_for(geom, inlet, "NASH")
    doms = [DeviceDomain.from_shape(caps, iolets, shape, Q, partition=("slabs", 2, first), rank=r, nranks=R)
            for r in range(R)]
    from hemelb_b200.lbm import gzs_device_domain_needs
    asks = [gzs_device_domain_needs(d)[1] for d in doms]
    assert sum(len(a) for a in asks) > 0
    gpus = [GpuLBM.from_device_domain(d, kernel, "GZS", inlet, "NASH", tau=0.8, inlets=inlets, outlets=outlets,
                                      all_gather=lambda obj: asks) for d in doms]
    sim = O.OracleSim(O.OracleDomains(geom, Q, ros, R), kernel, "GZS", inlet, "NASH", tau=0.8, inlets=inlets, outlets=outlets)
    for r, (d, g) in enumerate(zip(doms, gpus)):
        f0 = anisotropic_f(d.N, Q, d.totalSharedFs, site_offset=5 * r)
        g.set_f(f0)
        sim.set_f(f0, r)
    for _ in range(5):
        for g in gpus:
            g.exchange_site_halo()
        sends = [g.get_gzs_send() for g in gpus]
        for r, g in enumerate(gpus):
            rows = np.zeros((g.gzs_row_owner.size, Q))
            for p in range(R):
                mine = np.nonzero(g.gzs_row_owner == p)[0]
                theirs = np.nonzero(gpus[p].gzs_serve[:, 0] == r)[0]
                assert mine.size == theirs.size
                rows[mine] = sends[p][theirs]
            g.set_gzs_ghost(rows)
        for g in gpus:
            g.request_comms()
            g.pre_send()
            g.pre_receive()
        halos = [g.get_halo(1) for g in gpus]
        for a, (d, g) in enumerate(zip(doms, gpus)):
            recv = np.zeros(d.totalSharedFs)
            for (p, cnt, fst) in d.procs:
                o = int(fst) - (d.N * Q + 1)
                back = doms[p].procs
                j = int(np.nonzero(back[:, 0] == a)[0][0])
                po = int(back[j, 2]) - (doms[p].N * Q + 1)
                recv[o:o + cnt] = halos[p][po:po + cnt]
            g.set_halo(recv, 0)
        for g in gpus:
            g.post_receive()
            g.swap_old_and_new()
            g.state.increment()
        sim.step(1)
    for r, (d, g) in enumerate(zip(doms, gpus)):
        if kernel == "LBGK":
            assert np.array_equal(g.get_f()[:d.N * Q], sim.get_f(r)[:d.N * Q]), r
        else:
            assert np.abs(g.get_f()[:d.N * Q] - sim.get_f(r)[:d.N * Q]).max() <= 1e-13, r


@pytest.mark.parametrize("name", ["cylinder", "tree"])
@pytest.mark.parametrize("Q", (15, 19, 27))
def test_typed_block_counts_and_weighted_decomposition(name, Q):
    """Per-block boundary-typed site counts (the vertex weights of the weighted decomposition) equal
    the counts derived from the built tables; the weighted k-way block partition made from them is a
    valid site -> rank rule: every rank's tables equal the host builder's."""
    from hemelb_b200.devdomain import weighted_decomposition_of_counts
    caps, iolets, shape = _shape(name)
    dev = DeviceDomain.from_shape(caps, iolets, shape, Q)
    g = dev.geometry()
    counts, boundary = dev.count_block_sites_typed()
    assert np.array_equal(counts, dev.count_block_sites())
    # reference: boundary-typed sites of the single-rank tables, binned by block
    t = dev.tables()
    typed = (np.asarray(t["wallMask"]) != 0) | (np.asarray(t["ioletMask"]) != 0)
    bc = (np.asarray(t["globalCoords"]).reshape(-1, 3) // g.block_size).astype(np.int64)
    want = np.zeros(counts.shape, np.int64)
    np.add.at(want, (bc[typed, 0], bc[typed, 1], bc[typed, 2]), 1)
    assert np.array_equal(boundary, want)
    R = 4
    rob = weighted_decomposition_of_counts(counts, boundary, R, "BFL")
    assert set(np.unique(rob[rob >= 0]).tolist()) == set(range(R))
    bd = g.block_dims.astype(np.int64)
    gc = (g.coords // g.block_size).astype(np.int64)
    ros = rob[(gc[:, 0] * bd[1] + gc[:, 1]) * bd[2] + gc[:, 2]].astype(np.int32)
    host = build_domains(g, Q, ros, R)
    for r in range(R):
        d = DeviceDomain.from_shape(caps, iolets, shape, Q, partition=("blocks", rob), rank=r, nranks=R)
        _same(d, host[r], (name, Q, "weighted blocks", r))
        d.close()
    dev.close()


@pytest.mark.parametrize("name,Q,R,kind", [("four_cube", 15, 1, None), ("cylinder", 19, 3, "slab"), ("tree", 19, 4, "basic"),
                                           ("sac", 27, 2, "slab")])
def test_device_tables_equal_the_reference_domain(name, Q, R, kind):
    """The device builder (hlb_dom_*) against tables written by the reference's own geometry::Domain
    (tests/golden/domain_tables_*.npz, tests/test_domain_vs_ref.py), and against that library itself where it
    travelled with the snapshot."""
    import os
    from tests.test_domain_vs_ref import KEYS as REF_KEYS, decomposition
    geom = geometry(name)
    rank = decomposition(geom, R, kind)
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "domain_tables_%s_q%d_r%d.npz" % (name, Q, R)))
    live = O.RefDomains(geom, Q, rank, R) if O.ref_domain_lib() is not None else None
    for r in range(R):
        dev = DeviceDomain.from_geometry(geom, Q, rank, r, R)
        t = dev.tables()
        bs = dev.boundary_sites()
        for k in REF_KEYS:
            for src in ([gold["r%d_%s" % (r, k)]] + ([live.tables(r)[k]] if live else [])):
                a, b = np.asarray(t[k]), np.asarray(src)
                if k == "wallNormal":  # (read for boundary-typed sites only)
                    a, b = a.reshape(-1, 3)[bs], b.reshape(-1, 3)[bs]
                if k == "siteType":
                    continue  # the device tables carry the collision type; types are compared through counts
                assert a.shape == b.shape and np.array_equal(a, b), (name, Q, R, r, k)
        assert t["N"] == int(gold["r%d_N" % r]) and t["totalSharedFs"] == int(gold["r%d_totalSharedFs" % r])
        dev.close()
