#!/usr/bin/env python
"""configs[2] / configs[3] / configs[4]: the synthetic bifurcating vascular tree (or, --geometry sac, the
rough-walled aneurysm-like sac), built and partitioned on the GPUs.

  python bench_tree.py [--sites 1.2e8] [--generations 6] [--steps 50] [--kernel LBGK --wall BFL]
  python bench_tree.py --kernel MRT --wall GZS --inlet LADD                       # configs[3]
  python bench_tree.py --geometry sac --lattice 27 --kernel TRT --wall BFL        # configs[4]
  python -m torch.distributed.run --nproc-per-node N ... bench_tree.py --sites 1e9

Strong target size: ``--sites`` is the WHOLE tree (default 1.2e8 per GPU x N).  Pipeline, all on the
devices (hlb_dom_*): per-block fluid counts of the analytic capsule tree (each rank counts an
x-slab of blocks, all-gathered) -> the reference's BasicDecomposition over Morton-ordered blocks
(Code/geometry/decomposition/BasicDecomposition.cc:21-96) -> each rank voxelises its own blocks
plus one voxel of rim and builds its Domain tables -> engine handles device-to-device -> NCCL halo.
One JSON line: MLUPS (max over ranks, CUDA events), roofline fractions at 380 B/site (D3Q19),
set-up seconds, sites per rank, halo sizes.  A secondary measurement: the contract line is bench.py.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def tree_for_sites(total_sites: float, generations: int):
    """Root radius / length so that the tree holds about ``total_sites`` fluid sites: every
    generation carries ~pi r^2 L (Murray's law with L shrinking by 0.8), L = 4.17 r."""
    per_gen = total_sites / generations
    r = (per_gen / (np.pi * 4.17)) ** (1.0 / 3.0)
    return float(r), float(4.17 * r)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sites", type=float, default=0.0)
    ap.add_argument("--generations", type=int, default=6)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--kernel", default="LBGK")
    ap.add_argument("--wall", default="BFL")
    ap.add_argument("--inlet", default="NASH")
    ap.add_argument("--outlet", default="NASH")
    ap.add_argument("--lattice", type=int, default=19)
    ap.add_argument("--geometry", default="tree", choices=["tree", "sac"])
    ap.add_argument("--roughness", type=float, default=3.0, help="sac: wall roughness amplitude in voxels")
    ap.add_argument("--decomposition", default="basic", choices=["basic", "weighted"],
                    help="basic: the reference's BasicDecomposition over Morton blocks; weighted: the METIS-free "
                         "weighted k-way block partition (hemelb_b200/partition.py)")
    ap.add_argument("--partition-start", default="morton", choices=["morton", "rcb", "inertial", "best"],
                    help="weighted only: start from the bisection of the Morton-ordered blocks (as measured in "
                         "profiles/), from recursive coordinate / inertial bisection, or from the best of the three")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    Q = args.lattice
    total = args.sites or 1.2e8 * world

    from hemelb_b200.capi import iolet_record
    from hemelb_b200.devdomain import DeviceDomain, basic_decomposition_of_counts, sac_shape, tree_shape
    from hemelb_b200.lbm import GpuLBM, prepare_boundary_objects

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    t0 = time.time()
    rough = None
    if args.geometry == "sac":
        # sphere of radius R (4/3 pi R^3 sites) crossed by a neck of radius R/4 reaching R/3 beyond it
        r0 = float((total / (4.0 / 3.0 * np.pi)) ** (1.0 / 3.0))
        l0 = r0 / 3.0
        caps, iolets, shape, rough = sac_shape(r0, r0 / 4.0, l0, roughness=args.roughness)
    else:
        r0, l0 = tree_for_sites(total, args.generations)
        caps, iolets, shape = tree_shape(args.generations, r0, l0)
    dom = DeviceDomain.from_shape(caps, iolets, shape, Q, 8, None, rank, world, local_rank, build=False, roughness=rough)
    bd = dom.block_dims
    # per-block fluid counts: each rank counts an x-slab of blocks
    xs = [int(bd[0]) * r // world for r in range(world + 1)]
    weighted = args.decomposition == "weighted" and world > 1
    if weighted:
        mine = np.stack(dom.count_block_sites_typed([xs[rank], 0, 0], [xs[rank + 1], int(bd[1]), int(bd[2])]))
    else:
        mine = dom.count_block_sites([xs[rank], 0, 0], [xs[rank + 1], int(bd[1]), int(bd[2])])
    if dist is not None:
        import torch
        parts = [None] * world
        dist.all_gather_object(parts, mine)
        counts = np.concatenate(parts, 1 if weighted else 0)
    else:
        counts = mine
    boundary_counts = None
    if weighted:
        counts, boundary_counts = counts[0], counts[1]
    t_count = time.time() - t0
    n_global = int(counts.sum())
    if world > 1:
        if weighted:
            from hemelb_b200.devdomain import weighted_decomposition_of_counts
            rob = weighted_decomposition_of_counts(counts, boundary_counts, world, args.wall,
                                                   initial=args.partition_start)
        else:
            rob = basic_decomposition_of_counts(counts, world)
        dom.set_partition(("blocks", rob))
    t1 = time.time()
    dom.build()
    t_build = time.time() - t1

    def rec(p, bc, inlet):
        if bc == "LADD":
            return iolet_record(1, tuple(p.normal), tuple(p.position), radius=p.radius - 1.0, max_speed=0.01)
        return iolet_record(0, tuple(p.normal), tuple(p.position), radius=p.radius,
                            density_mean=1.0005 if inlet else 0.9995, density_amp=0.0, period=1000.0)
    ins = [rec(p, args.inlet, True) for p in dom.meta["inlets"]]
    outs = [rec(p, args.outlet, False) for p in dom.meta["outlets"]]
    prepare_boundary_objects(ins, outs)
    t2 = time.time()
    def all_gather(obj):
        if dist is None:
            return [obj]
        parts = [None] * world
        dist.all_gather_object(parts, obj)
        return parts
    gpu = GpuLBM.from_device_domain(dom, args.kernel, args.wall, args.inlet, args.outlet, tau=0.8, inlets=ins, outlets=outs,
                                    all_gather=all_gather)
    if world > 1:
        uid = [GpuLBM.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        gpu.comm_init(uid[0])
    gpu.set_equilibrium(1.0, (0.0, 0.0, 0.0))
    t_engine = time.time() - t2
    setup = time.time() - t0

    def barrier():
        gpu.sync()
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    gpu.step(max(args.warmup, 3))
    barrier()
    ms, fused_ms, fused_sites = gpu.time_steps_detail(args.steps)  # product schedule: fused mid-domain kernel
    barrier()
    gpu.set_overlap(False)           # second region, kernels back to back: the bulk kernel's own duration
    gpu.step(2)
    barrier()
    serial_ms, bulk_ms, bulk_sites = gpu.time_steps_detail(args.steps)
    gpu.set_overlap(True)
    barrier()
    mon = gpu.monitor()
    per_rank = [dom.N]
    halo = [int(dom.totalSharedFs)]
    nbrs = [int(dom.procs.shape[0])]
    if dist is not None:
        import torch
        t = torch.tensor([ms, serial_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, serial_ms = float(t[0].item()), float(t[1].item())
        g = [None] * world
        dist.all_gather_object(g, (dom.N, int(dom.totalSharedFs), int(dom.procs.shape[0]), float(mon["min_f"])))
        per_rank, halo, nbrs = [x[0] for x in g], [x[1] for x in g], [x[2] for x in g]
        mon["min_f"] = min(x[3] for x in g)
    if rank != 0:
        return 0
    peak = 6650.0
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    B = 20 * Q
    mlups = sum(per_rank) * args.steps / (ms * 1e-3) / 1e6
    nb = int(dom.N - dom.mid[0] - dom.edge[0])
    what = ("tree: %d generations, root r=%.1f l=%.1f" % (args.generations, r0, l0) if args.geometry == "tree" else
            "sac: rough sphere r=%.1f (value noise +-%.1f voxels) + neck r=%.1f" % (r0, args.roughness, r0 / 4.0))
    wall_links = None
    line = {"config": "%s, D3Q%d %s + %s walls, %s inlet / %s outlets" % (what, Q, args.kernel, args.wall, args.inlet, args.outlet),
            "gzs_remote_links_rank0": int(gpu.gzs_need.shape[0]),
            "rank0_target_runs": list(gpu.target_runs()),
            "n_gpus": world, "sites": sum(per_rank), "sites_counted": n_global, "sites_per_rank": per_rank,
            "halo_doubles_per_rank": halo, "neighbours_per_rank": nbrs,
            "decomposition": ("weighted k-way over blocks, %s start (hemelb_b200/partition.py)" % args.partition_start
                              if args.decomposition == "weighted"
                              else "BasicDecomposition over Morton-ordered blocks") if world > 1 else "single rank",
            "MLUPS": mlups, "ms_per_step": ms / args.steps, "bytes_per_site": B,
            "whole_step_frac_of_hbm_roofline": mlups * 1e6 * B / 1e9 / peak / world,
            "rank0_bulk_kernel_frac": (fused_sites * B / 1e9 / (fused_ms * 1e-3)) / peak if fused_ms else None,
            "rank0_bulk_kernel_frac_plain_order": (bulk_sites * B / 1e9 / (bulk_ms * 1e-3)) / peak if bulk_ms else None,
            "fused_mid_kernel": bool(fused_sites > bulk_sites),
            
            "rank0_bulk_share_of_step": fused_ms / ms if ms else None, "serial_ms_per_step": serial_ms / args.steps,
            "rank0_boundary_fraction": nb / max(dom.N, 1), "n_outlets": len(outs),
            "lattice_blocks": [int(x) for x in bd],
            "setup_seconds": {"total": setup, "count_blocks": t_count, "domain_build": t_build,
                              "domain_build_kernels": dom.build_seconds, "engine": t_engine},
            "stable": bool(mon["min_f"] > 0), "monitor": mon, "peak_GBps": peak}
    print(json.dumps(line), flush=True)
    gpu.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
