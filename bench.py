#!/usr/bin/env python
"""MLUPS benchmark of the collide-and-stream hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

The workload is BASELINE.json's north-star configuration, configs[2]: the synthetic bifurcating
vascular tree (6 generations, Murray's law), D3Q19 LBGK + BouzidiFirdaousLallemand walls, Nash pressure
iolets, ~1.1e8 fluid sites per GPU (weak scaling: the tree grows with N; 8 GPUs = 8.8e8 sites).  For
N > 1 (torchrun, one rank per GPU) the lattice blocks are partitioned by the reference's own
BasicDecomposition over Morton-ordered blocks (or --decomposition weighted: the METIS-free weighted
k-way partition), each rank voxelises and builds the tables of its own part on its GPU, and the halo
travels over NCCL send/recv.

One JSON line on rank 0:
  value      whole-job MLUPS with everything resident in HBM (CUDA events, max over ranks);
  e2e        the same steps driven phase by phase through the C ABI, the per-step host scalars copied
             H2D and a monitor read back D2H every step;
  roofline   the site kernel over the mid-domain part (the dominant kernel), timed live with CUDA
             events on the engine's stream, against the measured HBM copy bandwidth;
  secondary  configs[1], the straight cylinder (r=146, l=1500 per GPU, z-slabs for N > 1), measured in
             the same run;
  cpu_baseline  the reference's own streamers / kernels (oracle/_ref) on the host cores, a bounded
             sample of the same tree.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import re
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "MLUPS (fluid-site updates/s) D3Q19 LBGK+BFL"
Q = 19
BYTES_PER_SITE = 20 * Q  # 2*Q*8 B distributions + Q*4 B neighbour indices
TAU = 0.8
GENERATIONS = 6
SITE_KERNEL = "collide_stream_kernel<19,LBGK,BFL,NASH,NASH>"
REFERENCE_SAMPLE_SITES = 2.0e7


def kernel_source_hash():
    """Identifies the site kernel a DRAM-traffic capture belongs to: the kernel sources without their
    comments and white space (an edited comment does not make a capture stale, an edited statement does)."""
    h = hashlib.sha256()
    for f in ("kernels.cuh", "lattice.cuh", "instantiate.cuh"):
        with open(os.path.join(ROOT, "hemelb_b200", "csrc", f), "r") as fh:
            text = fh.read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        text = re.sub(r"//[^\n]*", "", text)
        h.update("".join(text.split()).encode())
    return h.hexdigest()[:16]


def measured_traffic(sites_per_launch):
    """DRAM bytes per launch of the site kernel from a committed ncu --set full capture -- only when
    that capture was taken on this very kernel source and launch shape; else None."""
    p = os.path.join(ROOT, "profiles", "r02_site_kernel_traffic.json")
    try:
        with open(p) as fh:
            t = json.load(fh)
        if t.get("kernel_source_hash") == kernel_source_hash() and int(t["sites_per_launch"]) == int(sites_per_launch):
            return float(t["dram_bytes_read"]) + float(t["dram_bytes_write"])
    except Exception:
        pass
    return None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        self.index = index

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                self.samples.append(float(parts[0]))
                self.max_mhz = float(parts[1])
                for n, v in zip(names, parts[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=5)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------------
def tree_for_sites(total_sites: float, generations: int = GENERATIONS):
    """Root radius / length so that the tree holds about ``total_sites`` fluid sites: every
    generation carries ~pi r^2 L (Murray's law with L shrinking by 0.8), L = 4.17 r."""
    per_gen = total_sites / generations
    r = (per_gen / (np.pi * 4.17)) ** (1.0 / 3.0)
    return float(r), float(4.17 * r)


def pressure_iolets(meta, radius_of=lambda p: p.radius):
    from hemelb_b200.capi import iolet_record
    from hemelb_b200.lbm import prepare_boundary_objects
    ins = [iolet_record(0, tuple(p.normal), tuple(p.position), radius=radius_of(p), density_mean=1.0005,
                        density_amp=0.0, period=1000.0) for p in meta["inlets"]]
    outs = [iolet_record(0, tuple(p.normal), tuple(p.position), radius=radius_of(p), density_mean=0.9995,
                         density_amp=0.0, period=1000.0) for p in meta["outlets"]]
    prepare_boundary_objects(ins, outs)
    return ins, outs


def tree_block_partition(dom, nparts, decomposition, partition_start, gather=None, rank=0, world=1):
    """rank of every lattice block (.gmy block order) for ``nparts`` ranks.  The per-block fluid-site
    counts come from the device (each of ``world`` processes counts an x-slab of blocks)."""
    from hemelb_b200.devdomain import basic_decomposition_of_counts, weighted_decomposition_of_counts
    bd = dom.block_dims
    xs = [int(bd[0]) * r // world for r in range(world + 1)]
    weighted = decomposition == "weighted"
    lo, hi = [xs[rank], 0, 0], [xs[rank + 1], int(bd[1]), int(bd[2])]
    mine = np.stack(dom.count_block_sites_typed(lo, hi)) if weighted else dom.count_block_sites(lo, hi)
    parts = gather(mine) if gather is not None else [mine]
    counts = np.concatenate(parts, 1 if weighted else 0)
    if weighted:
        return weighted_decomposition_of_counts(counts[0], counts[1], nparts, "BFL", initial=partition_start), int(counts[0].sum())
    return basic_decomposition_of_counts(counts, nparts), int(counts.sum())


def build_tree(total_sites, rank, world, device, decomposition="basic", partition_start="inertial", gather=None):
    """This rank's Domain tables of the tree, voxelised and built on its GPU (hlb_dom_*)."""
    from hemelb_b200.devdomain import DeviceDomain, tree_shape
    r0, l0 = tree_for_sites(total_sites)
    caps, iolets, shape = tree_shape(GENERATIONS, r0, l0)
    dom = DeviceDomain.from_shape(caps, iolets, shape, Q, 8, None, rank, world, device, build=False)
    if world > 1:
        rob, _ = tree_block_partition(dom, world, decomposition, partition_start, gather, rank, world)
        dom.set_partition(("blocks", rob))
    dom.build()
    ins, outs = pressure_iolets(dom.meta)
    return dom, ins, outs, (r0, l0)


def build_cylinder(radius, length, rank, nranks, device=0, block_size=8):
    """configs[1]: this rank's z-slab of a cylinder nranks times longer, built on the GPU."""
    from hemelb_b200.devdomain import DeviceDomain, cylinder_shape
    total_len = length * nranks
    caps, iolets, shape = cylinder_shape(radius, total_len)
    partition = None
    if nranks > 1:
        margin, per, big = 2, total_len // nranks, 2 ** 60
        partition = ("slabs", 2, [-big] + [margin + r * per for r in range(1, nranks)] + [big])
    dom = DeviceDomain.from_shape(caps, iolets, shape, Q, block_size, partition, rank, nranks, device)
    ins, outs = pressure_iolets(dom.meta, lambda p: radius)
    return dom, ins, outs


def workload_text(sites_per_gpu):
    return ("configs[2]: synthetic bifurcating vascular tree, %d generations (Murray's law, half-angle 35 deg, seed "
            "20261017), ~%.3g fluid sites per GPU, D3Q19 LBGK + BFL walls + Nash pressure iolets (1 inlet, 32 "
            "outlets), tau=%g" % (GENERATIONS, sites_per_gpu, TAU))


# ------------------------------------------------------------------------------------------------
# the reference's own code on the host cores
# ------------------------------------------------------------------------------------------------
def usable_cores():
    """Host threads this process may really use: the affinity mask, capped by a cgroup CPU quota when there is one
    (the emulated ranks of the reference arm poll for their neighbours as MPI ranks do: more threads than cores
    would cost the reference its rate)."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        pass
    for path in ("/sys/fs/cgroup/cpu.max", ):
        try:
            quota, period = open(path).read().split()[:2]
            if quota != "max":
                n = min(n, max(1, int(float(quota) / float(period))))
        except (OSError, ValueError):
            pass
    try:
        q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
        per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
        if q > 0 and per > 0:
            n = min(n, max(1, q // per))
    except (OSError, ValueError):
        pass
    return max(1, n)


def reference_cross_check(timeout=300):
    """Is the timed arm a fair stand-in for the reference?  scripts/reference_arm_ab.py, in a child process (a crash
    there cannot cost this line): the arm's rank driver against the reference's whole lb::LBM over its own net::Net
    and StepManager (oracle/_ref/libhemelb_reflbm_sse3.so) on the same 1e6-site cylinder, decomposition and cores,
    two interleaved repetitions each.  MLUPS lists, or why there are none."""
    import oracle as O
    if O.ref_lbm_lib(True) is None:
        return {"unavailable": "oracle/_ref/libhemelb_reflbm_sse3.so was not built / shipped"}
    try:
        run = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "reference_arm_ab.py"), "40", "200", "--json",
                              "--sse3-only", "--reps", "2"], capture_output=True, text=True, timeout=timeout, cwd=ROOT)
        if run.returncode != 0:
            return {"unavailable": "child ended with %d: %s" % (run.returncode, run.stderr.strip()[-200:])}
        return json.loads(run.stdout.strip().splitlines()[-1])
    except Exception as e:  # noqa: BLE001 -- whatever went wrong, the bench line itself stands
        return {"unavailable": repr(e)[:200]}


def cpu_reference_run(steps, warmup, target_seconds=15.0, sample_sites=REFERENCE_SAMPLE_SITES):
    """The reference's own streamers / kernels (oracle/_ref, SSE3 build = the x86-64 default) on all
    host cores: one emulated rank (thread) per core, BasicDecomposition over Morton blocks, in-memory
    halo copies, on a bounded sample of the bench workload: the same tree generator at
    ``sample_sites`` sites (its tables are built on the GPU when there is one -- untimed set-up, and
    the only way to get a tree too large for the host caches in seconds; else a small numpy-built tree)."""
    import oracle as O
    from hemelb_b200 import capi
    cores = usable_cores()
    sse3 = O.ref_lib(True) is not None
    have_ref = O.ref_lib(sse3) is not None
    R = cores if have_ref else 1
    import ctypes as C
    ndev = C.c_int(0)
    on_gpu = capi.lib().hlb_gpu_device_count(C.byref(ndev)) == 0 and ndev.value > 0
    if on_gpu:
        from hemelb_b200.devdomain import DeviceDomain, tree_shape
        r0, l0 = tree_for_sites(sample_sites)
        caps, iolets, shape = tree_shape(GENERATIONS, r0, l0)
        probe = DeviceDomain.from_shape(caps, iolets, shape, Q, 8, None, 0, R, 0, build=False)
        rob = None
        if R > 1:
            rob, _ = tree_block_partition(probe, R, "basic", "morton")
        meta = probe.meta
        probe.close()
        tables = []
        for r in range(R):
            d = DeviceDomain.from_shape(caps, iolets, shape, Q, 8, ("blocks", rob) if rob is not None else None, r, R, 0)
            tables.append(d.tables())
            d.close()
        built = "tables built on the GPU by hlb_dom_* (untimed set-up)"
        geom_text = "tree, %d generations, root r=%.1f l=%.1f" % (GENERATIONS, r0, l0)
        geom = None
    else:
        from hemelb_b200 import geometry as G
        from hemelb_b200.domain import build_domains
        geom = G.capsule_tree(3, 9.0, 30.0)
        rank_of = G.basic_decomposition(geom, R) if R > 1 else None
        doms = build_domains(geom, Q, rank_of, R)
        tables = [d.tables() for d in doms]
        meta = geom.meta
        built = "tables built on the host (numpy); no GPU here, so a small tree"
        geom_text = "tree, 3 generations, root r=9 l=30"
    n_sites = int(sum(t["N"] for t in tables))
    inlets, outlets = pressure_iolets(meta)
    kind = "reference"
    if have_ref:
        # tau = 0.8 through LbmParameters(dt, dx, rho, eta): dt = (tau - 0.5) * Cs2 * dx^2 * rho / eta
        dt = (TAU - 0.5) / 3.0 * 1000.0 / 0.004
        sim = O.RefSim(tables, Q, "LBGK", "BFL", "NASH", "NASH", dt=dt, dx=1.0, rho=1000.0, eta=0.004, inlets=inlets,
                       outlets=outlets, sse3=sse3)
        stepper = sim.step_mt
    else:
        kind = "port"
        sim = O.OracleSim(O.OracleDomains(geom, Q, None, 1), "LBGK", "BFL", tau=TAU, inlets=inlets, outlets=outlets)
        stepper = sim.step
        cores = 1
    _, w, _ = O.lattice(Q)
    for r, t in enumerate(tables):
        f = np.zeros(t["N"] * Q + 1 + t["totalSharedFs"])
        f[:t["N"] * Q] = np.tile(w, t["N"])
        sim.set_f(f, r)
        sim.set_f(f, r, 1)
    stepper(max(1, warmup))
    t0 = time.perf_counter()
    stepper(1)
    per = time.perf_counter() - t0
    if steps is None:
        steps = int(max(3, min(200, target_seconds / max(per, 1e-6))))
    t0 = time.perf_counter()
    stepper(steps)
    dtm = time.perf_counter() - t0
    mlups = n_sites * steps / dtm / 1e6
    sample = "%s (%d sites), D3Q19 LBGK+BFL+Nash, %d steps, %d threads (one emulated rank each for the whole run, BasicDecomposition, neighbour waits + one barrier per step), %s build; %s" % (
        geom_text, n_sites, steps, cores, "SSE3" if sse3 else "scalar", built)
    return dict(value=mlups, unit="MLUPS", cores=cores, kind=kind, sample=sample, sites=n_sites), dtm / steps * 1e3, steps


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)   # time >= 200 steps
    ap.add_argument("--warmup", type=int, default=20)  # ... after >= 20 warm-up steps
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sites-per-gpu", type=float, default=1.18e8)  # (the generator lands ~6 % under: 1.11e8 per GPU, 8.9e8 on eight)
    ap.add_argument("--decomposition", default="basic", choices=["basic", "weighted"])
    ap.add_argument("--partition-start", default="inertial", choices=["morton", "rcb", "inertial", "best"])
    ap.add_argument("--radius", type=float, default=146.0, help="secondary record: cylinder radius")
    ap.add_argument("--length", type=int, default=1500, help="secondary record: cylinder length per GPU")
    ap.add_argument("--no-secondary", action="store_true", help="skip the configs[1] cylinder record")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cross-check", action="store_true",
                    help="--impl reference: skip the comparison of the timed arm with the reference's whole lb::LBM")
    ap.add_argument("--no-reorder", action="store_true")
    args = ap.parse_args()
    # stdout carries the JSON line and nothing else: whatever a library prints to fd 1 (NCCL's version
    # banner, for one) goes to stderr
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    warmup = max(args.warmup, 3)
    decomposition = ("single rank" if world == 1 else
                     "BasicDecomposition over Morton-ordered 8^3 blocks (Code/geometry/decomposition/BasicDecomposition.cc), rank r -> GPU r"
                     if args.decomposition == "basic" else
                     "weighted k-way over 8^3 blocks, %s start (hemelb_b200/partition.py), rank r -> GPU r" % args.partition_start)
    config = {"workload": workload_text(args.sites_per_gpu), "decomposition": decomposition,
              "l2": "inputs (>= 40 GB per GPU at the default size) exceed the 126 MB L2; no flush needed",
              "reference_arm": "--impl reference times the reference's own streamers on the host cores on a bounded "
                               "sample: the same tree generator at ~%.3g sites, one emulated rank per core "
                               "(cpu_baseline.sample has the exact figures)" % REFERENCE_SAMPLE_SITES}

    if args.impl == "reference":
        if rank != 0:
            return 0
        base, ms, steps = cpu_reference_run(args.steps if args.steps else None, warmup)
        line = {"metric": METRIC, "value": base["value"], "unit": "MLUPS", "n_gpus": args.gpus, "steps": steps,
                "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference", "config": config,
                "cpu_baseline": base,
                "e2e": {"value": base["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        if world == 1 and not args.no_cross_check:
            line["cross_check"] = reference_cross_check()
        print(json.dumps(line), file=json_out, flush=True)
        return 0

    from hemelb_b200.lbm import GpuLBM
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def gather(x):
        parts = [None] * world
        dist.all_gather_object(parts, x)
        return parts

    def reduce_max(vals):
        if dist is None:
            return [float(v) for v in vals]
        import torch
        t = torch.tensor(list(vals), dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    def measure(dom, inlets, outlets, steps, e2e):
        """value / roofline / e2e figures of one engine over one Domain."""
        gpu = GpuLBM.from_device_domain(dom, "LBGK", "BFL", "NASH", "NASH", tau=TAU, inlets=inlets, outlets=outlets,
                                        reorder=not args.no_reorder)
        if world > 1:
            uid = [GpuLBM.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            gpu.comm_init(uid[0])
        gpu.set_equilibrium(1.0, (0.0, 0.0, 0.0))

        def barrier():
            gpu.sync()
            if dist is not None:
                import torch
                dist.barrier()
                torch.cuda.synchronize()

        out = {"target_runs": gpu.target_runs()}
        gpu.step(warmup)
        barrier()
        sampler = ClockSampler(local_rank)
        sampler.start()
        l0 = gpu.launch_count()
        ms, part_ms, part_sites = gpu.time_steps_detail(steps)  # CUDA events around the mid-domain site kernel
        out["launches"] = gpu.launch_count() - l0
        barrier()
        out["clocks"] = sampler.stop()
        out["ms"], = reduce_max([ms])
        out["part_ms"], out["part_sites"] = part_ms, part_sites
        if e2e:
            # end to end through the phase API with host scalars every step
            gpu.set_cache_mask(256)  # HLB_CACHE_MONITOR: stability / incompressibility monitors gathered in-kernel
            gpu.do_time_step()
            gpu.monitor()
            barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                gpu.do_time_step()   # set_step_scalars (H2D of iolet densities) + the LBM phase calls
                mon = gpu.monitor()  # D2H: {min f, min/max density, max speed}
            gpu.sync()
            out["e2e_s"], = reduce_max([time.perf_counter() - t0])
            out["monitor"] = mon
        else:
            out["monitor"] = gpu.monitor()
        gpu.close()
        return out

    # ---- primary: the tree
    t_setup = time.time()
    total = args.sites_per_gpu * world
    dom, inlets, outlets, (r0, l0) = build_tree(total, rank, world, local_rank, args.decomposition, args.partition_start,
                                                gather if world > 1 else None)
    t_setup = time.time() - t_setup
    per_rank = [(dom.N, int(dom.totalSharedFs), int(dom.procs.shape[0]), int(dom.N - dom.mid[0] - dom.edge[0]),
                 int(dom.N - dom.mid.sum()))]
    if dist is not None:
        per_rank = gather(per_rank[0])
    n_sites_global = int(sum(p[0] for p in per_rank))
    mid_sites0, by_mid, by_edge, S0, build_s = int(dom.mid.sum()), [int(x) for x in dom.mid], [int(x) for x in dom.edge], \
        int(dom.totalSharedFs), dom.build_seconds
    res = measure(dom, inlets, outlets, args.steps, True)
    dom.close()
    mlups = n_sites_global * args.steps / (res["ms"] * 1e-3) / 1e6
    e2e_mlups = n_sites_global * args.steps / res["e2e_s"] / 1e6

    # ---- secondary: the cylinder
    secondary = None
    if not args.no_secondary:
        cdom, cin, cout = build_cylinder(args.radius, args.length, rank, world, local_rank)
        cn = [cdom.N]
        if dist is not None:
            cn = gather(cdom.N)
        cres = measure(cdom, cin, cout, args.steps, False)
        cdom.close()
        cyl_mlups = int(sum(cn)) * args.steps / (cres["ms"] * 1e-3) / 1e6
        secondary = {"workload": "configs[1]: straight cylinder r=%g l=%d per GPU (z-slabs, one per GPU), same policies"
                                 % (args.radius, args.length),
                     "sites": int(sum(cn)), "value": cyl_mlups, "unit": "MLUPS", "ms_per_step": cres["ms"] / args.steps}

    if rank != 0:
        return 0
    peak, peak_kind = measured_peak()
    part_gbs = (res["part_sites"] * BYTES_PER_SITE / 1e9) / (res["part_ms"] * 1e-3) if res["part_ms"] > 0 else 0.0
    if secondary is not None:
        cg = (cres["part_sites"] * BYTES_PER_SITE / 1e9) / (cres["part_ms"] * 1e-3) if cres["part_ms"] > 0 else 0.0
        secondary["site_kernel_frac"] = cg / peak
        secondary["whole_step_frac"] = (cyl_mlups * 1e6 * BYTES_PER_SITE / 1e9 / world) / peak
    line = {
        "metric": METRIC, "value": mlups, "unit": "MLUPS", "n_gpus": world, "steps": args.steps, "warmup": warmup,
        "ms_per_step": res["ms"] / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": config,
        "sites": {"global": n_sites_global, "per_rank": [p[0] for p in per_rank],
                  "halo_doubles_per_rank": [p[1] for p in per_rank], "neighbours_per_rank": [p[2] for p in per_rank],
                  "boundary_typed_per_rank": [p[3] for p in per_rank], "domain_edge_per_rank": [p[4] for p in per_rank],
                  "rank0_by_type_mid": by_mid, "rank0_by_type_edge": by_edge, "tree_root_radius": r0, "tree_root_length": l0},
        "roofline": {"bound": "hbm", "achieved": part_gbs, "peak": peak, "unit": "GB/s",
                     "frac": part_gbs / peak if peak else None,
                     "traffic": measured_traffic(mid_sites0),
                     "traffic_unit": "bytes per launch (ncu dram read+write); null unless a capture of this very kernel "
                                     "source and launch shape is committed (profiles/r02_site_kernel_traffic.json)",
                     "algorithmic_bytes_per_launch": mid_sites0 * BYTES_PER_SITE,
                     "kernel": SITE_KERNEL + " over rank 0's mid-domain part (all six collision types in one launch, "
                               "sites in lattice order)",
                     "kernel_source_hash": kernel_source_hash(),
                     "bytes_per_site": BYTES_PER_SITE,
                     "streaming_targets": "read as <= 2 runs per direction (8 B) for %d of rank 0's %d groups of 32 sites, "
                                          "from the index planes (128 B) for the rest: the kernel moves fewer bytes than the "
                                          "algorithmic 20 Q per site, so frac can pass 1" % tuple(res["target_runs"]),
                     "peak_kind": peak_kind + " HBM copy (burst)",
                     "kernel_share_of_step": res["part_ms"] / res["ms"] if res["ms"] else None,
                     "timed_in": "the timed region of `value` itself, CUDA events on the engine's stream",
                     "whole_step_frac": (mlups * 1e6 * BYTES_PER_SITE / 1e9 / world) / peak},
        "e2e": {"value": e2e_mlups, "unit": "MLUPS", "h2d_bytes_per_step": 8 * (len(inlets) + len(outlets)),
                "d2h_bytes_per_step": 32,
                "path": "hlb_gpu_set_step_scalars + request_comms/stream_and_collide x12/edge_done/copy_received/"
                        "post_step x12/swap + hlb_gpu_monitor per step, from Python over ctypes"},
        "secondary": secondary,
        "gpu_launches": int(res["launches"]), "clocks": res["clocks"], "setup_seconds": t_setup,
        "tables": "device (hlb_dom_*), %.3f s of kernels on rank 0" % build_s,
        "monitor": res["monitor"],
    }
    if not args.no_cpu_baseline and world == 1:
        try:
            base, _, _ = cpu_reference_run(None, 1)
            line["cpu_baseline"] = base
        except Exception as e:  # the baseline must not take the GPU line down
            line["cpu_baseline"] = {"value": None, "unit": "MLUPS", "cores": 0, "kind": "unavailable", "sample": repr(e)}
    print(json.dumps(line), file=json_out, flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
