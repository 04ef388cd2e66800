#!/usr/bin/env python
"""MLUPS benchmark of the collide-and-stream hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N = 1: configs[1] -- synthetic straight cylinder (radius 146, length 1500 => ~1.0e8 fluid sites),
D3Q19 LBGK + BouzidiFirdaousLallemand walls, Nash pressure iolets, on one B200.  N > 1 (torchrun,
one rank per GPU): the same cylinder made N times longer and cut into N z-slabs (weak scaling,
~1e8 sites per GPU), halo over NCCL send/recv.

One JSON line on rank 0: value = whole-job MLUPS with everything resident in HBM (CUDA events, max
over ranks); e2e = the same steps driven phase by phase through the C ABI with the per-step host
scalars copied H2D and a monitor read back D2H every step; roofline = the mid-fluid (bulk) kernel
timed live with CUDA events against the measured HBM copy bandwidth; cpu_baseline = the reference's
own streamers/kernels (oracle/_ref) on the host cores, a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
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


def measured_traffic(bulk_sites_per_launch):
    """DRAM bytes per launch of the bulk kernel from the committed ncu --set full capture, if that
    capture was taken on this very launch shape; else None."""
    p = os.path.join(ROOT, "profiles", "r01_bulk_fullsize_traffic.json")
    try:
        with open(p) as fh:
            t = json.load(fh)
        if int(t["sites_per_launch"]) == int(bulk_sites_per_launch):
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


def cylinder_iolets(geom_meta, radius):
    from hemelb_b200.capi import iolet_record
    from hemelb_b200.lbm import prepare_boundary_objects
    inl, outl = geom_meta["inlets"][0], geom_meta["outlets"][0]
    inlets = [iolet_record(0, tuple(inl.normal), tuple(inl.position), radius=radius, density_mean=1.0005,
                           density_amp=0.0, period=1000.0)]
    outlets = [iolet_record(0, tuple(outl.normal), tuple(outl.position), radius=radius, density_mean=0.9995,
                            density_amp=0.0, period=1000.0)]
    prepare_boundary_objects(inlets, outlets)
    return inlets, outlets


def build_workload(radius, length, rank, nranks, block_size=8, device=0):
    """This rank's Domain tables, built on the GPU straight from the cylinder's analytic shape
    (hlb_dom_*: voxelisation, site order, neighbourIndices, halo tables).  N > 1: z-slabs of a
    cylinder nranks times longer; each rank voxelises only its own slab plus one voxel of rim."""
    from hemelb_b200.devdomain import DeviceDomain, cylinder_shape
    total_len = length * nranks
    caps, iolets, shape = cylinder_shape(radius, total_len)
    partition = None
    if nranks > 1:
        margin, per, big = 2, total_len // nranks, 2 ** 60
        partition = ("slabs", 2, [-big] + [margin + r * per for r in range(1, nranks)] + [big])
    dom = DeviceDomain.from_shape(caps, iolets, shape, Q, block_size, partition, rank, nranks, device)
    inlets, outlets = cylinder_iolets(dom.meta, radius)
    return dom, inlets, outlets


def build_workload_host(radius, length, rank, nranks, block_size=8):
    """The same workload through the host (numpy) voxeliser and Domain builder (--host-tables)."""
    from hemelb_b200 import geometry as G
    from hemelb_b200.domain import DomainBuilder
    total_len = length * nranks
    if nranks == 1:
        geom = G.cylinder_extruded(radius, total_len, block_size)
        rank_of_site = None
    else:
        geom, rank_of_site = G.cylinder_slab(radius, total_len, nranks, rank, block_size)
    dom = DomainBuilder(geom, Q, rank_of_site, nranks).domains[rank]
    inlets, outlets = cylinder_iolets(geom.meta, radius)
    return dom, inlets, outlets


def cpu_reference_run(steps, warmup, target_seconds=12.0, radius=40.0, length=320):
    """The reference's own streamers / kernels (oracle/_ref, SSE3 build = the x86-64 default) on all
    host cores: one emulated rank (thread) per core, in-memory halo copies."""
    import oracle as O
    from hemelb_b200 import geometry as G
    from hemelb_b200.domain import build_domains
    from hemelb_b200.capi import iolet_record
    from hemelb_b200.lbm import prepare_boundary_objects
    cores = os.cpu_count() or 1
    sse3 = O.ref_lib(True) is not None
    kind = "reference"
    geom = G.cylinder_extruded(radius, length)
    R = cores
    rank = G.slab_decomposition(geom, R) if R > 1 else None
    doms = build_domains(geom, Q, rank, R)
    inl, outl = geom.meta["inlets"][0], geom.meta["outlets"][0]
    inlets = [iolet_record(0, tuple(inl.normal), tuple(inl.position), radius=radius, density_mean=1.0005)]
    outlets = [iolet_record(0, tuple(outl.normal), tuple(outl.position), radius=radius, density_mean=0.9995)]
    prepare_boundary_objects(inlets, outlets)
    if O.ref_lib(sse3) is not None:
        # tau = 0.8 through LbmParameters(dt, dx, rho, eta): dt = (tau - 0.5) * Cs2 * dx^2 * rho / eta
        dt = (TAU - 0.5) / 3.0 * 1000.0 / 0.004
        sim = O.RefSim([d.tables() for d in doms], Q, "LBGK", "BFL", "NASH", "NASH", dt=dt, dx=1.0, rho=1000.0,
                       eta=0.004, inlets=inlets, outlets=outlets, sse3=sse3)
        stepper = sim.step_mt
    else:
        kind = "port"
        sim = O.OracleSim(O.OracleDomains(geom, Q, rank, R), "LBGK", "BFL", tau=TAU, inlets=inlets, outlets=outlets)
        stepper = sim.step
        cores = 1
    _, w, _ = O.lattice(Q)
    for r, d in enumerate(doms):
        f = np.zeros(d.N * Q + 1 + d.totalSharedFs)
        f[:d.N * Q] = np.tile(w, d.N)
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
    mlups = geom.n_sites * steps / dtm / 1e6
    sample = "cylinder r=%g l=%d (%d sites), %d steps, %d threads (one emulated rank each), %s build" % (
        radius, length, geom.n_sites, steps, cores, "SSE3" if sse3 else "scalar")
    return dict(value=mlups, unit="MLUPS", cores=cores, kind=kind, sample=sample), dtm / steps * 1e3, steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)   # time >= 200 steps
    ap.add_argument("--warmup", type=int, default=20)  # ... after >= 20 warm-up steps
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--radius", type=float, default=146.0)
    ap.add_argument("--length", type=int, default=1500)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reorder", action="store_true")
    ap.add_argument("--host-tables", action="store_true", help="voxelise and build the Domain tables on the host (numpy)")
    ap.add_argument("--block-size", type=int, default=8, help="sites per block side of the synthetic .gmy (HemeLB default 8)")
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
    config = {"workload": "configs[1]: straight cylinder r=%g l=%d per GPU, D3Q19 LBGK + BFL walls + Nash pressure "
                          "iolets, tau=%g" % (args.radius, args.length, TAU),
              "decomposition": "z-slabs, one per GPU" if world > 1 else "single rank",
              "l2": "inputs (>= 30 GB per GPU at the default size) exceed the 126 MB L2; no flush needed"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        base, ms, steps = cpu_reference_run(args.steps if args.steps else None, warmup)
        line = {"metric": METRIC, "value": base["value"], "unit": "MLUPS", "n_gpus": args.gpus, "steps": steps,
                "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference", "config": config,
                "cpu_baseline": base,
                "e2e": {"value": base["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), file=json_out, flush=True)
        return 0

    from hemelb_b200.lbm import GpuLBM
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    t_setup = time.time()
    if args.host_tables:
        dom, inlets, outlets = build_workload_host(args.radius, args.length, rank, world, args.block_size)
        gpu = GpuLBM(dom, "LBGK", "BFL", "NASH", "NASH", tau=TAU, inlets=inlets, outlets=outlets, device=local_rank,
                     reorder=not args.no_reorder)
    else:
        dom, inlets, outlets = build_workload(args.radius, args.length, rank, world, args.block_size, local_rank)
        gpu = GpuLBM.from_device_domain(dom, "LBGK", "BFL", "NASH", "NASH", tau=TAU, inlets=inlets, outlets=outlets,
                                        reorder=not args.no_reorder)
    if world > 1:
        import torch
        uid = [GpuLBM.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        gpu.comm_init(uid[0])
    gpu.set_equilibrium(1.0, (0.0, 0.0, 0.0))
    t_setup = time.time() - t_setup

    def barrier():
        gpu.sync()
        if dist is not None:
            dist.barrier()
            import torch
            torch.cuda.synchronize()

    n_sites_local = dom.N
    n_sites_global = n_sites_local
    if dist is not None:
        import torch
        t = torch.tensor([n_sites_local], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        n_sites_global = int(t.item())

    # ---- device-resident throughput
    gpu.step(warmup)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = gpu.launch_count()
    # product schedule: the mid-fluid kernel pre-writes the slots the boundary ranges fill after it,
    # every range its own kernel on one stream; CUDA events around each mid-fluid launch
    ms, bulk_ms, bulk_sites = gpu.time_steps_detail(args.steps)
    launches = gpu.launch_count() - l0
    barrier()
    # A/B region, same K steps in the reference's plain write order (hlb_gpu_set_overlap(0))
    gpu.set_overlap(False)
    gpu.step(2)
    barrier()
    serial_ms, plain_bulk_ms, plain_bulk_sites = gpu.time_steps_detail(args.steps)
    gpu.set_overlap(True)
    barrier()
    clocks = sampler.stop()
    if dist is not None:
        import torch
        t = torch.tensor([ms, serial_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, serial_ms = float(t[0].item()), float(t[1].item())
    mlups = n_sites_global * args.steps / (ms * 1e-3) / 1e6

    # ---- end to end through the phase API with host scalars every step
    gpu.set_cache_mask(256)  # HLB_CACHE_MONITOR: stability / incompressibility monitors gathered in-kernel
    gpu.do_time_step()
    gpu.monitor()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        gpu.do_time_step()   # set_step_scalars (H2D of iolet densities) + the LBM phase calls
        mon = gpu.monitor()  # D2H: {min f, min/max density, max speed}
    gpu.sync()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        import torch
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_mlups = n_sites_global * args.steps / e2e_s / 1e6

    if rank != 0:
        return 0
    peak, peak_kind = measured_peak()
    bulk_gbs = (bulk_sites * BYTES_PER_SITE / 1e9) / (bulk_ms * 1e-3) if bulk_ms > 0 else 0.0
    line = {
        "metric": METRIC, "value": mlups, "unit": "MLUPS", "n_gpus": world, "steps": args.steps, "warmup": warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": config,
        "sites": {"global": n_sites_global, "rank0": n_sites_local, "rank0_by_type_mid": [int(x) for x in dom.mid],
                  "rank0_by_type_edge": [int(x) for x in dom.edge], "halo_doubles_rank0": int(dom.totalSharedFs)},
        "roofline": {"bound": "hbm", "achieved": bulk_gbs, "peak": peak, "unit": "GB/s",
                     "frac": bulk_gbs / peak if peak else None,
                     "traffic": measured_traffic(int(dom.mid[0])), "traffic_unit": "bytes per launch (ncu dram read+write)",
                     "algorithmic_bytes_per_launch": int(dom.mid[0]) * BYTES_PER_SITE,
                     "kernel": "collide_stream_kernel<19,LBGK,none,none> (mid-fluid range)",
                     "bytes_per_site": BYTES_PER_SITE, "peak_kind": peak_kind + " HBM copy (burst)",
                     "kernel_share_of_step": bulk_ms / ms if ms else None,
                     "timed_in": "the timed region of `value` itself (product schedule: the mid-fluid kernel runs "
                                 "alone on the engine's stream, the boundary ranges follow it)",
                     "plain_order": {"what": "same K steps without the hole pre-write (hlb_gpu_set_overlap(0))",
                                     "ms_per_step": serial_ms / args.steps,
                                     "bulk_kernel_frac": (plain_bulk_sites * BYTES_PER_SITE / 1e9) /
                                                         (plain_bulk_ms * 1e-3) / peak if plain_bulk_ms else None},
                     "whole_step_frac": (mlups * 1e6 * BYTES_PER_SITE / 1e9 / world) / peak},
        "e2e": {"value": e2e_mlups, "unit": "MLUPS", "h2d_bytes_per_step": 8 * (len(inlets) + len(outlets)),
                "d2h_bytes_per_step": 32,
                "path": "hlb_gpu_set_step_scalars + request_comms/stream_and_collide x12/edge_done/copy_received/"
                        "post_step x12/swap + hlb_gpu_monitor per step, from Python over ctypes"},
        "gpu_launches": int(launches), "clocks": clocks, "setup_seconds": t_setup,
        "tables": "host (numpy)" if args.host_tables else "device (hlb_dom_*), %.3f s of kernels" % dom.build_seconds,
        "monitor": mon,
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
