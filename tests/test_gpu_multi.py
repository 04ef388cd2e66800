"""Multi-GPU parity: one process per GPU, halo over NCCL send/recv (hlb_gpu_comm_init), against
the oracle's emulated-rank run.  Needs >= 2 GPUs (gpurun --gpus 2); skipped otherwise."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

WORKER = r'''
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, %(root)r)
import oracle as O
from hemelb_b200 import geometry as G
from hemelb_b200.domain import DomainBuilder
from hemelb_b200.lbm import GpuLBM
from tests.cases import anisotropic_f, iolets_for

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
Q, L, radius, steps = 19, 64, 6.3, 12
for (kernel, wall, inlet, outlet) in (("LBGK", "BFL", "NASH", "NASH"), ("MRT", "SBB", "LADD", "NASH"),
                                     ("LBGK", "GZS", "LADD", "NASH")):
    if os.environ.get("HLB_CASE") == "tree_sites":
        # site-granular partition of the tree (inertial start + site stage): shared blocks, several neighbours
        from hemelb_b200 import partition as P
        from tests.cases import geometry
        from tests.test_partition import collision_types
        full = geometry("tree")
        full_rank, _ = P.partition_sites(full, collision_types(full, Q), Q, nranks=world, initial="inertial", native=True)
        dom = DomainBuilder(full, Q, full_rank, world).domains[rank]
    else:
        sub, sub_rank = G.cylinder_slab(radius, L, world, rank)
        full = G.cylinder_extruded(radius, L)
        full_rank = np.minimum((full.coords[:, 2].astype(np.int64) - 2) // (L // world), world - 1).astype(np.int32)
        if wall == "GZS":
            # the phase-0 site halo (NeighbouringDataManager.cc:101-142) names sites of other ranks:
            # tables from the whole geometry
            dom = DomainBuilder(full, Q, full_rank, world).domains[rank]
        else:
            dom = DomainBuilder(sub, Q, sub_rank, world).domains[rank]
    inlets, outlets = iolets_for(full, inlet, outlet)
    gpu = GpuLBM(dom, kernel, wall, inlet, outlet, tau=0.8, inlets=inlets, outlets=outlets, device=rank)
    if wall == "GZS":
        needs = [None] * world
        dist.all_gather_object(needs, int(gpu.gzs_need.shape[0]))
        assert sum(needs) > 0, "no GZS link extrapolates across a rank boundary: the case tests nothing"
    uid = [GpuLBM.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    gpu.comm_init(uid[0])
    gpu.set_f(anisotropic_f(dom.N, Q, dom.totalSharedFs, site_offset=7 * rank))
    # half the steps inside the library, half driven phase by phase from the host
    gpu.step(steps // 2)
    for _ in range(steps - steps // 2):
        gpu.do_time_step()
    mine = gpu.get_f()[:dom.N * Q]
    odom = O.OracleDomains(full, Q, full_rank, world)
    ref = O.OracleSim(odom, kernel, wall, inlet, outlet, tau=0.8, inlets=inlets, outlets=outlets)
    for r in range(world):
        t = odom.tables(r)
        ref.set_f(anisotropic_f(t["N"], Q, t["totalSharedFs"], site_offset=7 * r), r)
    ref.step(steps)
    want = ref.get_f(rank)[:dom.N * Q]
    err = float(np.abs(mine - want).max())
    assert err <= 1e-13, (kernel, wall, err)
    assert np.array_equal(mine, want), (kernel, wall, "not bit-identical", err)
    # the monitors over all ranks (ncclAllReduce) against the per-rank ones combined on the host
    if os.environ.get("HLB_CHECK_MONITOR") != "1":
        gpu.close()
        dist.barrier()
        continue
    local = gpu.monitor()
    every = [None] * world
    dist.all_gather_object(every, local)
    glob = gpu.monitor_global()
    assert glob["min_f"] == min(m["min_f"] for m in every)
    assert glob["min_density"] == min(m["min_density"] for m in every)
    assert glob["max_density"] == max(m["max_density"] for m in every)
    assert glob["max_speed"] == max(m["max_speed"] for m in every)
    gpu.close()
    dist.barrier()
print("rank", rank, "ok")
'''


def _gpu_count():
    from hemelb_b200 import capi
    import ctypes as C
    n = C.c_int(0)
    if capi.lib().hlb_gpu_device_count(C.byref(n)) != 0:
        return 0
    return n.value


def run_workers(tmp_path, world, case, check_monitor=False):
    if _gpu_count() < world:
        pytest.skip("needs %d GPUs" % world)
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   HLB_CASE=case, HLB_CHECK_MONITOR="1" if check_monitor else "0")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=600)
        except subprocess.TimeoutExpired:
            p.kill()
            out, _ = p.communicate()
        outs.append(out)
    for r, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, "rank %d failed:\n%s" % (r, out[-3000:])


@pytest.mark.parametrize("world", [2, 4, 8])
def test_nccl_halo_matches_oracle(tmp_path, world):
    run_workers(tmp_path, world, "cylinder_slabs")
