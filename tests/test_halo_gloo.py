"""N > 1 host logic on CPU: two real processes (torch.distributed, gloo) each own one rank's Domain
tables, run that rank's ranges with the oracle's per-range API, and exchange the halo slices named
by ``neighbouringProcs`` with send/recv -- exactly the offsets and pairing the NCCL path uses
(FieldData.cc:27-48).  Result must equal the single-process multi-rank oracle run."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, %(root)r)
import oracle as O
from hemelb_b200 import geometry as G
from hemelb_b200.domain import DomainBuilder
from tests.cases import anisotropic_f, iolets_for

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
Q, L, radius, steps = 19, 48, 5.3, 6
if os.environ.get("HLB_CASE") == "tree_sites":
    # a site-granular partition of the tree (hemelb_b200/partition.py, inertial start + site stage):
    # ranks share blocks and have several neighbours each
    from hemelb_b200 import partition as P
    from tests.cases import geometry
    from tests.test_partition import collision_types
    full = geometry("tree")
    full_rank, quality = P.partition_sites(full, collision_types(full, Q), Q, nranks=world, initial="inertial", native=True)
    dom = DomainBuilder(full, Q, full_rank, world).domains[rank]
    assert dom.procs.shape[0] >= 1
else:
    # rank-local construction: own slab + one halo slice (what bench.py does per GPU)
    sub, sub_rank = G.cylinder_slab(radius, L, world, rank)
    dom = DomainBuilder(sub, Q, sub_rank, world).domains[rank]
    # the oracle needs a geometry it can run: the full one, but this process only drives its own rank
    full = G.cylinder_extruded(radius, L)
    full_rank = np.minimum((full.coords[:, 2].astype(np.int64) - 2) // (L // world), world - 1).astype(np.int32)
inlets, outlets = iolets_for(full, "NASH", "NASH")
odom = O.OracleDomains(full, Q, full_rank, world)
t = odom.tables(rank)
for k in ("neighbourIndices", "streamingIndices", "procs", "counts", "wallMask", "distanceToWall"):
    assert np.array_equal(np.asarray(t[k]), np.asarray(dom.tables()[k])), k
sim = O.OracleSim(odom, "LBGK", "BFL", tau=0.8, inlets=inlets, outlets=outlets)
N, S = dom.N, dom.totalSharedFs
f0 = anisotropic_f(N, Q, S, site_offset=7 * rank)
sim.set_f(f0, rank)
mid_total = int(dom.mid.sum())
for step in range(steps):
    off = mid_total
    for s in range(6):
        sim.stream_and_collide(s, off, int(dom.edge[s]), rank); off += int(dom.edge[s])
    off = 0
    for s in range(6):
        sim.stream_and_collide(s, off, int(dom.mid[s]), rank); off += int(dom.mid[s])
    fnew = sim.get_f(rank, 1)
    fold = sim.get_f(rank, 0)
    reqs = []
    bufs = []
    for (p, cnt, first) in dom.procs:
        send = torch.from_numpy(fnew[first:first + cnt].copy())
        recv = torch.zeros(int(cnt), dtype=torch.float64)
        reqs.append(dist.isend(send, int(p)))
        reqs.append(dist.irecv(recv, int(p)))
        bufs.append((int(first), int(cnt), recv, send))
    for r in reqs:
        r.wait()
    for first, cnt, recv, _ in bufs:
        fold[first:first + cnt] = recv.numpy()
    # CopyReceived
    fnew[dom.streamingIndices] = fold[N * Q + 1:N * Q + 1 + S]
    sim.set_f(fnew, rank, 1)
    sim.set_f(fold, rank, 0)
    off = mid_total
    for s in range(6):
        sim.post_step(s, off, int(dom.edge[s]), rank); off += int(dom.edge[s])
    off = 0
    for s in range(6):
        sim.post_step(s, off, int(dom.mid[s]), rank); off += int(dom.mid[s])
    # swap + time
    a, b = sim.get_f(rank, 0), sim.get_f(rank, 1)
    sim.set_f(b, rank, 0); sim.set_f(a, rank, 1)
    sim.set_time(step + 2)
mine = sim.get_f(rank, 0)[:N * Q]
# single-process reference run of all ranks
ref = O.OracleSim(O.OracleDomains(full, Q, full_rank, world), "LBGK", "BFL", tau=0.8, inlets=inlets, outlets=outlets)
for r in range(world):
    tr = odom.tables(r)
    ref.set_f(anisotropic_f(tr["N"], Q, tr["totalSharedFs"], site_offset=7 * r), r)
ref.step(steps)
assert np.array_equal(mine, ref.get_f(rank)[:N * Q]), "rank %%d differs" %% rank
dist.barrier()
print("rank", rank, "ok")
'''


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,case", [(2, "cylinder_slabs"), (3, "tree_sites")])
def test_two_process_halo_exchange(tmp_path, world, case):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    port = _free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   HLB_CASE=case)
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=240)
        except subprocess.TimeoutExpired:
            p.kill()
            out, _ = p.communicate()
        outs.append(out)
    for r, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, "rank %d failed:\n%s" % (r, out[-3000:])
        assert "ok" in out
