"""Host logic of the extraction / checkpoint mirror (hemelb_b200/extraction.py) on CPU: the file
layout across ranks over a real 2-process gloo group, the offset-file naming, and the header
validation of the checkpoint reader (every check LocalDistributionInput.cc:167-308 makes happens
before anything touches the GPU)."""
import os
import socket
import struct
import subprocess
import sys

import numpy as np
import pytest

from oracle import xtr as X
from hemelb_b200.capi import HlbError
from hemelb_b200.extraction import GpuLocalDistributionInput, SingleComm, extraction_to_offset, write_layout
from tests.xtr_cases import CASES, DT, DX, ORIGIN, REF_PRESSURE, RHO, make_sim, rank_data, xfields

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_offset_file_name():
    assert extraction_to_offset("results/Extracted/whole.xtr") == "results/Extracted/whole.off"
    assert extraction_to_offset("a.b/c.d.xtr") == "a.b/c.d.off"
    with pytest.raises(HlbError, match="Cannot split extension"):
        extraction_to_offset("noextension")


def test_single_rank_layout_matches_oracle():
    case = CASES["four_cube_all"]
    sim, T = make_sim("oracle", case)
    po = X.PropertyOutput(xfields(case["fields"]), "whole", (), X.UnitConverter(DT, DX, ORIGIN, RHO, REF_PRESSURE), 15,
                          rank_data(sim, T, 15))
    lay = write_layout(SingleComm(), po.local_counts[0], po.site_len, len(po.header))
    assert lay["local_write_start"] == po.local_start[0] == len(po.header)
    assert lay["local_data_write_length"] == po.local_len[0]
    assert lay["global_data_write_length"] == po.global_len


WORKER = r'''
import os, sys, json
import torch.distributed as dist
sys.path.insert(0, %(root)r)
from hemelb_b200.extraction import TorchComm, write_layout
dist.init_process_group("gloo")
c = TorchComm()
counts = json.loads(%(counts)r)
lay = write_layout(c, counts[c.rank], %(site_len)d, %(header)d)
lay["bcast"] = c.broadcast("from-io" if c.rank == 0 else None, 0)
lay["scatter"] = c.scatter([[10, 11], [20, 21]] if c.rank == 0 else None, 0)
c.barrier()
open(os.path.join(%(out)r, "layout%%d.json" %% c.rank), "w").write(json.dumps(lay))
dist.destroy_process_group()
'''


def test_two_rank_layout_over_gloo(tmp_path):
    """Scan / AllReduce / Broadcast / Scatter over torch.distributed give every rank the offsets
    the reference's MPI calls would (compared with the oracle's 2-rank layout)."""
    case = CASES["cylinder_surface_r2"]
    sim, T = make_sim("oracle", case)
    po = X.PropertyOutput(xfields(case["fields"]), "surface", (), X.UnitConverter(DT, DX, ORIGIN, RHO, REF_PRESSURE), 19,
                          rank_data(sim, T, 19))
    import json
    script = tmp_path / "w.py"
    script.write_text(WORKER % dict(root=ROOT, counts=json.dumps(po.local_counts), site_len=po.site_len, header=len(po.header),
                                      out=str(tmp_path)))
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", str(port), str(script)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    seen = {rk: json.loads((tmp_path / ("layout%d.json" % rk)).read_text()) for rk in (0, 1)}
    for rk in (0, 1):
        assert seen[rk]["local_write_start"] == po.local_start[rk]
        assert seen[rk]["local_data_write_length"] == po.local_len[rk]
        assert seen[rk]["global_site_count"] == po.global_count
        assert seen[rk]["global_data_write_length"] == po.global_len
        assert seen[rk]["bcast"] == "from-io"
        assert seen[rk]["scatter"] == [[10, 11], [20, 21]][rk]


class _NoGpu:
    Q = 19

    @property
    def domain(self):
        raise AssertionError("the checkpoint reader must fail before it reaches the device")

    h = None


def _golden(name):
    g = os.path.join(ROOT, "tests", "golden")
    return open(os.path.join(g, name + ".xtr"), "rb").read(), open(os.path.join(g, name + ".off"), "rb").read()


def _write(tmp_path, name, xb, ob):
    (tmp_path / (name + ".xtr")).write_bytes(xb)
    (tmp_path / (name + ".off")).write_bytes(ob)
    return tmp_path / (name + ".xtr")


def test_checkpoint_reader_rejects_before_the_device(tmp_path):
    xb, ob = _golden("xtr_cylinder_checkpoint_r2")
    one_rank_off = struct.pack(">IIIi", 0x686C6221, 0x6F666604, 1, 1) + struct.pack(">QQ", 92, 92 + (len(xb) - 92) // 3)
    cases = [
        ("magic", b"\0\0\0\0" + xb[4:], one_rank_off, "does not start with the HemeLB magic number"),
        ("xmagic", xb[:4] + b"\0\0\0\1" + xb[8:], one_rank_off, "does not have the extraction magic number"),
        ("version", xb[:8] + struct.pack(">I", 4) + xb[12:], one_rank_off, "Version number incorrect. Supported: 5 Input: 4"),
        ("nfields", xb[:52] + struct.pack(">I", 2) + xb[56:], one_rank_off, "exactly one field"),
        ("fhlen", xb[:56] + struct.pack(">I", 28) + xb[60:], one_rank_off, "must be 32 B long, but is 28 B"),
        ("name", xb[:64] + b"distributionz" + xb[77:], one_rank_off, "field named 'distributions', but has 'distributionz'"),
        ("type", xb[:84] + struct.pack(">I", 0) + xb[88:], one_rank_off, "wrong data type"),
        ("noff", xb[:88] + struct.pack(">I", 1) + xb[92:], one_rank_off, "should not have offsets"),
        ("offmagic", xb, one_rank_off[:4] + b"\0\0\0\0" + one_rank_off[8:], "does not have the offset magic number"),
        ("ranks", xb, ob, "wrong number of MPI ranks. Running with: 1 Input: 2"),
        ("length", xb + b"\0" * 8, one_rank_off, "not consistent with integer number of checkpoints"),
    ]
    for name, x, o, msg in cases:
        with pytest.raises(HlbError, match=msg):
            GpuLocalDistributionInput(_write(tmp_path, name, x, o)).load_distribution(_NoGpu())
    q15 = _NoGpu()
    q15.Q = 15
    with pytest.raises(HlbError, match="contains 19 distributions but this build of HemeLB requires 15"):
        GpuLocalDistributionInput(_write(tmp_path, "q", xb, one_rank_off)).load_distribution(q15)
    # time steps 3, 5, 7 are in the file; 4 is absent, and the reference's search also misses 5
    # (it ends on a probe of another record, LocalDistributionInput.cc:79-96) -- same here
    for t in (4, 5):
        with pytest.raises(HlbError, match="Target timestep %d not found" % t):
            GpuLocalDistributionInput(_write(tmp_path, "t%d" % t, xb, one_rank_off)).load_distribution(_NoGpu(), t)
