"""Golden extraction / checkpoint files written by the compiled reference
(tests/golden/make_golden_xtr.py): the oracle (restated streamers + oracle/xtr.py) must reproduce
them byte for byte wherever it runs -- no reference checkout needed."""
import os

import numpy as np
import pytest

from oracle import xtr as X
from tests.xtr_cases import CASES, DT, DX, ORIGIN, REF_PRESSURE, RHO, make_sim, rank_data, xfields

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    with open(os.path.join(GOLD, "xtr_%s.xtr" % name), "rb") as a, open(os.path.join(GOLD, "xtr_%s.off" % name), "rb") as b:
        return a.read(), b.read()


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_reference_files(name):
    case = CASES[name]
    sim, T = make_sim("oracle", case)
    conv = X.UnitConverter(DT, DX, ORIGIN, RHO, REF_PRESSURE)
    want_xtr, want_off = golden(name)
    got = None
    t = case["steps"]
    for more in case["writes"]:
        sim.step(more)
        t += more
        po = X.PropertyOutput(xfields(case["fields"]), case["selector"], case["params"], conv, case["Q"],
                              rank_data(sim, T, case["Q"]))
        if got is None:
            got = po.header
            assert po.offset_file() == want_off
        if t % case["frequency"] == 0:
            got += po.record(t)
    assert got == want_xtr


def test_checkpoint_golden_loads():
    case = CASES["cylinder_checkpoint_r2"]
    sim, T = make_sim("oracle", case)
    xb, ob = golden("cylinder_checkpoint_r2")
    t, f = X.load_checkpoint(xb, ob, case["Q"], [x["globalCoords"] for x in T], 3)
    assert t == 3
    for r in range(case["R"]):
        assert np.array_equal(f[r].ravel(), sim.get_f(r)[:T[r]["N"] * case["Q"]])
