"""Multi-GPU cases written after the round's GPU budget was spent: they have not had their first
run yet, so they live in a module that sorts after every verified GPU test (the suite runs with -x).
Same worker as tests/test_gpu_multi.py; needs >= 2 GPUs, skipped otherwise.

* the NCCL halo of a site-granular tree partition (inertial start + site stage: ranks share blocks
  and have several neighbours each) against the oracle's emulated-rank run;
* hlb_gpu_monitor_global (one ncclAllReduce) against the per-rank monitors combined on the host."""
import pytest

from tests.test_gpu_multi import run_workers

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 4])
def test_nccl_halo_of_a_site_granular_partition(tmp_path, world):
    run_workers(tmp_path, world, "tree_sites")


def test_global_monitor_over_nccl(tmp_path):
    run_workers(tmp_path, 2, "cylinder_slabs", check_monitor=True)
