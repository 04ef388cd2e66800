#!/bin/bash
# GPU calls queued by the end of round 1 (the round's 180 GPU-minutes were spent before these changes
# were written).  Each block is one `gpurun` call; run them in this order and copy what should be
# judged from gpurun_out/ into profiles/.  Nothing here is executed by the tests or the bench.
set -euo pipefail
G=/usr/local/graft/bin/gpurun

# 1. everything that has a GPU test, one GPU (about 25 s of run time)
$G --timeout 300 -- 'python -m pytest tests -m gpu -q 2>&1 | tail -5; python -c "import __graft_entry__ as g; g.smoke()"'

# 2. the contract bench and its launch list (defaults are now 20 warm-up + 200 timed steps)
$G --timeout 600 -- 'mkdir -p gpurun_out; python bench.py 2>gpurun_out/r2_bench.err | tee gpurun_out/r2_bench.json | cut -c1-400;
  python bench.py --impl reference > gpurun_out/r2_bench_ref.json 2>/dev/null;
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 3 --warmup 3 > /dev/null 2>&1'

# 3. first multi-GPU run of: hlb_gpu_monitor_global, the site-granular NCCL cases, the 2-rank C++ harness
$G --gpus 2 --timeout 600 -- 'python -m pytest tests/test_gpu_multi.py tests/test_zgpu_multi_next.py tests/test_host_lbm.py -m gpu -q 2>&1 | tail -8'

# 4. the tree over 2 GPUs with each decomposition start (host-side figures: profiles/r01_partition_quality.jsonl)
for start in morton rcb inertial; do
  $G --gpus 2 --timeout 600 -- "python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench_tree.py --sites 2.2e8 --steps 30 --decomposition weighted --partition-start $start 2>/dev/null | grep '^{' | tee gpurun_out/r2_tree_n2_$start.json | cut -c1-600"
done

# 5. A/B of the per-step stream drain in upload_densities: make the
#    cudaStreamSynchronize conditional on the ring wrapping, rebuild, rerun 1 and 2.

# 6. the C++ policy classes driving full steps, wall-clock: write a case
#    with tests/test_host_lbm.py::write_case for a 1e7-site cylinder and run
#    HLB_HOST_TIMING=1 tests/_build/host_lbm_run case.bin out.bin
