mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multi.py --deselect tests/test_zgpu_multi_next.py 2>&1 | tail -12 | tee gpurun_out/r2d_pytest.log
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "baseline_configs_small or phase_api" 2>&1 | tail -12 | tee gpurun_out/r2d_sanitizer.log
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>gpurun_out/r2d_bench.err | tee gpurun_out/r2d_bench.json | cut -c1-3000
HLB_TMA=0 timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>gpurun_out/r2d_bench_notma.err | tee gpurun_out/r2d_bench_notma.json | cut -c1-600
timeout 600 ncu --set full --clock-control none --import-source on -k regex:site_tma -s 2 -c 1 -o gpurun_out/r2d_tree_full python bench_tree.py --sites 3e7 --steps 2 --warmup 1 > gpurun_out/r2d_ncu.log 2>&1
tail -3 gpurun_out/r2d_ncu.log
