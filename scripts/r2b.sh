mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multi.py --deselect tests/test_zgpu_multi_next.py 2>&1 | tail -25 | tee gpurun_out/r2b_pytest.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "baseline_configs_small or single_ranges or phase_api or gzs_site_halo or schedules" 2>&1 | tail -15 | tee gpurun_out/r2b_sanitizer.log
python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>gpurun_out/r2b_bench.err | tee gpurun_out/r2b_bench.json | cut -c1-1500
python bench_tree.py --sites 1.1e8 --steps 50 2>gpurun_out/r2b_tree.err | grep "^{" | tee gpurun_out/r2b_tree.json | cut -c1-1200
