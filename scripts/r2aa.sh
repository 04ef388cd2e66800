mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_multi.py --deselect tests/test_zgpu_multi_next.py > gpurun_out/r2aa_pytest.log 2>&1
tail -3 gpurun_out/r2aa_pytest.log
timeout 300 python bench_host_cxx.py --steps 100 > gpurun_out/r2aa_host_cxx.json 2> gpurun_out/r2aa_host_cxx.err; cat gpurun_out/r2aa_host_cxx.json
