# final 1-GPU validation of the round: suite, contract line, ncu captures, reference arm, smoke
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_multi.py --deselect tests/test_zgpu_multi_next.py > gpurun_out/r2y_pytest.log 2>&1
tail -3 gpurun_out/r2y_pytest.log
timeout 600 python bench.py > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err
tail -c 300 gpurun_out/r2y_bench.json
timeout 400 ncu --set full --clock-control none --import-source on -k regex:collide_stream -s 3 -c 1 -o gpurun_out/r2y_site_full python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r2y_ncu.log 2>&1
tail -2 gpurun_out/r2y_ncu.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"collide_stream|gzs_links|post_links|copy_received|monitor|stability" -c 60 --csv --log-file gpurun_out/r2y_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > /dev/null 2>&1
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2y_bench_reference.json 2> gpurun_out/r2y_bench_reference.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 600 python bench_host_cxx.py > gpurun_out/r2y_host_cxx.json 2> gpurun_out/r2y_host_cxx.err
cat gpurun_out/r2y_host_cxx.json
