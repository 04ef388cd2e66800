mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multi.py --deselect tests/test_zgpu_multi_next.py 2>&1 | tail -8 | tee gpurun_out/r2c_pytest.log
python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>gpurun_out/r2c_bench.err | tee gpurun_out/r2c_bench.json | cut -c1-1200
python bench_tree.py --sites 1.1e8 --steps 50 2>gpurun_out/r2c_tree.err | grep "^{" | tee gpurun_out/r2c_tree.json | cut -c1-900
ncu --set full --clock-control none --import-source on -k regex:collide_stream -s 8 -c 1 -o gpurun_out/r2c_tree_full python bench_tree.py --sites 3e7 --steps 2 --warmup 1 > gpurun_out/r2c_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/r2c_tree_launches.csv python bench_tree.py --sites 1.1e8 --steps 4 --warmup 3 > /dev/null 2>&1
tail -3 gpurun_out/r2c_ncu.log
