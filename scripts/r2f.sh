mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q -x 2>&1 | tail -2
python bench_tree.py --sites 1.1e8 --steps 50 2>/dev/null | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('tree', d['MLUPS'], d['ms_per_step'], d['whole_step_frac_of_hbm_roofline'], d['rank0_bulk_kernel_frac'], d['rank0_bulk_share_of_step'])"
python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2>gpurun_out/r2f_bench.err | tee gpurun_out/r2f_bench.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['whole_step_frac'], d['e2e']['value'], d['secondary'])"
