mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r2q_pytest.log
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py 2>gpurun_out/r2q_bench.err | tee gpurun_out/r2q_bench.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['whole_step_frac'], d['e2e']['value'], d['secondary']['value'], d['cpu_baseline'])"
python bench.py --impl reference 2>gpurun_out/r2q_ref.err | tee gpurun_out/r2q_bench_ref.json | cut -c1-700
tail -3 gpurun_out/r2q_ref.err
