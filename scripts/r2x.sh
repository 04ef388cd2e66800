mkdir -p gpurun_out
timeout 300 python bench.py --no-cpu-baseline --no-secondary --steps 100 > gpurun_out/r2x_bench_quick.json 2> gpurun_out/r2x_bench_quick.err
python -c "
import json
l=json.loads(open('gpurun_out/r2x_bench_quick.json').read().strip().splitlines()[-1])
print('quick MLUPS %.0f site-kernel frac %.3f whole %.3f e2e %.0f'%(l['value'], l['roofline']['frac'], l['roofline']['whole_step_frac'], l['e2e']['value']))"
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_multi.py --deselect tests/test_zgpu_multi_next.py > gpurun_out/r2x_pytest.log 2>&1
tail -3 gpurun_out/r2x_pytest.log
timeout 600 python bench.py > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err
tail -c 300 gpurun_out/r2x_bench.json
timeout 400 ncu --set full --clock-control none --import-source on -k regex:collide_stream -s 3 -c 1 -o gpurun_out/r2x_site_full python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r2x_ncu.log 2>&1
tail -2 gpurun_out/r2x_ncu.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"collide_stream|gzs_links|post_links|copy_received|monitor|stability" -c 60 --csv --log-file gpurun_out/r2x_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > /dev/null 2>&1
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2x_bench_reference.json 2> gpurun_out/r2x_bench_reference.err
tail -c 300 gpurun_out/r2x_bench_reference.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
