# after the cheaper in-kernel monitors: whole -m gpu suite, contract line, ncu --set full of the site kernel (new source hash)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2final2_pytest.log 2>&1
tail -2 gpurun_out/r2final2_pytest.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:collide_stream -s 3 -c 1 -o gpurun_out/r2final2_site_full python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r2final2_ncu.log 2>&1
tail -1 gpurun_out/r2final2_ncu.log
timeout 400 python bench.py > gpurun_out/r2final2_bench.json 2> gpurun_out/r2final2_bench.err
python -c "
import json
l=json.loads(open('gpurun_out/r2final2_bench.json').read().strip().splitlines()[-1])
print('MLUPS %.0f site %.3f whole %.3f e2e %.0f traffic %s cyl %.0f'%(l['value'], l['roofline']['frac'], l['roofline']['whole_step_frac'], l['e2e']['value'], l['roofline']['traffic'], l['secondary']['value']))"
