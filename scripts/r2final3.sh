# default size raised to ~1.11e8 sites per GPU (8.9e8 on eight): contract line and ncu capture at the new size
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:collide_stream -s 3 -c 1 -o gpurun_out/r2final3_site_full python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r2final3_ncu.log 2>&1
tail -1 gpurun_out/r2final3_ncu.log
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/r2final3_bench.json 2> gpurun_out/r2final3_bench.err
python -c "
import json
l=json.loads(open('gpurun_out/r2final3_bench.json').read().strip().splitlines()[-1])
print('sites %d MLUPS %.0f site %.3f whole %.3f e2e %.0f cyl %.0f alg %d'%(l['sites']['global'], l['value'], l['roofline']['frac'], l['roofline']['whole_step_frac'], l['e2e']['value'], l['secondary']['value'], l['roofline']['algorithmic_bytes_per_launch']))"
