# last check of the round on one GPU: the whole -m gpu suite, the contract line, the reference arm, smoke
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2final_pytest.log 2>&1
tail -3 gpurun_out/r2final_pytest.log
timeout 600 python bench.py > gpurun_out/r2final_bench.json 2> gpurun_out/r2final_bench.err
python -c "
import json
l=json.loads(open('gpurun_out/r2final_bench.json').read().strip().splitlines()[-1])
print('MLUPS %.0f site %.3f whole %.3f e2e %.0f traffic %s cyl %.0f cpu %.1f launches %d'%(l['value'], l['roofline']['frac'], l['roofline']['whole_step_frac'], l['e2e']['value'], l['roofline']['traffic'], l['secondary']['value'], l['cpu_baseline']['value'], l['gpu_launches']))"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 | tail -c 200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
