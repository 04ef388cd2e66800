mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_bench_contract.py -m gpu -x -q -k "monitor or bench or error" > gpurun_out/r2ae_pytest.log 2>&1
tail -3 gpurun_out/r2ae_pytest.log
timeout 600 python bench.py --no-cpu-baseline --no-secondary --steps 100 > gpurun_out/r2ae_bench.json 2> gpurun_out/r2ae_bench.err
python -c "
import json
l=json.loads(open('gpurun_out/r2ae_bench.json').read().strip().splitlines()[-1])
print('MLUPS %.0f e2e %.0f blocking %.0f'%(l['value'], l['e2e']['value'], l['e2e']['value_with_blocking_monitor']))"
tail -3 gpurun_out/r2ae_bench.err
