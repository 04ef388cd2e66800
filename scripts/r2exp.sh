mkdir -p gpurun_out
HLB_LIB=$PWD/experiments/libhemelb_b200_c10.so timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "baseline_configs or runs" > gpurun_out/r2exp_pytest.log 2>&1
tail -2 gpurun_out/r2exp_pytest.log
for lib in experiments/libhemelb_b200_c10.so hemelb_b200/libhemelb_b200.so experiments/libhemelb_b200_c10.so hemelb_b200/libhemelb_b200.so; do
HLB_LIB=$PWD/$lib timeout 300 python bench.py --no-cpu-baseline --steps 100 2>/dev/null | python -c "
import sys,json
l=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$lib MLUPS %.0f site %.3f whole %.3f cyl %.0f'%(l['value'], l['roofline']['frac'], l['roofline']['whole_step_frac'], l['secondary']['value']))"
done
