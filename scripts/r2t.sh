mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "runs or four_cube or baseline_configs" > gpurun_out/r2t_pytest_first.log 2>&1
tail -3 gpurun_out/r2t_pytest_first.log
timeout 300 python bench.py --no-cpu-baseline --no-secondary --steps 100 > gpurun_out/r2t_bench_runs.json 2> gpurun_out/r2t_bench_runs.err
HLB_NBR_RUNS=0 timeout 300 python bench.py --no-cpu-baseline --no-secondary --steps 100 > gpurun_out/r2t_bench_planes.json 2> gpurun_out/r2t_bench_planes.err
python - <<'PY'
import json
for n in ("runs","planes"):
    try:
        l=json.loads(open("gpurun_out/r2t_bench_%s.json"%n).read().strip().splitlines()[-1])
        print(n, "MLUPS %.0f"%l["value"], "site kernel frac %.3f"%l["roofline"]["frac"], "whole %.3f"%l["roofline"]["whole_step_frac"], "e2e %.0f"%l["e2e"]["value"], l["roofline"].get("streaming_targets","")[:60])
    except Exception as e:
        print(n, "failed", e)
PY
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_multi.py --deselect tests/test_zgpu_multi_next.py > gpurun_out/r2t_pytest.log 2>&1
tail -5 gpurun_out/r2t_pytest.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:collide_stream -s 3 -c 1 -o gpurun_out/r2t_site_full python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r2t_ncu.log 2>&1
tail -2 gpurun_out/r2t_ncu.log
