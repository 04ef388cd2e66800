mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_multi.py --deselect tests/test_zgpu_multi_next.py > gpurun_out/r2v_pytest.log 2>&1
tail -3 gpurun_out/r2v_pytest.log
HLB_GZS_OVERLAP=1 timeout 1500 python -m pytest tests -m gpu -x -q -k "gzs or GZS or reference_inputs or four_cube or baseline_configs" --deselect tests/test_gpu_multi.py --deselect tests/test_zgpu_multi_next.py > gpurun_out/r2v_pytest_overlap.log 2>&1
tail -3 gpurun_out/r2v_pytest_overlap.log
for ov in 0 1; do
  HLB_GZS_OVERLAP=$ov timeout 400 python bench_tree.py --kernel MRT --wall GZS --inlet LADD --sites 1.1e8 > gpurun_out/r2v_configs3_ov$ov.json 2> gpurun_out/r2v_configs3_ov$ov.err
  HLB_GZS_OVERLAP=$ov timeout 400 python bench_tree.py --kernel LBGK --wall GZS --sites 1.1e8 > gpurun_out/r2v_lbgk_gzs_ov$ov.json 2> gpurun_out/r2v_lbgk_gzs_ov$ov.err
done
python - <<'PY'
import json
for n in ("configs3_ov0","configs3_ov1","lbgk_gzs_ov0","lbgk_gzs_ov1"):
    try:
        l=json.loads(open("gpurun_out/r2v_%s.json"%n).read().strip().splitlines()[-1])
        print(n, "MLUPS %.0f whole-step frac %.3f ms %.3f serial ms %.3f"%(l["MLUPS"], l["whole_step_frac_of_hbm_roofline"], l["ms_per_step"], l["serial_ms_per_step"]))
    except Exception as e:
        print(n, "failed", e)
PY
