mkdir -p gpurun_out
HLB_GZS_OVERLAP=1 timeout 900 python -m pytest tests -m gpu -x -q -k "gzs or GZS or reference_inputs or four_cube or baseline_configs" --deselect tests/test_gpu_multi.py --deselect tests/test_zgpu_multi_next.py > gpurun_out/r2w_pytest_overlap.log 2>&1
tail -3 gpurun_out/r2w_pytest_overlap.log
timeout 900 python -m pytest tests -m gpu -x -q -k "gzs or GZS" --deselect tests/test_gpu_multi.py --deselect tests/test_zgpu_multi_next.py > gpurun_out/r2w_pytest.log 2>&1
tail -3 gpurun_out/r2w_pytest.log
for ov in 0 1 2; do
  HLB_GZS_OVERLAP=$ov timeout 400 python bench_tree.py --kernel MRT --wall GZS --inlet LADD --sites 1.1e8 > gpurun_out/r2w_configs3_ov$ov.json 2> gpurun_out/r2w_configs3_ov$ov.err
  HLB_GZS_OVERLAP=$ov timeout 400 python bench_tree.py --kernel LBGK --wall GZS --sites 1.1e8 > gpurun_out/r2w_lbgk_gzs_ov$ov.json 2> gpurun_out/r2w_lbgk_gzs_ov$ov.err
done
python - <<'PY'
import json
for n in ("configs3_ov0","configs3_ov1","configs3_ov2","lbgk_gzs_ov0","lbgk_gzs_ov1","lbgk_gzs_ov2"):
    try:
        l=json.loads(open("gpurun_out/r2w_%s.json"%n).read().strip().splitlines()[-1])
        print(n, "MLUPS %.0f whole-step frac %.3f ms %.3f"%(l["MLUPS"], l["whole_step_frac_of_hbm_roofline"], l["ms_per_step"]))
    except Exception as e:
        print(n, "failed", e)
PY
