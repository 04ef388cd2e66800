mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "gzs or GZS or four_cube or baseline_configs or reference_inputs" --deselect tests/test_gpu_multi.py --deselect tests/test_zgpu_multi_next.py > gpurun_out/r2ad_pytest.log 2>&1
tail -2 gpurun_out/r2ad_pytest.log
timeout 400 python bench_tree.py --kernel MRT --wall GZS --inlet LADD --sites 1.1e8 > gpurun_out/r2ad_configs3.json 2> gpurun_out/r2ad_configs3.err
timeout 400 python bench_tree.py --kernel LBGK --wall GZS --sites 1.1e8 > gpurun_out/r2ad_lbgk_gzs.json 2> gpurun_out/r2ad_lbgk_gzs.err
python -c "
import json
for n in ('configs3','lbgk_gzs'):
    l=json.loads(open('gpurun_out/r2ad_%s.json'%n).read().strip().splitlines()[-1]); print(n, l['MLUPS'], l['whole_step_frac_of_hbm_roofline'], l['ms_per_step'])"
