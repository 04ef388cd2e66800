mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_multi.py --deselect tests/test_zgpu_multi_next.py > gpurun_out/r2u_pytest.log 2>&1
tail -4 gpurun_out/r2u_pytest.log
for pf in 10240 20480 30720 51200 61440 81920; do
  HLB_PREFETCH=$pf timeout 200 python bench.py --no-cpu-baseline --no-secondary --steps 60 > gpurun_out/r2u_pf_$pf.json 2> gpurun_out/r2u_pf_$pf.err
  python -c "
import json,sys
l=json.loads(open('gpurun_out/r2u_pf_$pf.json').read().strip().splitlines()[-1])
print('prefetch $pf MLUPS %.0f site-kernel frac %.3f'%(l['value'], l['roofline']['frac']))" 2>&1 | tail -1
done
timeout 400 python bench_tree.py --kernel MRT --wall GZS --inlet LADD --sites 1.1e8 > gpurun_out/r2u_configs3.json 2> gpurun_out/r2u_configs3.err
timeout 400 python bench_tree.py --geometry sac --lattice 27 --kernel TRT --wall BFL --sites 1.1e8 > gpurun_out/r2u_configs4.json 2> gpurun_out/r2u_configs4.err
timeout 400 python bench_tree.py --kernel MRT --wall BFL --sites 1.1e8 > gpurun_out/r2u_mrt_bfl.json 2> gpurun_out/r2u_mrt_bfl.err
timeout 400 python bench_tree.py --lattice 15 --kernel LBGK --wall SBB --sites 1.1e8 > gpurun_out/r2u_q15_sbb.json 2> gpurun_out/r2u_q15_sbb.err
python - <<'PY'
import json
for n in ("configs3","configs4","mrt_bfl","q15_sbb"):
    try:
        l=json.loads(open("gpurun_out/r2u_%s.json"%n).read().strip().splitlines()[-1])
        print(n, "MLUPS %.0f whole-step frac %.3f bulk %.3f runs %s"%(l["MLUPS"], l["whole_step_frac_of_hbm_roofline"], l["rank0_bulk_kernel_frac"] or 0, l.get("rank0_target_runs")))
    except Exception as e:
        print(n, "failed", e)
PY
timeout 600 python bench.py > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err
tail -c 400 gpurun_out/r2u_bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"collide_stream|gzs_links|post_links|copy_received|monitor|stability" -c 60 --csv --log-file gpurun_out/r2u_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > /dev/null 2>&1
