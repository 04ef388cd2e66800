mkdir -p gpurun_out
for d in 20 40 60 80 100 125 150 200 250; do
  echo "prefetch $d"
  HLB_PREFETCH=$d python bench_tree.py --sites 1.1e8 --steps 30 2>/dev/null | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['MLUPS'], d['ms_per_step'], d['whole_step_frac_of_hbm_roofline'], d['rank0_bulk_kernel_frac'])"
done
