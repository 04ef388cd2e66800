mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"collide_stream|gzs_links|post_links|copy_received" -s 4 -c 8 --csv --log-file gpurun_out/r2k_cfg3_launches.csv python bench_tree.py --sites 1.1e8 --steps 3 --warmup 3 --kernel MRT --wall GZS --inlet LADD > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2k_cfg3_launches.csv')))
i0=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
for r in rows[i0+1:]:
    print(r[4][:90], r[8], r[-1])
PY
python bench_tree.py --sites 1.1e8 --steps 30 --kernel MRT --wall BFL --inlet NASH 2>/dev/null | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('mrt_bfl', d['sites'], d['MLUPS'], d['ms_per_step'], d['whole_step_frac_of_hbm_roofline'])"
python bench_tree.py --sites 1.1e8 --steps 30 --kernel LBGK --wall GZS --inlet LADD 2>/dev/null | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('lbgk_gzs', d['sites'], d['MLUPS'], d['ms_per_step'], d['whole_step_frac_of_hbm_roofline'])"
