mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "monitor or stability or phase_api" 2>&1 | tail -3
python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-secondary 2>gpurun_out/r2j_bench.err | tee gpurun_out/r2j_bench.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['whole_step_frac'], d['e2e']['value'])"
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 12 --csv --log-file gpurun_out/r2j_cfg3_launches.csv python bench_tree.py --sites 1.1e8 --steps 3 --warmup 3 --kernel MRT --wall GZS --inlet LADD > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2j_cfg3_launches.csv')))
i0=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
for r in rows[i0+1:]:
    print(r[4][:90], r[8], r[-1])
PY
