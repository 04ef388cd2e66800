# final multi-GPU validation: parity suites on however many GPUs the box has, then the contract line at N = all, N/2
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "gpus $N"
timeout 1500 python -m pytest tests/test_gpu_multi.py tests/test_zgpu_multi_next.py tests/test_host_lbm.py -m gpu -q -rs 2>&1 | tail -8 | tee gpurun_out/r2z_multi_${N}gpu.log
for n in $N $((N/2)); do
  if [ $n -ge 2 ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2956$n bench.py --gpus $n --steps 100 --warmup 10 2>gpurun_out/r2z_bench$n.err | tee gpurun_out/r2z_bench_n$n.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['whole_step_frac'], d['e2e']['value'], d['secondary'])"
  fi
done
if [ $N -ge 8 ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581 bench_tree.py --kernel MRT --wall GZS --inlet LADD --sites 8.8e8 > gpurun_out/r2z_configs3_8gpu.json 2> gpurun_out/r2z_configs3_8gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29582 bench_tree.py --geometry sac --lattice 27 --kernel TRT --wall BFL --sites 8.8e8 > gpurun_out/r2z_configs4_8gpu.json 2> gpurun_out/r2z_configs4_8gpu.err
python -c "
import json
for n in ('configs3','configs4'):
    try:
        l=json.loads(open('gpurun_out/r2z_%s_8gpu.json'%n).read().strip().splitlines()[-1]); print(n, l['MLUPS'], l['whole_step_frac_of_hbm_roofline'])
    except Exception as e: print(n,'failed',e)
"
fi
