mkdir -p gpurun_out
nvidia-smi -L | wc -l
python -m pytest tests/test_gpu_multi.py tests/test_zgpu_multi_next.py tests/test_host_lbm.py -m gpu -q -rs 2>&1 | tail -8 | tee gpurun_out/r2h_multi.log
for n in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2956$n bench.py --gpus $n --steps 100 --warmup 10 2>gpurun_out/r2h_bench$n.err | tee gpurun_out/r2h_bench_n$n.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['whole_step_frac'], d['e2e']['value'], d['sites']['halo_doubles_per_rank'], d['sites']['neighbours_per_rank'], d['secondary'])"
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 8 --steps 100 --warmup 10 --decomposition weighted --partition-start inertial --no-secondary 2>gpurun_out/r2h_bench8w.err | tee gpurun_out/r2h_bench_n8_weighted.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('weighted', d['value'], d['ms_per_step'], d['roofline']['whole_step_frac'], d['sites']['halo_doubles_per_rank'], d['sites']['neighbours_per_rank'], d['sites']['per_rank'])"
