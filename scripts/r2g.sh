mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py tests/test_zgpu_multi_next.py tests/test_host_lbm.py -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r2g_multi.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 100 --warmup 10 2>gpurun_out/r2g_bench2.err | tee gpurun_out/r2g_bench_n2.json | cut -c1-1500
tail -5 gpurun_out/r2g_bench2.err
