mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench_tree.py "${@:3}" 2>gpurun_out/r2n_$2.err | grep "^{" | tee gpurun_out/r2n_$2.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$2', d['n_gpus'], d['sites'], d['MLUPS'], d['ms_per_step'], d['whole_step_frac_of_hbm_roofline'], d['rank0_boundary_fraction'], d['stable'], d.get('gzs_remote_links_rank0'), d['neighbours_per_rank'])"; tail -2 gpurun_out/r2n_$2.err | cut -c1-300; }
run 29581 cfg3_n8 --sites 8.8e8 --steps 50 --kernel MRT --wall GZS --inlet LADD
run 29582 cfg4_n8 --geometry sac --sites 8.0e8 --lattice 27 --kernel TRT --wall BFL --steps 50
run 29583 cfg4_n8_rough6 --geometry sac --roughness 6 --sites 8.0e8 --lattice 27 --kernel TRT --wall BFL --steps 50
