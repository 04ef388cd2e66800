mkdir -p gpurun_out
python bench_host_cxx.py 2>gpurun_out/r2o_cxx.err | tee gpurun_out/r2o_host_cxx.json | cut -c1-900
tail -3 gpurun_out/r2o_cxx.err
python bench_tree.py --sites 1.1e8 --steps 30 --kernel MRT --wall GZS --inlet LADD 2>/dev/null | grep "^{" > gpurun_out/r2o_cfg3_n1.json
python bench_tree.py --geometry sac --sites 1.0e8 --lattice 27 --kernel TRT --wall BFL --steps 30 2>/dev/null | grep "^{" > gpurun_out/r2o_cfg4_n1.json
python bench_tree.py --geometry sac --roughness 6 --sites 1.0e8 --lattice 27 --kernel TRT --wall BFL --steps 30 2>/dev/null | grep "^{" > gpurun_out/r2o_cfg4_n1_rough6.json
python bench_tree.py --sites 1.1e8 --steps 30 --kernel MRT --wall BFL 2>/dev/null | grep "^{" > gpurun_out/r2o_mrt_bfl_n1.json
python bench_tree.py --sites 1.1e8 --steps 30 --lattice 15 --wall SBB 2>/dev/null | grep "^{" > gpurun_out/r2o_q15_sbb_n1.json
for f in cfg3_n1 cfg4_n1 cfg4_n1_rough6 mrt_bfl_n1 q15_sbb_n1; do python -c "import sys,json; d=json.load(open('gpurun_out/r2o_$f.json')); print('$f', d['sites'], d['MLUPS'], d['ms_per_step'], d['whole_step_frac_of_hbm_roofline'], d['rank0_boundary_fraction'], d['stable'])"; done
