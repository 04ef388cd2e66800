mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on -k regex:gzs_links -s 2 -c 1 -o gpurun_out/r2ac_gzs_full python bench_tree.py --kernel MRT --wall GZS --inlet LADD --sites 1.1e8 --steps 3 --warmup 3 > gpurun_out/r2ac_ncu.log 2>&1
tail -2 gpurun_out/r2ac_ncu.log
