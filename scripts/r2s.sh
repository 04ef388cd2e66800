mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"collide_stream|gzs_links|post_links|copy_received|monitor|stability" -c 60 --csv --log-file gpurun_out/r2s_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:collide_stream -s 3 -c 1 -o gpurun_out/r2s_site_full python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r2s_ncu.log 2>&1
tail -2 gpurun_out/r2s_ncu.log
ncu --set full --clock-control none -k regex:post_links -s 3 -c 1 -o gpurun_out/r2s_post_full python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r2s_ncu2.log 2>&1
tail -2 gpurun_out/r2s_ncu2.log
