python bench.py --impl reference 2>gpurun_out/r2r_ref.err | tee gpurun_out/r2r_bench_ref.json | cut -c1-1200
tail -3 gpurun_out/r2r_ref.err
