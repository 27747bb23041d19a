python bench.py 2> gpurun_out/r02_bench_final3.err | tee gpurun_out/r02_bench_final3.json | cut -c1-160
