# round 2, final build: GPU suite, smoke, the driver's bench command, launch list of the default step
python -m pytest tests -q -m gpu 2>&1 | tail -8 | tee gpurun_out/r02_gpu_suite_final2.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee gpurun_out/r02_smoke_final2.log
python bench.py 2> gpurun_out/r02_bench_final2.err | tee gpurun_out/r02_bench_final2.json | cut -c1-400
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-gpu-reference --no-literal-loop"
$CMD > /dev/null 2>&1 && ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches_default_final.csv $CMD > gpurun_out/ncu_default_final.log 2>&1
python tools/launch_list_summary.py gpurun_out/r02_launches_default_final.csv 1 | head -20 | tee gpurun_out/r02_launches_default_final_summary.txt
python bench.py --workload hidden200 --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-literal-loop 2>/dev/null | tee gpurun_out/r02_bench_hidden200_final2.json | cut -c1-200
