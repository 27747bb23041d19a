CMD="python bench.py --workload config5 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-gpu-reference --no-literal-loop"
$CMD > /dev/null 2>&1 && ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02_config5_launches.csv $CMD > gpurun_out/ncu_c5.log 2>&1
python tools/launch_list_summary.py gpurun_out/r02_config5_launches.csv 1 | head -24
