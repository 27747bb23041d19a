CMD="python bench.py --hidden 208 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-gpu-reference"
$CMD > gpurun_out/r02_q_plain.json 2> gpurun_out/r02_q.err && \
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gates_tc_pair -s 20 -c 1 -o gpurun_out/r02_gates_h208_full $CMD > gpurun_out/r02_q_ncu.log 2>&1
tail -2 gpurun_out/r02_q_ncu.log
