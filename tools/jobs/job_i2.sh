# last build: ncu --set full of the 16-warp gate kernel at hidden_dim 200 (after the same command exited 0 without ncu)
CMD="python bench.py --workload hidden200 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-gpu-reference --no-literal-loop"
$CMD > gpurun_out/r02_i2_plain.json 2> gpurun_out/r02_i2.err && \
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gates_tc_pair -s 20 -c 1 -o gpurun_out/r02_gates_h200_final_full $CMD > gpurun_out/r02_i2_ncu.log 2>&1
tail -2 gpurun_out/r02_i2_ncu.log; ls -la gpurun_out/r02_gates_h200_final_full.ncu-rep
