CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-gpu-reference"
$CMD > gpurun_out/r02_h_plain.json 2> gpurun_out/r02_h_plain.err && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_launches_default.csv $CMD > gpurun_out/r02_h_ncu1.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gates_tc_pair -s 20 -c 1 -o gpurun_out/r02_gates_full $CMD > gpurun_out/r02_h_ncu2.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:ruiz_chain -s 2 -c 3 -o gpurun_out/r02_ruiz_chain_full $CMD > gpurun_out/r02_h_ncu3.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:kkt_pass -s 10 -c 2 -o gpurun_out/r02_kkt_full $CMD > gpurun_out/r02_h_ncu4.log 2>&1
tail -2 gpurun_out/r02_h_ncu4.log; ls -la gpurun_out/*.ncu-rep | tail -5; cut -c1-300 gpurun_out/r02_h_plain.json
