mkdir -p gpurun_out
python bench.py > gpurun_out/bench_il.json 2> gpurun_out/bench_il.err; tail -2 gpurun_out/bench_il.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench_il.err
python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_short.json 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 800 --csv --log-file gpurun_out/launches_il.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
PROBE_REPS=1 PROBE_K=20 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gates_tc_pair -s 10 -c 1 -o gpurun_out/gates_il_v2 -f python tools/gate_probe.py > gpurun_out/ncu_il2.log 2>&1
PROBE_REPS=1 PROBE_K=20 timeout 600 ncu --set full --clock-control none --import-source on -k regex:kkt_pass -s 30 -c 2 -o gpurun_out/kkt_il_v2 -f python tools/gate_probe.py > gpurun_out/ncu_kkt2.log 2>&1
ls -la gpurun_out; cat gpurun_out/bench_il.json
