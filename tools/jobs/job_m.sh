CMD="python bench.py --workload train --batch 32 --steps 1 --warmup 3 --tl 10"
$CMD > gpurun_out/r02_m_plain.json 2> gpurun_out/r02_m.err && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_train_launches_b32_tl10.csv $CMD > gpurun_out/r02_m_ncu.log 2>&1
tail -2 gpurun_out/r02_m_ncu.log
python tools/launch_list_summary.py gpurun_out/r02_train_launches_b32_tl10.csv | head -40
