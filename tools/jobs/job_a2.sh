# round 2, re-entry: Ruiz column-partial fold -- parity tests, same-box A/B (development build), config-5 and default bench lines
python -m pytest tests/test_gpu_parity.py -q -m gpu -k "ruiz" 2>&1 | tail -15 | tee gpurun_out/r02_ruiz_fold_tests.log
python -m pytest tests/test_gpu_production_shapes.py -q -m gpu -k "config5 or unmodified_reference" 2>&1 | tail -3 | tee -a gpurun_out/r02_ruiz_fold_tests.log
DEVLIB=$PWD/i-admm-lstm_b200/iadmm_b200/libiadmm_b200_dev.so
: > gpurun_out/r02_ruiz_fold_ab.jsonl
for cfg in "256 1000" "24 5000" "64 2000"; do
  set -- $cfg
  for f in 1 0; do
    RZ_B=$1 RZ_N=$2 IADMM_B200_LIB=$DEVLIB IADMM_RUIZ_FOLD=$f python tools/ruiz_ab.py 2>&1 | tail -1 | tee -a gpurun_out/r02_ruiz_fold_ab.jsonl
  done
done
python bench.py --workload config5 --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-literal-loop 2> gpurun_out/r02_c5.err | tee gpurun_out/r02_bench_config5_v5.json
