timeout 300 python -m pytest tests/test_gpu_training.py tests/test_gpu_production_shapes.py -x -q -k "window or gradient or recomputed" 2>&1 | tail -4
DEV=$PWD/i-admm-lstm_b200/iadmm_b200/libiadmm_b200_dev.so
rm -f gpurun_out/r02_gemm_pair_ab.jsonl
for b in 32 8; do
  IADMM_B200_LIB=$DEV IADMM_GEMM_PAIR=0 timeout 300 python bench.py --workload train --batch $b --steps 3 --warmup 3 >> gpurun_out/r02_gemm_pair_ab.jsonl 2>> gpurun_out/r02_l.err
  IADMM_B200_LIB=$DEV timeout 300 python bench.py --workload train --batch $b --steps 3 --warmup 3 >> gpurun_out/r02_gemm_pair_ab.jsonl 2>> gpurun_out/r02_l.err
done
IADMM_B200_LIB=$DEV IADMM_GEMM_PAIR=0 timeout 300 python bench.py --workload train --batch 2 --graph --steps 5 --warmup 3 >> gpurun_out/r02_gemm_pair_ab.jsonl 2>> gpurun_out/r02_l.err
IADMM_B200_LIB=$DEV timeout 300 python bench.py --workload train --batch 2 --graph --steps 5 --warmup 3 >> gpurun_out/r02_gemm_pair_ab.jsonl 2>> gpurun_out/r02_l.err
tail -3 gpurun_out/r02_l.err
python -c "
import json
for i,l in enumerate(open('gpurun_out/r02_gemm_pair_ab.jsonl')):
    d=json.loads(l); print('single' if i%2==0 else 'pair  ', d['config']['batch_per_gpu'], round(d['value'],1), round(d['ms_per_step'],1), {k:round(v,3) for k,v in d['phase_ms_per_iteration'].items()}, round(d['roofline']['frac'],3), d['config']['loss'], d['clocks']['sm_mhz'])
"
