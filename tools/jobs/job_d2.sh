# round 2, last build: BASELINE config 4 (batch-size sweep at the headline shape) on one GPU
: > gpurun_out/r02_batch_sweep_final.jsonl
for b in 4 16 64 256 1024 2048; do
  timeout 200 python bench.py --batch $b --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-gpu-reference --no-literal-loop 2>> gpurun_out/r02_sweep.err | tee -a gpurun_out/r02_batch_sweep_final.jsonl | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('batch $b', round(d['value'],1), {k:round(v,3) for k,v in d['phase_ms_per_iteration'].items()}, d['clocks']['sm_mhz'], round(d['hbm_roofline_frac_whole_solve'],3))"
done
