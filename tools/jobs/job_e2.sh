# round 2, last build: hidden_dim 200 and config 5 lines carrying hbm_roofline_frac_whole_solve
for w in hidden200 config5; do
  python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-literal-loop 2>> gpurun_out/r02_e2.err | tee gpurun_out/r02_bench_${w}_final3.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$w', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), {k:round(v,3) for k,v in d['phase_ms_per_iteration'].items()}, d['clocks']['sm_mhz'], 'iterations-only', round(d['hbm_roofline_frac_whole_path'],3), 'whole solve', round(d['hbm_roofline_frac_whole_solve'],3))"
done
