python -m pytest tests/test_gpu_parity.py tests/test_gpu_production_shapes.py tests/test_gpu_training.py tests/test_gpu_resident.py -x -q 2>&1 | tail -2
for hh in 200 72 800; do python bench.py --hidden $hh --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-gpu-reference 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('hidden $hh', round(d['value'],1), {k:round(v,3) for k,v in d['phase_ms_per_iteration'].items()}, d['clocks']['sm_mhz'])"; done
