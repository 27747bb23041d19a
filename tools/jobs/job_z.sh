nvidia-smi --query-gpu=name,temperature.gpu,power.limit,clocks.max.sm,clocks.max.mem --format=csv
python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-literal-loop 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('default', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['phase_ms_per_iteration'], d['clocks'])"
python bench.py --workload hidden200 --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-literal-loop 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('hidden200', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['phase_ms_per_iteration'], d['clocks'])"
