for w in hidden200 solve; do
python bench.py --workload $w --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-gpu-reference --no-literal-loop 2>gpurun_out/r02_j2.err | tee gpurun_out/r02_j2_$w.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('$w', round(d['value'],1), {k: (round(v,4) if isinstance(v,float) else v) for k,v in r.items() if k in ('bound','achieved','peak','unit','frac','tensor_frac','hbm_frac','traffic','state_bytes_per_launch','peak_kind')})"
done; tail -3 gpurun_out/r02_j2.err
