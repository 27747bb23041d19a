python -m pytest tests -m gpu -q 2>&1 | tail -4
rm -f gpurun_out/r02_batch_sweep.jsonl
for b in 64 256 512 1024; do python bench.py --batch $b --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-gpu-reference >> gpurun_out/r02_batch_sweep.jsonl 2>> gpurun_out/r02_i.err; done
python bench.py --workload sparse --family SVM --steps 3 --no-cpu-baseline >> gpurun_out/r02_batch_sweep.jsonl 2>> gpurun_out/r02_i.err
python bench.py --workload solve --sparse auto --steps 5 --no-cpu-baseline --no-gpu-reference >> gpurun_out/r02_batch_sweep.jsonl 2>> gpurun_out/r02_i.err
tail -3 gpurun_out/r02_i.err
python -c "
import json
for l in open('gpurun_out/r02_batch_sweep.jsonl'):
    d=json.loads(l); print(d['config']['workload'][:20], d['config']['matrix_form'][:6], d['config']['batch_per_gpu'], round(d['value'],1), {k:round(v,3) for k,v in d['phase_ms_per_iteration'].items()}, round(d['roofline']['frac'],3), round(d['roofline_kkt']['frac'],3), d['clocks']['sm_mhz'])
"
