timeout 300 python -m pytest tests/test_gpu_training.py tests/test_gpu_production_shapes.py -x -q -k "window or gradient or recomputed" 2>&1 | tail -3
rm -f gpurun_out/r02_train_v4.jsonl
for args in "--batch 32" "--batch 8" "--batch 2 --graph"; do timeout 300 python bench.py --workload train $args --steps 3 --warmup 3 >> gpurun_out/r02_train_v4.jsonl 2>> gpurun_out/r02_n.err; done
tail -3 gpurun_out/r02_n.err
python -c "
import json
for i,l in enumerate(open('gpurun_out/r02_train_v4.jsonl')):
    d=json.loads(l); print(d['config']['batch_per_gpu'], round(d['value'],1), round(d['ms_per_step'],1), {k:round(v,3) for k,v in d['phase_ms_per_iteration'].items()}, round(d['roofline']['frac'],3), d['config']['loss'], d['clocks']['sm_mhz'])
"
