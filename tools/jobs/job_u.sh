python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py > gpurun_out/r02_bench_default_v2.json 2>gpurun_out/r02_u.err; python -c "
import json
d=json.load(open('gpurun_out/r02_bench_default_v2.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['value'], d.get('gpu_reference',{}).get('value'), d['clocks'])"
python bench.py --workload hidden200 --steps 3 --warmup 3 > gpurun_out/r02_bench_hidden200_v2.json 2>>gpurun_out/r02_u.err; python -c "
import json
d=json.load(open('gpurun_out/r02_bench_hidden200_v2.json')); print(d['config']['workload'], d['value'], d['e2e']['value'], d['roofline'], d['hbm_roofline_frac_whole_path'])"
