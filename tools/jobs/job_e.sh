python -m pytest tests -m gpu -q 2>&1 | tail -6
python __graft_entry__.py smoke 2>&1 | tail -3
rm -f gpurun_out/r02_e_bench.jsonl
for args in "--workload solve --steps 10 --warmup 3" "--workload solve --sparse auto --steps 5 --no-cpu-baseline --no-gpu-reference" "--workload hidden200 --steps 5 --no-cpu-baseline --no-gpu-reference" "--workload sparse --family SVM --steps 3 --no-cpu-baseline" "--workload sparse --family SVM --sparse off --steps 3 --no-cpu-baseline" "--workload config5 --steps 3 --no-cpu-baseline --no-gpu-reference"; do
  python bench.py $args >> gpurun_out/r02_e_bench.jsonl 2>> gpurun_out/r02_e.err
done
tail -3 gpurun_out/r02_e.err
python -c "
import json
for l in open('gpurun_out/r02_e_bench.jsonl'):
    d=json.loads(l); print(d['config']['workload'][:34], '|', d['config']['matrix_form'][:6], round(d['value'],1), round(d['e2e']['value'],1), {k:round(v,3) for k,v in d['phase_ms_per_iteration'].items()}, round(d['roofline']['frac'],3), round(d['roofline_kkt']['frac'],3), round(d['hbm_roofline_frac_whole_path'],3), d.get('gpu_reference',{}).get('value'), d['clocks']['sm_mhz'])
"
