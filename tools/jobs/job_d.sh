python -m pytest tests/test_gpu_sparse.py tests/test_gpu_parity.py -x -q 2>&1 | tail -3
export DEVLIB=$PWD/i-admm-lstm_b200/iadmm_b200/libiadmm_b200_dev.so
rm -f gpurun_out/r02_ruiz_ab.jsonl
for e in "A=1" "IADMM_RUIZ_CHAIN=0"; do
  env $e IADMM_B200_LIB=$DEVLIB python tools/ruiz_ab.py 2>gpurun_out/ruiz_ab.err | grep '"form"' >> gpurun_out/r02_ruiz_ab.jsonl
  env $e RZ_N=5000 RZ_B=16 IADMM_B200_LIB=$DEVLIB python tools/ruiz_ab.py 2>>gpurun_out/ruiz_ab.err | grep '"form"' >> gpurun_out/r02_ruiz_ab.jsonl
  env $e RZ_N=203 RZ_B=64 IADMM_B200_LIB=$DEVLIB python tools/ruiz_ab.py 2>>gpurun_out/ruiz_ab.err | grep '"form"' >> gpurun_out/r02_ruiz_ab.jsonl
done
tail -2 gpurun_out/ruiz_ab.err
python -c "
import sys,json
for l in open('gpurun_out/r02_ruiz_ab.jsonl'):
    d=json.loads(l); print(d['form'], d['B'], d['n'], round(d['ms'],2), round(d['passes_equivalent_at_6559GBps'],1), d['sha256'][:12])"
rm -f gpurun_out/r02_sparse_bench.jsonl
for args in "--workload sparse --family SVM" "--workload sparse --family SVM --sparse off" "--workload sparse --family Random_QP" "--workload solve --sparse auto --steps 5" "--workload solve --steps 5"; do
  python bench.py $args --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-reference >> gpurun_out/r02_sparse_bench.jsonl 2>> gpurun_out/r02_sparse.err
done
tail -3 gpurun_out/r02_sparse.err
python -c "
import json
for l in open('gpurun_out/r02_sparse_bench.jsonl'):
    d=json.loads(l); print(d['config']['workload'][:40], '|', d['config']['matrix_form'][:12], round(d['value'],1), {k:round(v,3) for k,v in d['phase_ms_per_iteration'].items()}, d['roofline_kkt']['bytes_per_iteration']/1e9, round(d['roofline_kkt']['frac'],3), d['roofline_kkt'].get('sparse'))
"
( export IADMM_B200_LIB=$DEVLIB; export PROBE_H=208; rm -f gpurun_out/r02_shr208_ab.jsonl
for i in 1 2; do python tools/gate_probe.py >> gpurun_out/r02_shr208_ab.jsonl 2>>gpurun_out/r02_shr.err; IADMM_TC_EPI=3 python tools/gate_probe.py >> gpurun_out/r02_shr208_ab.jsonl 2>>gpurun_out/r02_shr.err; done )
python -c "
import json
for l in open('gpurun_out/r02_shr208_ab.jsonl'):
    d=json.loads(l); print(d['env'].get('IADMM_TC_EPI'), round(d['gate_ms'],3), round(d['kkt_ms'],3), round(d['solve_ms_per_iter'],3), d['clocks']['sm_mhz'], {k:float('%.1e'%v) for k,v in d['parity_vs_simt_K3'].items()})
"
