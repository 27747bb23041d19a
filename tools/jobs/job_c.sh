python -m pytest tests/test_gpu_parity.py -x -q -k "ruiz or config2 or sharding or poisoned or solve_vs_reference" 2>&1 | tail -3
export DEVLIB=$PWD/i-admm-lstm_b200/iadmm_b200/libiadmm_b200_dev.so
rm -f gpurun_out/r02_ruiz_ab.jsonl
for e in "A=1" "IADMM_RUIZ_CHAIN=0"; do
  env $e IADMM_B200_LIB=$DEVLIB python tools/ruiz_ab.py 2>/dev/null | grep form >> gpurun_out/r02_ruiz_ab.jsonl
  env $e RZ_N=5000 RZ_B=16 IADMM_B200_LIB=$DEVLIB python tools/ruiz_ab.py 2>/dev/null | grep form >> gpurun_out/r02_ruiz_ab.jsonl
  env $e RZ_N=203 RZ_B=64 IADMM_B200_LIB=$DEVLIB python tools/ruiz_ab.py 2>/dev/null | grep form >> gpurun_out/r02_ruiz_ab.jsonl
done
python -c "
import sys,json
for l in open('gpurun_out/r02_ruiz_ab.jsonl'):
    d=json.loads(l); print(d['form'], d['B'], d['n'], round(d['ms'],2), round(d['passes_equivalent_at_6559GBps'],1), d['sha256'][:12])"
rm -f gpurun_out/r02_train_v3.jsonl
for args in "--batch 2" "--batch 2 --graph" "--batch 8 --graph"; do python bench.py --workload train $args --steps 5 --warmup 3 >> gpurun_out/r02_train_v3.jsonl 2>> gpurun_out/r02_train.err; done
tail -3 gpurun_out/r02_train.err
python -c "
import json
for l in open('gpurun_out/r02_train_v3.jsonl'):
    d=json.loads(l); print(d['config']['batch_per_gpu'], d['config']['launch'][:20], round(d['value'],1), round(d['ms_per_step'],1), round(d['e2e']['value'],1), d['config']['loss'])
"
export IADMM_B200_LIB=$DEVLIB; export PROBE_H=208; rm -f gpurun_out/r02_shr208_ab.jsonl
for i in 1 2; do python tools/gate_probe.py >> gpurun_out/r02_shr208_ab.jsonl 2>>gpurun_out/r02_shr.err; IADMM_TC_EPI=3 python tools/gate_probe.py >> gpurun_out/r02_shr208_ab.jsonl 2>>gpurun_out/r02_shr.err; done
python -c "
import json
for l in open('gpurun_out/r02_shr208_ab.jsonl'):
    d=json.loads(l); print(d['env'].get('IADMM_TC_EPI'), round(d['gate_ms'],3), round(d['kkt_ms'],3), round(d['solve_ms_per_iter'],3), d['clocks']['sm_mhz'], {k:float('%.1e'%v) for k,v in d['parity_vs_simt_K3'].items()})
"
