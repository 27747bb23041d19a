NEW=$PWD/i-admm-lstm_b200/iadmm_b200/libiadmm_b200_dev.so; OLD=$PWD/i-admm-lstm_b200/iadmm_b200/libiadmm_b200_prev.so
rm -f gpurun_out/r02_ragged_ab.jsonl
for H in 208 800; do for i in 1 2; do
  IADMM_B200_LIB=$OLD PROBE_H=$H TAG=old python tools/gate_probe.py >> gpurun_out/r02_ragged_ab.jsonl 2>>gpurun_out/r02_k.err
  IADMM_B200_LIB=$NEW PROBE_H=$H TAG=new python tools/gate_probe.py >> gpurun_out/r02_ragged_ab.jsonl 2>>gpurun_out/r02_k.err
done; done
tail -2 gpurun_out/r02_k.err
python -c "
import json
for i,l in enumerate(open('gpurun_out/r02_ragged_ab.jsonl')):
    d=json.loads(l); print('old' if i%2==0 else 'new', d['h'], round(d['gate_ms'],3), round(d['kkt_ms'],3), round(d['solve_ms_per_iter'],3), d['clocks']['sm_mhz'], {k:float('%.1e'%v) for k,v in d['parity_vs_simt_K3'].items()})
"
python -m pytest tests/test_gpu_parity.py tests/test_gpu_production_shapes.py tests/test_gpu_resident.py -x -q 2>&1 | tail -2
