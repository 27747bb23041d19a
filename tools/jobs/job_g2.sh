# round 2: unit-pair gate-column order of the row-interleaved kernels -- full GPU suite on the new build, bit-identity against
# the previous build, same-box A/B of the two builds
python -m pytest tests -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r02_gpu_suite_final3.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r02_smoke_final3.log
PREV=$PWD/i-admm-lstm_b200/iadmm_b200/libiadmm_b200_prev.so
: > gpurun_out/r02_gate_col_order_ab.jsonl
python tools/build_ab_hash.py 2>&1 | tail -6 > gpurun_out/hash_new.txt
IADMM_B200_LIB=$PREV python tools/build_ab_hash.py 2>&1 | tail -6 > gpurun_out/hash_prev.txt
if cmp -s gpurun_out/hash_new.txt gpurun_out/hash_prev.txt; then echo '{"bit_identical_to_previous_build": true}'; else echo '{"bit_identical_to_previous_build": false}'; diff gpurun_out/hash_new.txt gpurun_out/hash_prev.txt; fi | tee -a gpurun_out/r02_gate_col_order_ab.jsonl
cat gpurun_out/hash_new.txt >> gpurun_out/r02_gate_col_order_ab.jsonl
for hh in 800 200 800 200; do for b in prev new; do
  if [ $b = prev ]; then export IADMM_B200_LIB=$PREV; else unset IADMM_B200_LIB; fi
  python bench.py --hidden $hh --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-gpu-reference --no-literal-loop 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(json.dumps({'hidden': $hh, 'build': '$b', 'solves_per_s': round(d['value'],1), 'phase_ms': {k:round(v,4) for k,v in d['phase_ms_per_iteration'].items()}, 'sm_mhz': d['clocks']['sm_mhz'], 'whole_path': round(d['hbm_roofline_frac_whole_path'],4)}))" | tee -a gpurun_out/r02_gate_col_order_ab.jsonl
done; done
unset IADMM_B200_LIB
