# round 2: exp-only tanh in the 16-warp gate kernel (hidden_dim <= 384), same-box A/B on the development build
export IADMM_B200_LIB=$PWD/i-admm-lstm_b200/iadmm_b200/libiadmm_b200_dev.so
: > gpurun_out/r02_tanh_exp_ab.jsonl
for e in 0 2; do
  if [ $e = 2 ]; then export IADMM_TC_EPI=2; else unset IADMM_TC_EPI; fi
  python tools/qualify_tanh.py 200 384 2>&1 | tail -4 | tee -a gpurun_out/r02_tanh_exp_ab.jsonl
done
for hh in 200 384; do for e in 0 2 0 2; do
  if [ $e = 2 ]; then export IADMM_TC_EPI=2; else unset IADMM_TC_EPI; fi
  python bench.py --hidden $hh --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-gpu-reference --no-literal-loop 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(json.dumps({'hidden': $hh, 'IADMM_TC_EPI': '$e', 'solves_per_s': round(d['value'],1), 'phase_ms': {k:round(v,4) for k,v in d['phase_ms_per_iteration'].items()}, 'sm_mhz': d['clocks']['sm_mhz'], 'whole_path': round(d['hbm_roofline_frac_whole_path'],4)}))" | tee -a gpurun_out/r02_tanh_exp_ab.jsonl
done; done
