export IADMM_B200_LIB=$PWD/i-admm-lstm_b200/iadmm_b200/libiadmm_b200_dev.so
for hh in 200 208 800; do for ex in 0 8; do IADMM_TC_EXP=$ex python bench.py --hidden $hh --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-gpu-reference 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('hidden $hh exp $ex', round(d['value'],1), {k:round(v,3) for k,v in d['phase_ms_per_iteration'].items()}, d['clocks']['sm_mhz'])"; done; done
