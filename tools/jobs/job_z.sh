export IADMM_B200_LIB=$PWD/i-admm-lstm_b200/iadmm_b200/libiadmm_b200_dev.so
for hh in 208 256; do for ex in 0 1 2 3 4 5; do IADMM_TC_EXP=$ex timeout 300 python bench.py --hidden $hh --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-gpu-reference 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('hidden $hh exp $ex', round(d['value'],1), {k:round(v,3) for k,v in d['phase_ms_per_iteration'].items()}, d['clocks']['sm_mhz'])" | tee -a gpurun_out/r02_h208_ablation.txt; done; done
