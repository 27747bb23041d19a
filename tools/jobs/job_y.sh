export IADMM_B200_LIB=$PWD/i-admm-lstm_b200/iadmm_b200/libiadmm_b200_dev.so
python -m pytest tests/test_gpu_parity.py -x -q -k "row_interleaved or poisoned" 2>&1 | tail -1
IADMM_TC_IL_BK=32 python -m pytest tests/test_gpu_parity.py -x -q -k "row_interleaved or poisoned" 2>&1 | tail -1
for hh in 200 208 256 128 72; do for bk in 64 32; do IADMM_TC_IL_BK=$bk python bench.py --hidden $hh --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-gpu-reference 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('hidden $hh bk $bk', round(d['value'],1), {k:round(v,3) for k,v in d['phase_ms_per_iteration'].items()}, d['clocks']['sm_mhz'])" | tee -a gpurun_out/r02_il_bk_ab.txt; done; done
