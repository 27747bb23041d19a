export IADMM_B200_LIB=$PWD/i-admm-lstm_b200/iadmm_b200/libiadmm_b200_dev.so
for mm in "7 7" "6 7" "5 7" "4 7" "7 6" "7 5" "7 4" "6 4" "4 4"; do
  set -- $mm
  echo "== GEMM1 mask $1 GEMM2 mask $2"
  IADMM_GEMM1_MASK=$1 IADMM_GEMM2_MASK=$2 python -m pytest tests/test_gpu_production_shapes.py -q -s -k "config3_shape" 2>&1 | grep -E "config-3 shape|passed|failed|assert" | cut -c1-400
done
for mm in "7 7" "6 4" "4 4"; do
  set -- $mm
  IADMM_GEMM1_MASK=$1 IADMM_GEMM2_MASK=$2 python bench.py --workload train --batch 32 --steps 2 --warmup 3 --graph 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1 $2', round(d['value'],1), round(d['ms_per_step'],1), {k:round(v,3) for k,v in d['phase_ms_per_iteration'].items()})"
done
