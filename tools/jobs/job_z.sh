export IADMM_B200_LIB=$PWD/i-admm-lstm_b200/iadmm_b200/libiadmm_b200_dev.so
timeout 300 python tools/kkt_tma_ab.py IADMM_PDL 2>gpurun_out/r02_pdl.err | grep "^{" | tee gpurun_out/r02_kkt_pdl_ab.jsonl
tail -3 gpurun_out/r02_pdl.err
for sw in 0 1 0 1; do IADMM_PDL=$sw python bench.py --workload hidden200 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-gpu-reference --no-literal-loop 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('hidden200 pdl=$sw', round(d['value'],1), d['phase_ms_per_iteration'])" | tee -a gpurun_out/r02_kkt_pdl_ab.jsonl; done
