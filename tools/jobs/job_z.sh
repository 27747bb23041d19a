export IADMM_B200_LIB=$PWD/i-admm-lstm_b200/iadmm_b200/libiadmm_b200_dev.so
for b in 2 8; do for w in 592 300 150 64; do IADMM_TRAIN_KKT_CTAS=$w python bench.py --workload train --batch $b --steps 2 --warmup 3 --graph 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('batch $b want $w', round(d['value'],2), round(d['ms_per_step'],1))" | tee -a gpurun_out/r02_train_kkt_chunks.txt; done; done
