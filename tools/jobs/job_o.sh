export IADMM_B200_LIB=$PWD/i-admm-lstm_b200/iadmm_b200/libiadmm_b200_dev.so
for sms in 148 128 112 96; do IADMM_TC_MAX_SMS=$sms OV_H=208 OV_K=50 python tools/overlap_experiment.py 2>&1 | tail -1; done
for sms in 148 128 112; do IADMM_TC_MAX_SMS=$sms OV_N=5000 OV_B=24 OV_K=20 python tools/overlap_experiment.py 2>&1 | tail -1; done
