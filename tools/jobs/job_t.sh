export IADMM_B200_LIB=$PWD/i-admm-lstm_b200/iadmm_b200/libiadmm_b200_dev.so
python tools/gemm_mask_study.py 2 100 1.0 2>gpurun_out/r02_t.err | tee gpurun_out/r02_gemm_mask_study.jsonl
python tools/gemm_mask_study.py 8 100 1.0 2>>gpurun_out/r02_t.err | tee -a gpurun_out/r02_gemm_mask_study.jsonl
python tools/gemm_mask_study.py 2 100 3.0 2>>gpurun_out/r02_t.err | tee -a gpurun_out/r02_gemm_mask_study.jsonl
tail -5 gpurun_out/r02_t.err
