python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/r02_gpu_suite_final.log
