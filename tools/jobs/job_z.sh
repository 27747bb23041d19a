python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/r02_gpu_suite_final.log
python -m pytest tests -q -m gpu -p no:randomly tests/test_gpu_dropin_loop.py tests/test_gpu_main_py.py 2>&1 | tail -2
