python -m pytest tests/test_gpu_main_py.py -x -q -m gpu -s 2>&1 | grep -v Warning | tail -40 | tee gpurun_out/r02_main_py_test.log
