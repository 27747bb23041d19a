python -m pytest tests/test_gpu_dropin_loop.py -x -q -m gpu 2>&1 | tail -15
