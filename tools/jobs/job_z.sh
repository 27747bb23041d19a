python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "bulk_copy" 2>&1 | tail -5
