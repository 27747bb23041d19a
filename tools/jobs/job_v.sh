python -m pytest tests/test_gpu_training.py -q -s -k "test_window_gradients_match_autograd" 2>&1 | grep -v "^$" | tail -40 | cut -c1-500
