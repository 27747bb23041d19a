python -m pytest tests/test_gpu_production_shapes.py -x -q -m gpu -s -k "hidden200" 2>&1 | grep "hidden \|passed\|failed" | tee gpurun_out/r02_k100_hidden_sizes.txt
