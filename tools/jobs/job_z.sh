python -m pytest tests/test_gpu_dropin_loop.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r02_dropin_tests.log
python tools/dropin_loop.py 256 800 100 2>gpurun_out/r02_dl.err | tee gpurun_out/r02_dropin_loop_v2.jsonl
python tools/dropin_loop.py 256 200 100 2>>gpurun_out/r02_dl.err | tee -a gpurun_out/r02_dropin_loop_v2.jsonl
