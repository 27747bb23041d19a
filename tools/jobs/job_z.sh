python tools/main_py_speed.py 16 16 800 100 1000 2>gpurun_out/r02_mps.err | tee gpurun_out/r02_main_py_speed.jsonl
python tools/main_py_speed.py 16 16 200 100 1000 2>>gpurun_out/r02_mps.err | tee -a gpurun_out/r02_main_py_speed.jsonl
tail -5 gpurun_out/r02_mps.err
