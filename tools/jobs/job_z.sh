python -m pytest tests/test_gpu_dropin_loop.py -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r02_dropin_tests.log
python tools/dropin_loop.py 256 800 100 2>gpurun_out/r02_dl.err | tee gpurun_out/r02_dropin_loop_v3.jsonl
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8 | tee gpurun_out/r02_smoke.log
python bench.py --steps 3 --warmup 3 2>gpurun_out/r02_bench.err | tee gpurun_out/r02_bench_default_v3.json
