python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r02_gpu_suite.log
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-gpu-reference --no-literal-loop"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-reference 2>gpurun_out/r02_b.err > gpurun_out/r02_bench_default_v4.json
python bench.py --workload hidden200 --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-reference 2>>gpurun_out/r02_b.err > gpurun_out/r02_bench_hidden200_v4.json
python bench.py --workload config5 --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-reference 2>>gpurun_out/r02_b.err > gpurun_out/r02_bench_config5_v4.json
python - <<'PY'
import json
for f in ("default","hidden200","config5"):
    d=json.loads(open(f"gpurun_out/r02_bench_{f}_v4.json").read().strip().splitlines()[-1])
    print(f, round(d["value"],1), "e2e", round(d["e2e"]["value"],1), d["phase_ms_per_iteration"], "kkt frac", round(d["roofline_kkt"]["frac"],3), "whole", round(d["hbm_roofline_frac_whole_path"],3), d["clocks"]["sm_mhz"])
PY
$CMD > /dev/null 2>&1 && ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:kkt_pass -s 10 -c 2 -o gpurun_out/r02_kkt_tma_full $CMD > gpurun_out/r02_kkt_tma_ncu.log 2>&1
tail -2 gpurun_out/r02_kkt_tma_ncu.log
