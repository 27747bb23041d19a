mkdir -p gpurun_out; rm -f gpurun_out/probe11.*
for h in 208 800; do for ew in 8 16; do
  PROBE_H=$h IADMM_TC_EPI_WARPS=$ew timeout 300 python tools/gate_probe.py >> gpurun_out/probe11.jsonl 2>> gpurun_out/probe11.err
done; done
tail -5 gpurun_out/probe11.err
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
