python tools/blocks_probe.py blocks > gpurun_out/bp1.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:kkt_pass -s 8 -c 8 --csv --log-file gpurun_out/r02_blocks_launches.csv python tools/blocks_probe.py blocks > gpurun_out/bp2.log 2>&1
python tools/blocks_probe.py dense > gpurun_out/bp3.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:kkt_pass -s 8 -c 8 --csv --log-file gpurun_out/r02_dense_launches.csv python tools/blocks_probe.py dense > gpurun_out/bp4.log 2>&1
tail -1 gpurun_out/bp1.log
python - <<'PY'
import csv
for f in ("gpurun_out/r02_blocks_launches.csv","gpurun_out/r02_dense_launches.csv"):
    rows=[r for r in csv.reader(open(f)) if len(r)>10]
    hdr=rows[0]; ki=hdr.index("Kernel Name"); mi=hdr.index("Metric Name"); vi=hdr.index("Metric Value")
    for r in rows[1:]:
        print(f.split('/')[-1][:10], r[ki][:40], r[mi], r[vi])
PY
