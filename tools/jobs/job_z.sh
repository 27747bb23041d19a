for b in 2 32; do python bench.py --workload train --batch $b --steps 2 --warmup 3 --graph 2>gpurun_out/r02_train.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('train batch $b', round(d['value'],2), 'inst/s', round(d['ms_per_step'],1), 'ms', d.get('phase_ms_per_iteration'))" | tee -a gpurun_out/r02_train_final2.txt; done
tail -2 gpurun_out/r02_train.err
