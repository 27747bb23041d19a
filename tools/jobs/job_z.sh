python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 4 --warmup 3 2>gpurun_out/r02_n8.err | tee gpurun_out/r02_bench_n8_final.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=8', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['config']['instances_per_gpu'], d['phase_ms_per_iteration'])"
tail -2 gpurun_out/r02_n8.err
