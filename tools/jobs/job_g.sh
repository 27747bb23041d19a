rm -f gpurun_out/r02_train_scaling.jsonl
for N in 1 2 4 8; do
  if [ $N -eq 1 ]; then
    python bench.py --workload train --batch 2 --graph --steps 10 --warmup 3 >> gpurun_out/r02_train_scaling.jsonl 2>> gpurun_out/r02_train_scaling.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --workload train --batch 2 --graph --steps 10 --warmup 3 >> gpurun_out/r02_train_scaling.jsonl 2>> gpurun_out/r02_train_scaling.err
  fi
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29650 bench.py --gpus 8 --workload train --batch 32 --steps 3 --warmup 3 >> gpurun_out/r02_train_scaling.jsonl 2>> gpurun_out/r02_train_scaling.err
python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -2
grep -v "OMP_NUM\|\*\*\*\*\|^$" gpurun_out/r02_train_scaling.err | tail -5
python -c "
import json
for l in open('gpurun_out/r02_train_scaling.jsonl'):
    d=json.loads(l); print(d['n_gpus'], d['config']['batch_per_gpu'], round(d['value'],1), round(d['ms_per_step'],2), d['allreduce'], d['config']['weights_identical_across_ranks'], round(d['e2e']['value'],1))
"
