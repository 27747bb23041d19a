# round 2, last build on 2 GPUs: the multi-GPU tests, the solve line and the training line (NCCL all-reduce)
python -m pytest tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -3 | tee gpurun_out/r02_multi_final2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 3 --warmup 3 2> gpurun_out/r02_n2.err | tee gpurun_out/r02_bench_n2_final2.json | cut -c1-300
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --workload train --batch 2 --graph --steps 10 --warmup 3 2>> gpurun_out/r02_n2.err | tee gpurun_out/r02_train_n2_final2.json | cut -c1-300
