python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --workload config5 --gpus 8 --steps 2 --warmup 3 2>gpurun_out/r02_n8c5.err | tee gpurun_out/r02_bench_config5_n8.json
tail -3 gpurun_out/r02_n8c5.err
