#!/bin/bash
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r2Y_bench_4gpu.json 2> gpurun_out/r2Y_bench_4gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus 4 --steps 3 --warmup 1 > gpurun_out/r2Y_bench_ref_4gpu.json 2> gpurun_out/r2Y_bench_ref_4gpu.err
