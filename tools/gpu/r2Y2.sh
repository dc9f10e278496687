#!/bin/bash
# 2-GPU job at the final state: NCCL tests + bench with the diag-sharded sub-record
python -m pytest tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2Y_multigpu_tests.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2Y_bench_2gpu.json 2> gpurun_out/r2Y_bench_2gpu.err
