#!/bin/bash
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2p_tests.log
for bits in 60,60,60,60 40,40,40,40 60,40,40,60; do python tools/ntt_bench.py --n 16384 --count 4096 --iters 20 --bits $bits; done > gpurun_out/r2p_ntt.jsonl 2>&1
python tools/ntt_bench.py --n 32768 --count 4096 --iters 10 >> gpurun_out/r2p_ntt.jsonl 2>&1
python tools/ntt_bench.py --n 8192 --count 4096 --iters 10 >> gpurun_out/r2p_ntt.jsonl 2>&1
python bench.py --no-cfg5 --no-cpu-baseline > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err
