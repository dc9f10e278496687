#!/bin/bash
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ntt or rescale or matvec" 2>&1 | tail -4 > gpurun_out/r2n_tests.log
for bits in 60,60,60,60 40,40,40,40 60,40,40,60; do python tools/ntt_bench.py --n 16384 --count 4096 --iters 20 --bits $bits; done > gpurun_out/r2n_ntt.jsonl 2>&1
python tools/ntt_bench.py --sweep >> gpurun_out/r2n_ntt.jsonl 2>&1
python bench.py --no-cfg5 --no-cpu-baseline > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err
