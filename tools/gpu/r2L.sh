#!/bin/bash
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2L_gputests.log
for i in 1 2; do timeout 300 python bench.py --no-cpu-baseline --no-imma --no-cfg5 --no-micro >> gpurun_out/r2L_bench.json 2>> gpurun_out/r2L_bench.err; done
for n in 16384 32768; do python tools/ntt_bench.py --n $n --count 4096 --iters 10 --bits 60,40,40,60 --check >> gpurun_out/r2L_ntt.jsonl 2>&1; done
