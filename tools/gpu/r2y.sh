#!/bin/bash
# exact Shoup product as PTX chains + approximate product in the fold loaders: tests, transform micro, bench
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2y_gputests.log
for bits in 60,60,60,60 60,40,40,60 50,50,50,50 40,40,40,40; do
  python tools/ntt_bench.py --n 16384 --count 4096 --iters 10 --bits $bits --check >> gpurun_out/r2y_ntt.jsonl 2>&1
done
for n in 8192 32768; do
  python tools/ntt_bench.py --n $n --count 4096 --iters 10 --bits 60,40,40,60 --check >> gpurun_out/r2y_ntt.jsonl 2>&1
done
timeout 300 python bench.py --no-cfg5 --no-imma > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err
