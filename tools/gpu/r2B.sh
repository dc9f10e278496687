#!/bin/bash
# two-level park for N = 32768: parity, transform micro and cfg5 against the one-level form
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_full_size.py -m gpu -x -q -k "32768 or cfg5 or ntt or least_squares" 2>&1 | tail -5 > gpurun_out/r2B_tests.log
for pk in 2 1; do
  for bits in 60,40,40,60 60,60,60,60 40,40,40,40; do
    HEGPU_PARK32K=$pk python tools/ntt_bench.py --n 32768 --count 4096 --iters 10 --bits $bits --check >> gpurun_out/r2B_ntt_park$pk.jsonl 2>&1
  done
  HEGPU_PARK32K=$pk timeout 600 python bench.py --steps 5 --no-micro --no-imma --no-cpu-baseline > gpurun_out/r2B_bench_park$pk.json 2> gpurun_out/r2B_bench_park$pk.err
done
