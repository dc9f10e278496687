#!/bin/bash
# job resolvers resolved once per CTA (+ optional limb-major order): full suite, A/B in the bench, plain transforms
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2D_gputests.log
Q="--no-cpu-baseline --no-cfg5 --no-imma"
for s in 0 1 0 1; do
  HEGPU_LIMB_MAJOR=$s timeout 300 python bench.py $Q >> gpurun_out/r2D_bench_limbmajor$s.json 2>> gpurun_out/r2D_bench_limbmajor$s.err
done
for bits in 60,40,40,60; do
  python tools/ntt_bench.py --n 16384 --count 4096 --iters 10 --bits $bits --check >> gpurun_out/r2D_ntt.jsonl 2>&1
  python tools/ntt_bench.py --n 32768 --count 4096 --iters 10 --bits $bits --check >> gpurun_out/r2D_ntt.jsonl 2>&1
done
