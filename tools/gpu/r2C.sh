#!/bin/bash
# limb-major job order of the mod-down / final NTT launches: full suite, then A/B in the bench
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2C_gputests.log
Q="--no-cpu-baseline --no-cfg5 --no-micro --no-imma"
for s in 0 1 0 1; do
  HEGPU_LIMB_MAJOR=$s timeout 300 python bench.py $Q >> gpurun_out/r2C_bench_limbmajor$s.json 2>> gpurun_out/r2C_bench_limbmajor$s.err
done
