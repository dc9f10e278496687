#!/bin/bash
# b_k kept in HBM across the launches of dh_inner (more than 4 giant steps): parity, then cfg 5 with and without
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_full_size.py -m gpu -x -q -k "matvec or cfg5 or cfg2" 2>&1 | tail -5 > gpurun_out/r2P_tests.log
for s in 1 0; do
  HEGPU_DH_BK=$s timeout 600 python bench.py --no-cpu-baseline --no-imma --no-micro --cfg5-steps 3 > gpurun_out/r2P_bench_bk$s.json 2> gpurun_out/r2P_bench_bk$s.err
done
