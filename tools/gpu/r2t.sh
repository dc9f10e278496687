#!/bin/bash
for tx in 16 8; do
  HEGPU_IMMA_TX=$tx timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "integer_mma" 2>&1 | tail -3 > gpurun_out/r2t_imma_tests_tx$tx.log
  HEGPU_IMMA_TX=$tx HEGPU_DH_IMMA=1 timeout 300 python bench.py --no-cfg5 --no-micro --no-cpu-baseline > gpurun_out/r2t_bench_tx$tx.json 2> gpurun_out/r2t_bench_tx$tx.err
done
