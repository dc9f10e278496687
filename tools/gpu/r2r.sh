#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "integer_mma" 2>&1 | tail -30 > gpurun_out/r2r_imma_tests.log
for m in 0 1; do HEGPU_DH_IMMA=$m timeout 300 python bench.py --no-cfg5 --no-micro --no-cpu-baseline > gpurun_out/r2r_bench_imma$m.json 2> gpurun_out/r2r_bench_imma$m.err; done
