#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_full_size.py -m gpu -x -q -k "matvec or cfg5 or cfg2" 2>&1 | tail -3 > gpurun_out/r2V_tests.log
for i in 1 2 3; do timeout 300 python bench.py --no-cpu-baseline --no-imma --no-cfg5 --no-micro >> gpurun_out/r2V_bench.json 2>> gpurun_out/r2V_bench.err; done
