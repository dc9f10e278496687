#!/bin/bash
# batch fix + FP64-native fold loaders + explicit global loads: full suite, bench with cfg5
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2F_gputests.log
timeout 600 python bench.py --no-cpu-baseline --no-imma > gpurun_out/r2F_bench.json 2> gpurun_out/r2F_bench.err
timeout 300 python bench.py --no-cpu-baseline --no-imma --no-cfg5 --no-micro >> gpurun_out/r2F_bench.json 2>> gpurun_out/r2F_bench.err
