#!/bin/bash
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2K_gputests.log
for i in 1 2; do timeout 300 python bench.py --no-cpu-baseline --no-imma --no-cfg5 --no-micro >> gpurun_out/r2K_bench.json 2>> gpurun_out/r2K_bench.err; done
