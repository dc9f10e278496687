#!/bin/bash
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2Q_gputests.log
for i in 1 2; do timeout 300 python bench.py --no-cpu-baseline --no-imma --no-cfg5 --no-micro >> gpurun_out/r2Q_bench.json 2>> gpurun_out/r2Q_bench.err; done
timeout 300 python bench.py --mode exact --no-cpu-baseline --no-cfg5 --no-micro --no-imma --steps 10 > gpurun_out/r2Q_bench_exact.json 2> gpurun_out/r2Q_bench_exact.err
