#!/bin/bash
# GPU job: seal-pin GPU tests, default bench (new record layout), exact-mode bench
python -m pytest tests/test_seal_pin.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2c_sealpin.log
( time python bench.py > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err ) 2> gpurun_out/r2c_bench.time
python bench.py --mode exact --no-cfg5 --no-micro --steps 10 > gpurun_out/r2c_bench_exact.json 2> gpurun_out/r2c_bench_exact.err
