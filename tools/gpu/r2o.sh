#!/bin/bash
python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2o_gputests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2o_smoke.log 2>&1
( time python bench.py > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err ) 2> gpurun_out/r2o_bench.time
( time python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2o_bench_ref.json 2> gpurun_out/r2o_bench_ref.err ) 2> gpurun_out/r2o_bench_ref.time
