#!/bin/bash
# refresh the final-state records at HEAD
python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2Z_gputests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2Z_smoke.log 2>&1
( time python bench.py > gpurun_out/r2Z_bench.json 2> gpurun_out/r2Z_bench.err ) 2> gpurun_out/r2Z_bench.time
( time python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2Z_bench_ref.json 2> gpurun_out/r2Z_bench_ref.err ) 2> gpurun_out/r2Z_bench_ref.time
python bench.py --mode exact --no-cfg5 --no-micro --no-imma --steps 10 > gpurun_out/r2Z_bench_exact.json 2> gpurun_out/r2Z_bench_exact.err
python tools/cfg34_bench.py > gpurun_out/r2Z_cfg34.jsonl 2> gpurun_out/r2Z_cfg34.err
