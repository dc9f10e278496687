#!/bin/bash
python -m pytest tests/test_seal_pin.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2e_sealpin.log
for swz in 0 1; do
  HEGPU_DH_SWZ=$swz python bench.py --no-cfg5 --no-micro --no-cpu-baseline > gpurun_out/r2e_bench_swz$swz.json 2> gpurun_out/r2e_bench_swz$swz.err
done
