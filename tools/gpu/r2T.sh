#!/bin/bash
Q="--no-cpu-baseline --no-imma --no-cfg5 --no-micro"
for s in 2 3 4; do HEGPU_STREAMS=$s timeout 300 python bench.py $Q > gpurun_out/r2T_streams$s.json 2> gpurun_out/r2T_streams$s.err; done
for b in 384 512; do timeout 300 python bench.py $Q --batch $b > gpurun_out/r2T_batch$b.json 2> gpurun_out/r2T_batch$b.err; done
HEGPU_STREAMS=3 timeout 300 python bench.py $Q --batch 384 > gpurun_out/r2T_streams3_batch384.json 2> gpurun_out/r2T_streams3_batch384.err
