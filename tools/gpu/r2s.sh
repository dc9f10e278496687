#!/bin/bash
HEGPU_IMMA_TX=8 HEGPU_DH_IMMA=1 HEGPU_STREAMS=1 tools/ncu_capture.sh r2s_imma 'dh_imma_kernel' 2 -- python bench.py --batch 128 --steps 2 --warmup 3 --no-cpu-baseline --no-cfg5 --no-micro
rm -f gpurun_out/r2s_imma_sass.csv.gz
