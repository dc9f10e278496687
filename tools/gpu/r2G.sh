#!/bin/bash
export HEGPU_STREAMS=1
Q="--no-cpu-baseline --no-cfg5 --no-micro --no-imma"
B="python bench.py --steps 2 --warmup 3 $Q"
tools/ncu_capture.sh r2G_dhinner 'dh_inner_kernel' 3 -- $B
