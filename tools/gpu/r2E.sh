#!/bin/bash
# per-instruction sampling of the lift / mod-down / plain transforms (source page of one --set full capture each)
export HEGPU_STREAMS=1
Q="--no-cpu-baseline --no-cfg5 --no-micro --no-imma"
B="python bench.py --steps 2 --warmup 3 $Q"
tools/ncu_capture.sh r2E_lift 'KsLiftJob' 7 -- $B
tools/ncu_capture.sh r2E_moddown 'KsModDownJob' 3 -- $B
tools/ncu_capture.sh r2E_ksintt 'KsInttJob' 7 -- $B
