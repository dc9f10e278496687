#!/bin/bash
for cfg in "128:2" "128:3" "128:4" "256:2" "256:3" "256:4" "384:3"; do
  b=${cfg%%:*}; s=${cfg##*:}
  HEGPU_STREAMS=$s python bench.py --batch $b --no-cfg5 --no-micro --no-cpu-baseline --steps 20 > gpurun_out/r2m_b${b}_s${s}.json 2> gpurun_out/r2m_b${b}_s${s}.err
done
