#!/bin/bash
# round-2 evidence at the final state: streaming-store A/B, launch list, one --set full capture per hot kernel
Q="--no-cpu-baseline --no-cfg5 --no-micro --no-imma"
for s in 0 1; do
  HEGPU_DH_STCS=$s timeout 300 python bench.py $Q > gpurun_out/r2A_bench_stcs$s.json 2> gpurun_out/r2A_bench_stcs$s.err
done
export HEGPU_STREAMS=1
B="python bench.py --steps 2 --warmup 3 $Q"
$B > gpurun_out/r2A_bench_streams1.json 2> gpurun_out/r2A_bench_streams1.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -c 700 --csv --log-file gpurun_out/r2A_launches.csv $B > gpurun_out/r2A_launches.log 2>&1
tools/ncu_capture.sh r2A_dhinner 'dh_inner_kernel' 3 -- $B
HEGPU_DH_STCS=0 tools/ncu_capture.sh r2A_dhinner_plainstores 'dh_inner_kernel' 3 -- $B
tools/ncu_capture.sh r2A_lift 'KsLiftJob' 7 -- $B
tools/ncu_capture.sh r2A_ksintt 'KsInttJob' 7 -- $B
tools/ncu_capture.sh r2A_moddown 'KsModDownJob' 3 -- $B
tools/ncu_capture.sh r2A_halfintt 'HalfInttJob' 3 -- $B
tools/ncu_capture.sh r2A_finalntt 'FinalNttJob' 3 -- $B
tools/ncu_capture.sh r2A_ksinnersum 'ks_inner_sum_kernel' 3 -- $B
rm -f gpurun_out/r2A_*_sass.csv.gz
