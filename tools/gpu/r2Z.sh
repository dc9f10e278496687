#!/bin/bash
# round-2 final evidence: suite, default bench, reference arm, exact mode, launch list, one --set full capture per hot
# kernel (raw pages -> profiles/ncu_traffic.json), transform sweep, cfg 3 / cfg 4
python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2Z_gputests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2Z_smoke.log 2>&1
( time python bench.py > gpurun_out/r2Z_bench.json 2> gpurun_out/r2Z_bench.err ) 2> gpurun_out/r2Z_bench.time
( time python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2Z_bench_ref.json 2> gpurun_out/r2Z_bench_ref.err ) 2> gpurun_out/r2Z_bench_ref.time
python bench.py --mode exact --no-cfg5 --no-micro --no-imma --steps 10 > gpurun_out/r2Z_bench_exact.json 2> gpurun_out/r2Z_bench_exact.err
python tools/ntt_bench.py --sweep > gpurun_out/r2Z_ntt_sweep.json 2>&1
for bits in 60,60,60,60 40,40,40,40; do python tools/ntt_bench.py --n 16384 --count 4096 --iters 10 --bits $bits >> gpurun_out/r2Z_ntt_pure.jsonl 2>&1; done
python tools/cfg34_bench.py > gpurun_out/r2Z_cfg34.jsonl 2> gpurun_out/r2Z_cfg34.err
export HEGPU_STREAMS=1
Q="--no-cpu-baseline --no-cfg5 --no-micro --no-imma"
B="python bench.py --steps 2 --warmup 3 $Q"
$B > gpurun_out/r2Z_bench_streams1.json 2> gpurun_out/r2Z_bench_streams1.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -c 700 --csv --log-file gpurun_out/r2Z_launches.csv $B > gpurun_out/r2Z_launches.log 2>&1
tools/ncu_capture.sh r2Z_dhinner 'dh_inner_kernel' 3 -- $B
tools/ncu_capture.sh r2Z_lift 'KsLiftJob' 7 -- $B
tools/ncu_capture.sh r2Z_ksintt 'KsInttJob' 7 -- $B
tools/ncu_capture.sh r2Z_moddown 'KsModDownJob' 3 -- $B
tools/ncu_capture.sh r2Z_halfintt 'HalfInttJob' 3 -- $B
tools/ncu_capture.sh r2Z_finalntt 'FinalNttJob' 3 -- $B
tools/ncu_capture.sh r2Z_ksinnersum 'ks_inner_sum_kernel' 3 -- $B
rm -f gpurun_out/r2Z_halfintt_sass.csv.gz gpurun_out/r2Z_finalntt_sass.csv.gz gpurun_out/r2Z_ksinnersum_sass.csv.gz gpurun_out/r2Z_ksintt_sass.csv.gz
