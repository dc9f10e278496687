#!/bin/bash
# usage: tools/ncu_capture.sh <tag> <kernel-regex (demangled)> <skip> -- <cmd...>
# One `ncu --set full` capture of one launch; the .ncu-rep stays on the GPU box (/tmp), only the
# raw-metric CSV and the gzipped per-instruction (SASS) CSV come back under gpurun_out/.
tag=$1; regex=$2; skip=$3; shift 4
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$regex" -s "$skip" -c 1 \
    -f -o /tmp/prof_$tag "$@" > gpurun_out/ncu_$tag.log 2>&1
ncu -i /tmp/prof_$tag.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
ncu -i /tmp/prof_$tag.ncu-rep --page source --print-source sass --csv 2>/dev/null | gzip -9 > gpurun_out/${tag}_sass.csv.gz
ncu -i /tmp/prof_$tag.ncu-rep --page details > gpurun_out/${tag}_details.txt 2>/dev/null
ls -la /tmp/prof_$tag.ncu-rep gpurun_out/${tag}_*
