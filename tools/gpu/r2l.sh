#!/bin/bash
# ncu --set full of the plain forward transform at 4096 limbs: pure 60-bit chain, pure 40-bit chain, mixed
for case in "60,60,60,60:i60" "40,40,40,40:f40" "60,40,40,60:mix"; do
  bits=${case%%:*}; tag=${case##*:}
  python tools/ntt_bench.py --n 16384 --count 4096 --iters 5 --bits $bits > gpurun_out/r2l_ntt_$tag.json 2>&1
  tools/ncu_capture.sh r2l_fwd_$tag 'ntt_fwd_park_kernel' 4 -- python tools/ntt_bench.py --n 16384 --count 4096 --iters 3 --bits $bits
  rm -f gpurun_out/r2l_fwd_${tag}_sass.csv.gz
done
