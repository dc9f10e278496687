#!/bin/bash
P=homomorphic-encryption-algorithms-diploma-thesis_b200
for lib in new old; do
  if [ $lib = old ]; then cp $P/libhegpu.so /tmp/libhegpu_new.so; cp $P/libhegpu_exact.so.alt $P/libhegpu.so; fi
  for n in 4096 8192; do for bits in 60,60,60,60 40,40,40,40 50,50,50,50; do
    python tools/ntt_bench.py --n $n --count 4096 --iters 10 --bits $bits >> gpurun_out/r2z_ntt_$lib.jsonl 2>&1
  done; done
done
cp /tmp/libhegpu_new.so $P/libhegpu.so
