#!/bin/bash
for e in 3 4; do for bits in 40,40,40,40 60,60,60,60 60,40,40,60; do
  HEGPU_LOGE=$e python tools/ntt_bench.py --n 16384 --count 4096 --iters 10 --bits $bits >> gpurun_out/r2U_loge$e.jsonl 2>&1
done; done
