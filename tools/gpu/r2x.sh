#!/bin/bash
# A/B of the approximate-quotient Shoup butterfly: libhegpu.so (HEGPU_SHOUP_APPROX=1) against an exact-quotient build
P=homomorphic-encryption-algorithms-diploma-thesis_b200
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2x_gputests.log
for bits in 60,60,60,60 60,40,40,60 50,50,50,50 40,40,40,40; do
  python tools/ntt_bench.py --n 16384 --count 4096 --iters 10 --bits $bits --check >> gpurun_out/r2x_ntt_approx.jsonl 2>&1
done
timeout 300 python bench.py --no-cfg5 --no-imma > gpurun_out/r2x_bench_approx.json 2> gpurun_out/r2x_bench_approx.err
cp $P/libhegpu.so /tmp/libhegpu_approx.so; cp $P/libhegpu_exact.so.alt $P/libhegpu.so
for bits in 60,60,60,60 60,40,40,60 50,50,50,50; do
  python tools/ntt_bench.py --n 16384 --count 4096 --iters 10 --bits $bits >> gpurun_out/r2x_ntt_exact.jsonl 2>&1
done
timeout 300 python bench.py --no-cfg5 --no-imma > gpurun_out/r2x_bench_exact.json 2> gpurun_out/r2x_bench_exact.err
cp /tmp/libhegpu_approx.so $P/libhegpu.so
