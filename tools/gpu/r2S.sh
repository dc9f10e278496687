#!/bin/bash
# every runtime switch against the parity tests it can affect
K="ntt_bit_exact or rescale_relin_galois or double_hoisted_levels or matvec"
for v in HEGPU_PARK32K=1 HEGPU_PARK32K=0 HEGPU_PARK=0 HEGPU_LOGE=4 HEGPU_NO_FP64=1 HEGPU_STREAMS=1 HEGPU_DH_BK=0 HEGPU_LIMB_MAJOR=1 HEGPU_FUSE_FINAL=0 HEGPU_DH_FUSED=0 HEGPU_DH_F64=0 HEGPU_DH_STCS=0; do
  echo "== $v" >> gpurun_out/r2S_switches.log
  env $v timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$K" 2>&1 | tail -2 >> gpurun_out/r2S_switches.log
done
echo "== HEGPU_DH_BK=0 cfg5" >> gpurun_out/r2S_switches.log
HEGPU_DH_BK=0 timeout 600 python -m pytest tests/test_full_size.py -m gpu -x -q -k "cfg5" 2>&1 | tail -2 >> gpurun_out/r2S_switches.log
echo "== HEGPU_PARK32K=1 cfg5" >> gpurun_out/r2S_switches.log
HEGPU_PARK32K=1 timeout 600 python -m pytest tests/test_full_size.py -m gpu -x -q -k "cfg5" 2>&1 | tail -2 >> gpurun_out/r2S_switches.log
