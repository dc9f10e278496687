#!/bin/bash
python -m pytest tests/test_seal_pin.py -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r2d_sealpin.log
python - > gpurun_out/r2d_pipe.log 2>&1 <<'PY'
import sys
sys.path.insert(0, "tests")
import hegpu_loader
from fixtures import setup
hg = hegpu_loader.load()
S = setup(8192, (60, 40, 40, 60))
ctx = hg.Context(8192, S.moduli, device=0)
for k in range(6):
    print(k, ctx.pipe_peak(k))
PY
