#!/bin/bash
# GPU job: full -m gpu suite, default bench, pipe peaks
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2b_gputests.log
python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
python - > gpurun_out/r2b_pipe.log 2>&1 <<'PY'
import sys
sys.path.insert(0, "tests")
import hegpu_loader
from fixtures import setup
hg = hegpu_loader.load()
S = setup(8192, (60, 40, 40, 60))
ctx = hg.Context(8192, S.moduli, device=0)
for k in range(3):
    print(k, ctx.pipe_peak(k))
PY
