#!/bin/bash
# NOTE: needs the HEGPU_IMMA_DBG switches of a throw-away debug build of imma_kernels.cuh (skip MMAs / gathers / epilogue / unpack); kept as the record of how profiles/r2u_imma_components.jsonl was made
# component timing of the integer-MMA kernel (debug switches; results are wrong by construction, only the time matters)
for dbg in 0 1 2 4 8 3 9 11 15; do
  HEGPU_IMMA_DBG=$dbg HEGPU_DH_IMMA=1 HEGPU_STREAMS=1 python - <<PY >> gpurun_out/r2u_imma_components.jsonl 2>&1
import sys, json, numpy as np
sys.path.insert(0, "tests")
import hegpu_loader, torch
from fixtures import setup, rand_residues
hg = hegpu_loader.load()
n, n1, n2, B, L = 16384, 32, 4, 128, 3
S = setup(n, (60, 40, 40, 60))
ctx = hg.Context(n, S.moduli)
rng = np.random.default_rng(1)
steps = list(range(1, n1)) + [g * n1 for g in range(1, n2)]
ctx.load_galois_keys(S.gk(steps))
X = ctx.upload_ct(rand_residues(rng, S.moduli[:L], (B, 2), n), 2.0**40)
D = ctx.upload_pt_ext(rand_residues(rng, S.moduli[:L] + [S.moduli[-1]], (n1 * n2,), n), 2.0**40)
out = ctx.ct(B, 2)
for _ in range(2): ctx.matvec_bsgs(out, X, D, n1, n2, dh=True)
ctx.profile_reset(); ctx.profile(True)
for _ in range(5): ctx.matvec_bsgs(out, X, D, n1, n2, dh=True)
p = ctx.profile_read(); ctx.profile(False)
print(json.dumps({"dbg": $dbg, "dh_inner_ms": p["dh_inner"]["ms"] / 5}))
PY
done
