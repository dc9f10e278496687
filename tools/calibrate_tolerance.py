"""Calibration of tests/fixtures.py::ckks_tol on the CPU oracle: decrypted error of the BSGS matvec in its exact, hoisted+lazy and
double-hoisted modes at the test and bench shapes, next to the tolerance formulas.  Test tooling (runs the oracle)."""
import sys, time
import os; ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from fixtures import setup, ckks_tol
from oracle import oracle as orc
def run(n, dim, n1, n2, B=2, modes=("dh","exact","fast")):
    S = setup(n, (60,40,40,60)); L=3; scale=2.0**40
    rng = np.random.default_rng(5)
    M = rng.uniform(-1,1,(dim,dim)); V = rng.uniform(-1,1,(B,dim)); slots=n//2
    cts = np.stack([S.encrypt(np.tile(V[i], slots//dim), scale, L, seed=i) for i in range(B)])
    ptsx = np.empty((dim, L+1, n), dtype=np.uint64)
    r = np.arange(dim)
    for g in range(n2):
        for k in range(n1):
            d = g*n1+k
            ptsx[d] = S.enc.encode_ext(np.roll(np.tile(M[r,(r+d)%dim], slots//dim), g*n1), scale, L)
    pts = np.ascontiguousarray(ptsx[:,:L])
    bsteps, gsteps = list(range(1,n1)), [g*n1 for g in range(1,n2)]
    gk = S.gk(bsteps+gsteps)
    bk = [None]+[gk[orc.galois_elt_from_step(n,s)] for s in bsteps]
    gkeys = [None]+[gk[orc.galois_elt_from_step(n,s)] for s in gsteps]
    for mode in modes:
        t=time.time()
        if mode=="dh": out = S.o.matvec_bsgs(cts,n1,n2,ptsx,bk,gkeys,threads=8,dh=True)
        elif mode=="exact": out = S.o.matvec_bsgs(cts,n1,n2,pts,bk,gkeys,threads=8)
        else: out = S.o.matvec_bsgs(cts,n1,n2,pts,bk,gkeys,threads=8,fast=True)
        sc = scale*scale/S.moduli[L-1]
        errs=[np.max(np.abs(S.decrypt(out[i], sc).real[:dim]-M@V[i])) for i in range(B)]
        allslots=[np.max(np.abs(S.decrypt(out[i], sc).real-np.tile(M@V[i], slots//dim))) for i in range(B)]
        print(n,dim,n1,n2,mode, "err", max(errs), "allslots", max(allslots), "old tol", ckks_tol(dim,n,scale), "baseline tol", dim*4096/scale, "t", round(time.time()-t,1), flush=True)
run(8192,16,4,4)
run(16384,32,8,4)
run(16384,128,32,4, modes=("dh",))
run(16384,128,16,8, modes=("exact","fast"))
run(32768,512,32,16, modes=("dh",))
