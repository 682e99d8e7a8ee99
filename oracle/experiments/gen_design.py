"""CPU study behind DESIGN.md section 4 "Late in an optimisation" (ORACLE-SIDE EXPERIMENT, numpy/scipy only; nothing in the
product imports this).  Runs the oracle's SIMP (SuperLU) on a cantilever to convergence and stores the last three designs.
Run from the repo root:  python oracle/experiments/gen_design.py 256 128 100"""
import os
import sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, HERE)
OUT = os.environ.get("PROTO_DIR", "/tmp/proto")
os.makedirs(OUT, exist_ok=True)
import time
import numpy as np
import t1
nx, ny, nit = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
dirs, pos = np.array([[0, -1]]), np.array([[ny//4, 1]])
hist = {}
t0 = time.time()
rho = t1.simp(nx, ny, nx, ny, dirs, pos, nit, 3, history=hist)
n = len(hist["compliance"])
print("time", time.time() - t0, "iters", n, hist["compliance"][-1])
# rho entering design iterations n-1 and n (rho list holds the rho AFTER each iteration)
np.savez(os.path.join(OUT, "design_%dx%d_%d.npz" % (nx, ny, nit)), rho_prev=hist["rho"][-3], rho_cur=hist["rho"][-2], rho_next=hist["rho"][-1])
