"""Two BESO design iterations and two ESO_stress iterations on 1024x512 (BASELINE configs[2]; ncu target with TOPO_NO_GRAPH=1)."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "solidspy-opt_b200"))
import numpy as np
import solidspy_opt.optimize as opt
nx, ny = 1024, 512
d, p = np.array([[0, -1]]), np.array([[ny//4 + 1, 1]])
h = {}
opt.BESO(nx, ny, nx, ny, d, p, 2, 1e-3, 0.05, 0.5, verbose=False, history=h)
print("BESO cg", h["cg_iters"], flush=True)
h = {}
opt.ESO_stress(nx, ny, nx, ny, d, p, 2, 0.005, 0.05, 0.5, verbose=False, history=h)
print("ESO_stress cg", h["cg_iters"], flush=True)
