"""PCG iteration counts of BESO solves at config 3 (diagnostic)."""
import os, sys, time
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "solidspy-opt_b200"))
import numpy as np
from solidspy_opt.optimize import BESO
nx, ny = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1024, 512)
d, p = np.array([[0, -1]]), np.array([[ny//4, 1]])
h = {}
t0 = time.time()
ELS, nodes = BESO(nx, ny, nx, ny, d, p, int(os.environ.get("N_IT", "12")), 1e-3, 0.05, 0.5, verbose=False, history=h)
print("time %.3f s" % (time.time() - t0), "cg_iters", h.get("cg_iters"), "relres", ["%.1e" % r for r in h.get("relres", [])])
