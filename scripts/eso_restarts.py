"""How often does the ESO / BESO host loop fall back from multigrid-PCG to Jacobi-PCG (history['jacobi_restarts'])?"""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "solidspy-opt_b200"))
import numpy as np
import solidspy_opt.optimize as opt
for kind, nx, ny, niter in (("ESO_stiff", 96, 40, 30), ("ESO_stress", 96, 40, 12), ("ESO_stress", 150, 50, 12), ("ESO_stress", 60, 60, 50),
                            ("ESO_stress", 256, 128, 14), ("BESO", 200, 60, 8), ("BESO", 256, 128, 12)):
    dirs, pos = np.array([[0, -1]]), np.array([[ny//4 + 1, 1]])
    h = {}
    if kind == "BESO":
        opt.BESO(nx, ny, nx, ny, dirs, pos, niter, 1e-3, 0.05, 0.5, history=h, verbose=False)
    else:
        getattr(opt, kind)(nx, ny, nx, ny, dirs, pos, niter, 0.005, 0.05, 0.5, history=h, verbose=False)
    print(kind, nx, ny, "solves", len(h.get("cg_iters", [])), "cg", h.get("cg_iters"), "restarts", h.get("jacobi_restarts", 0),
          "max relres %.1e" % max(h.get("relres", [0])), flush=True)
