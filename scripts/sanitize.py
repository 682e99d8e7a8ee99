"""Small end-to-end runs of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python scripts/sanitize.py
SIMP (multigrid-PCG with the tile level kernels and the cluster tail, filter, OC), the wide-mesh product variant
(two node columns per lane, row length 1 mod 64: the BC-flag ring reads past the last node), BESO, ESO_stress, and the
row-slab path with two slabs in this process."""
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "solidspy-opt_b200"))
from solidspy_opt import _capi                                   # noqa: E402
from solidspy_opt.optimize import SIMP, BESO, ESO_stress, ESO_stiff   # noqa: E402
from solidspy_opt.Utils.solver import kloc_simp                  # noqa: E402

d, p = np.array([[0, -1]]), np.array([[6, 1]])
h = {}
SIMP(160, 80, 160, 80, d, np.array([[20, 1]]), 3, 3, history=h, verbose=False)
print("SIMP 160x80", h["compliance"], h["cg_iters"], flush=True)
h = {}
SIMP(96, 48, 96, 48, d, np.array([[12, 1]]), 2, 3, history=h, verbose=False, slabs=2)
print("SIMP 96x48 on two slabs", h["compliance"], h["cg_iters"], flush=True)
nx, ny = 1088, 8
with _capi.Topo(nx, ny, kloc_simp(206.8e9, 0.28)) as t:
    cons = np.zeros(((nx + 1)*(ny + 1), 2), np.int8)
    cons[::nx + 1] = -1
    t.set_cons(cons)
    t.fill(_capi.F_RHO, 0.7)
    t.simp_scale(3.0, 1e-9, 1.0)
    y = t.matvec(np.random.default_rng(0).standard_normal(2*(nx + 1)*(ny + 1)))
    print("matvec 1088x8 (two node columns per lane)", float(np.abs(y).max()), flush=True)
E, _ = BESO(48, 24, 48, 24, d, p, 3, 1e-3, 0.05, 0.5, verbose=False)
print("BESO 48x24 kept", E.shape[0], flush=True)
E, _ = ESO_stress(48, 24, 48, 24, d, p, 3, 0.005, 0.05, 0.5, verbose=False)
print("ESO_stress 48x24 kept", E.shape[0], flush=True)
E, _ = ESO_stiff(48, 24, 48, 24, d, p, 3, 0.005, 0.05, 0.5, verbose=False)
print("ESO_stiff 48x24 kept", E.shape[0], flush=True)
