"""Time BESO iterations at BASELINE config 3 (1024x512) through the drop-in entry point (no history collection)."""
import os, sys, time
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "solidspy-opt_b200"))
import numpy as np
from solidspy_opt.optimize import BESO
nx, ny = 1024, 512
d, p = np.array([[0, -1]]), np.array([[ny//4, 1]])
BESO(nx, ny, nx, ny, d, p, 2, 1e-3, 0.05, 0.5, verbose=False)          # warm-up (module load, first launches)
t = {}
for n in (2, 12):
    t0 = time.time()
    ELS, nodes = BESO(nx, ny, nx, ny, d, p, n, 1e-3, 0.05, 0.5, verbose=False)
    t[n] = time.time() - t0
print("BESO %dx%d: set-up + 2 iterations %.3f s; 12 iterations %.3f s => %.1f ms per design iteration; kept %d of %d"
      % (nx, ny, t[2], t[12], (t[12] - t[2])/10*1e3, ELS.shape[0], nx*ny))
