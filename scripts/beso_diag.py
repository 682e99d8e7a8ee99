import os, sys, time
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "solidspy-opt_b200"))
import numpy as np, torch
from solidspy_opt import _capi
from solidspy_opt.optimize import BESO
from solidspy_opt.Utils.solver import kloc_simp
nx, ny = 1024, 512
d, p = np.array([[0, -1]]), np.array([[ny//4, 1]])
def tb(tag):
    t = {}
    for n in (2, 12):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        BESO(nx, ny, nx, ny, d, p, n, 1e-3, 0.05, 0.5, verbose=False)
        t[n] = time.perf_counter() - t0
    print(tag, "%.1f ms/iter" % ((t[12]-t[2])/10*1e3), flush=True)
BESO(nx, ny, nx, ny, d, p, 2, 1e-3, 0.05, 0.5, verbose=False)
tb("fresh")
ke = kloc_simp(206.8e9, 0.28)
main = _capi.Topo(2048, 1024, ke)
tb("with a 2048x1024 handle alive")
cons = np.zeros((main.nn, 2), np.int8); cons[::2049] = -1
main.set_cons(cons); main.set_loads(np.array([[2049*256 - 1, 0.0, -1.0]])); main.fill(_capi.F_RHO, 0.5)
pp = _capi.SimpParams(penal=3.0, emin=1e-9, emax=1.0, rmin=4.0, dx=1.0, dy=1.0, move=0.2, rtol=1e-9, precond=1, maxit=1000, warm=1)
g = 0.0
for _ in range(3): g, r = main.simp_step(pp, g)
tb("after SIMP steps on it")
main.profile_matvec(True)
for _ in range(2): g, r = main.simp_step(pp, g)
main.profile_matvec(False)
tb("after a profiled pass")
big = _capi.Topo(8192, 4096, ke); big.close()
tb("after creating/closing an 8192x4096 handle")
