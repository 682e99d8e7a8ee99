import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.environ.get("PKG_ROOT", ROOT), "solidspy-opt_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import solidspy_opt.optimize as opt
import t1
kind = sys.argv[1] if len(sys.argv) > 1 else "eso_stiff"
nx, ny, niter = 96, 40, 12
fn, fn_ref = (opt.ESO_stiff, t1.eso_stiff) if kind == "eso_stiff" else (opt.ESO_stress, t1.eso_stress)
dirs, pos = np.array([[0, -1]]), np.array([[ny//4 + 1, 1]])
hist, ref = {}, {}
ELS, nodes = fn(nx, ny, nx, ny, dirs, pos, niter, 0.005, 0.05, 0.5, history=hist, verbose=False, precond=os.environ.get("PRECOND", "gmg"), maxit=int(os.environ.get("MAXIT", "20000")))
with np.errstate(all="ignore"):
    ELS_ref, nodes_ref = fn_ref(nx, ny, nx, ny, dirs, pos, niter, 0.005, 0.05, 0.5, history=ref)
print("keys", sorted(hist.keys()), sorted(ref.keys()))
print("mine present:", [a.size for a in hist["els_ids"][:10]])
print("ref  present:", [a.size for a in ref["els_ids"][:10]])
print("mine C:", ["%.6e" % c for c in hist.get("C", [])[:10]])
print("ref  C:", ["%.6e" % c for c in ref.get("C", [])[:10]])
print("cg", hist.get("cg_iters", [])[:12], ["%.1e" % r for r in hist.get("relres", [])[:12]])
for k, (a, b) in enumerate(zip(hist["els_ids"], ref["els_ids"])):
    if not np.array_equal(a, b):
        x = np.setxor1d(a, b)
        print("first difference at removal", k, "xor size", x.size, "ids", x[:20], "rows", (x[:20]//nx), "cols", (x[:20] % nx))
        break
