"""Diagnostic sweep on a GPU box: prints numbers instead of asserting (development aid)."""
import os, sys, time, traceback
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "solidspy-opt_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import scipy.sparse as sp
from scipy.sparse.linalg import spsolve
import t1, gmg_proto
from solidspy_opt import _capi
import solidspy_opt.optimize as opt

E0, NU = 206.8e9, 0.28
ke = t1.kloc_closed_form(E0, NU)

def section(name):
    print("\n==== %s ====" % name, flush=True)

def guarded(fn):
    try:
        fn()
    except Exception:
        traceback.print_exc()
        sys.stdout.flush()

def cant(nx, ny):
    cons = np.zeros(((nx+1)*(ny+1), 2), np.int8); cons[np.arange((nx+1)*(ny+1)) % (nx+1) == 0] = -1
    return cons

def p_matvec():
    section("matvec")
    for nx, ny in [(60, 60), (31, 5), (33, 17), (129, 40), (1, 1), (300, 2)]:
        rng = np.random.default_rng(nx)
        cons = cant(nx, ny); scale = 1e-9 + rng.uniform(0, 1, nx*ny)**3
        u = rng.standard_normal(2*(nx+1)*(ny+1))
        with _capi.Topo(nx, ny, ke) as t:
            t.set_cons(cons); t.set(_capi.F_SCALE, scale)
            y = t.matvec(u).reshape(-1)
        K = gmg_proto.full_K(nx, ny, scale, ke); free = (cons.reshape(-1) == 0).astype(float)
        ref = free*(K @ (u*free))
        print(nx, ny, "rel err", np.abs(y-ref).max()/np.abs(ref).max(), flush=True)

def p_solve():
    section("solve")
    for nx, ny in [(60, 60), (37, 23), (240, 120)]:
        rng = np.random.default_rng(nx)
        cons = cant(nx, ny); rho = rng.uniform(0.05, 1, nx*ny); scale = 1e-9 + rho**3
        loads = np.array([[(nx+1)*(ny//4) + nx, 0.0, -1.0]])
        K = gmg_proto.full_K(nx, ny, scale, ke); free = cons.reshape(-1) == 0
        idx = np.flatnonzero(free); f = np.zeros(K.shape[0]); f[2*int(loads[0,0])+1] = -1
        uref = np.zeros_like(f); uref[idx] = spsolve(K[idx][:, idx].tocsc(), f[idx])
        for pc, name in ((_capi.PRECOND_JACOBI, "jacobi"), (_capi.PRECOND_GMG, "gmg")):
            with _capi.Topo(nx, ny, ke) as t:
                t.set_cons(cons); t.set(_capi.F_SCALE, scale); t.set_loads(loads)
                t0 = time.time()
                it, rel, code = t.solve(pc, 1e-10, 50000, warm=False, raise_on_noconv=False)
                dt = time.time() - t0
                u = t.get(_capi.F_U).reshape(-1)
                print(nx, ny, name, "iters", it, "relres %.2e" % rel, "code", code, "err vs direct %.2e" % (np.abs(u-uref).max()/np.abs(uref).max()),
                      "levels", len(t.gmg_levels()), "%.1f ms" % (dt*1e3), "eq_ok", t.equilibrium_ok(), flush=True)
        # numpy prototype iteration count for the same problem
        mg = gmg_proto.GMG(nx, ny, scale, ke, ~free, min_coarse=1)
        _, itp, _ = gmg_proto.pcg(mg.levels[0]["A"], f, mg.vcycle, rtol=1e-10)
        print(nx, ny, "numpy-proto gmg iters", itp, flush=True)

def p_simp():
    section("SIMP c1")
    g = np.load(os.path.join(ROOT, "tests/golden/simp_c1.npz"))
    for pc in ("gmg", "jacobi"):
        hist = {"keep_rho": True}
        t0 = time.time()
        rho = opt.SIMP(60, 60, 60, 60, g["dirs"], g["positions"], 50, 3, precond=pc, history=hist, verbose=False)
        dt = time.time() - t0
        n = min(len(hist["compliance"]), len(g["compliance"]))
        print(pc, "iters", len(hist["compliance"]), "(ref %d)" % len(g["compliance"]), "%.2f s" % dt)
        print("  compliance rel err max", np.max(np.abs(np.array(hist["compliance"][:n]) - g["compliance"][:n])/g["compliance"][:n]))
        print("  rho abs err per iter", ["%.1e" % np.abs(hist["rho"][k]-g["rho"][k]).max() for k in range(0, n, 4)])
        print("  cg iters", hist["cg_iters"][:n:4], "oc steps", hist["oc_steps"][:5], flush=True)

def p_beso():
    section("BESO/ESO c1")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import unpack_ragged
    g = np.load(os.path.join(ROOT, "tests/golden/beso_c1.npz"))
    hist = {}
    t0 = time.time()
    ELS, nodes = opt.BESO(60, 60, 60, 60, g["dirs"], g["positions"], 50, 1e-3, 0.05, 0.5, history=hist, verbose=False)
    kept = unpack_ragged(g["kept_flat"], g["kept_off"])
    print("BESO %.2f s" % (time.time()-t0), "removals", len(hist.get("kept_ids", [])), "(ref %d)" % len(kept))
    for k, (a, b) in enumerate(zip(hist.get("kept_ids", []), kept)):
        if not np.array_equal(a, b):
            print("  first mismatch at removal", k, a.size, b.size, "symdiff", np.setxor1d(a, b)[:10]); break
    else:
        print("  all removal sets identical; ELS equal:", np.array_equal(ELS, g["ELS"]), "nodes equal:", np.array_equal(nodes, g["nodes"]))
    print("  cg iters", hist.get("cg_iters"), flush=True)
    for kind in ("eso_stiff", "eso_stress"):
        g = np.load(os.path.join(ROOT, "tests/golden/%s_c1.npz" % kind)); hist = {}
        fn = opt.ESO_stiff if kind == "eso_stiff" else opt.ESO_stress
        ELS, nodes = fn(60, 60, 60, 60, g["dirs"], g["positions"], 50, 0.005, 0.05, 0.5, history=hist, verbose=False)
        ids = unpack_ragged(g["ids_flat"], g["ids_off"])
        print(kind, [len(a) for a in hist.get("els_ids", [])], "ref", [len(a) for a in ids],
              "equal:", all(np.array_equal(a, b) for a, b in zip(hist.get("els_ids", []), ids)),
              "ELS", np.array_equal(ELS, g["ELS"]), "nodes", np.array_equal(nodes, g["nodes"]), flush=True)

def p_big():
    section("2048x1024")
    nx, ny = 2048, 1024
    with _capi.Topo(nx, ny, ke) as t:
        t.set_cons(cant(nx, ny)); t.fill(_capi.F_RHO, 0.5); t.simp_scale(3.0, 1e-9, 1.0)
        t.set(_capi.F_U, np.random.default_rng(0).standard_normal(2*t.nn))
        ms = t.matvec_device(20)
        print("matvec %.1f us -> %.0f GB/s algorithmic" % (ms*1e3, (32*t.nn + 8*t.ne)/ms/1e6), flush=True)
        t.set(_capi.F_U, np.zeros(2*t.nn))
        t.set_loads(np.array([[(nx+1)*(ny//4) - 1 + 0, 0.0, -1.0]]))
        p = _capi.SimpParams(penal=3.0, emin=1e-9, emax=1.0, rmin=4.0, dx=1.0, dy=1.0, move=0.2, rtol=1e-10,
                             precond=_capi.PRECOND_GMG, maxit=2000, warm=1)
        g = 0.0
        for i in range(6):
            if i == 5:
                t.profile_matvec(True)
            t0 = time.time()
            g, res = t.simp_step(p, g)
            tm = t.timing()
            print("it %d: %.1f ms wall | cg %d relres %.1e | oc %d | compliance %.6e change %.3f | solve %.1f sens %.2f filt %.2f oc %.2f setup %.2f total %.1f ms | launches %d"
                  % (i, (time.time()-t0)*1e3, res.cg_iters, res.relres, res.oc_steps, res.compliance, res.change,
                     tm.solve_ms, tm.sens_ms, tm.filter_ms, tm.oc_ms, tm.setup_ms, tm.total_ms, tm.kernel_launches), flush=True)
        rep = t.profile_report()
        tot = sum(r[2] for r in rep)
        print("in-situ kernel profile of the last iteration (sum %.2f ms):" % tot)
        for name, n, ms in rep:
            print("  %-28s n=%5d total=%8.3f ms avg=%8.2f us %5.1f%%" % (name, n, ms, ms/n*1e3, 100*ms/tot))

if __name__ == "__main__":
    which = sys.argv[1:] or ["matvec", "solve", "simp", "beso", "big"]
    print("devices:", _capi.device_count())
    for w in which:
        guarded({"matvec": p_matvec, "solve": p_solve, "simp": p_simp, "beso": p_beso, "big": p_big}[w])
