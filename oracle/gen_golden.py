"""Generate the golden fixtures under ``tests/golden/`` by running the UNTOUCHED reference (T0).

ORACLE / TEST INFRASTRUCTURE ONLY.  Run in the build container (needs ``/root/reference``):

    python oracle/gen_golden.py            # ~4 min, single thread

The fixtures are committed; nothing at test/bench/smoke time reads ``/root/reference``.
Configs: the reference's own test fixture (``tests/test_optimize.py:13-26``: 60x60, load at node 914)
plus a second, smaller case with two loads and non-unit square elements so that numbering, load
placement and scaling are exercised away from nx == ny == L == H.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import run_reference as rr  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")


def pack_ragged(lst):
    off = np.cumsum([0] + [len(a) for a in lst]).astype(np.int64)
    flat = np.concatenate(lst).astype(np.int32) if lst else np.zeros(0, np.int32)
    return flat, off


CASES = {
    # name: (length, height, nx, ny, dirs, positions)
    "c1": (60, 60, 60, 60, [[0, -1]], [[15, 1]]),                      # reference test fixture
    "c2": (60, 30, 40, 20, [[0, -2], [1, 0]], [[6, 1], [21, 1]]),      # dx = dy = 1.5, two loads
}


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, (L, H, nx, ny, dirs, pos) in CASES.items():
        meta = dict(length=L, height=H, nx=nx, ny=ny, dirs=np.array(dirs), positions=np.array(pos))
        t = time.time()
        niter = 50 if name == "c1" else 12
        r = rr.run_simp(L, H, nx, ny, dirs, pos, niter, 3)
        keep = sorted(set([0, 1, 9, len(r["dc"]) - 1]) & set(range(len(r["dc"]))))
        np.savez_compressed(os.path.join(OUT, "simp_%s.npz" % name), niter=niter, penal=3,
                            rho_final=r["rho_final"], compliance=r["compliance"], g=r["g"], rho=r["rho"],
                            dc_iters=np.array(keep), dc=r["dc"][keep], u_absmax=r["u_absmax"], **meta)
        print(name, "SIMP", len(r["compliance"]), "iters", round(time.time() - t, 1), "s", flush=True)

        t = time.time()
        ER, vf = (0.05, 0.5) if name == "c1" else (0.04, 0.6)
        r = rr.run_beso(L, H, nx, ny, dirs, pos, 50, 1e-3, ER, vf)
        flat, off = pack_ragged(r["kept_ids"])
        np.savez_compressed(os.path.join(OUT, "beso_%s.npz" % name), niter=50, t=1e-3, ER=ER, volfrac=vf,
                            ELS=r["ELS"], nodes=r["nodes"], kept_flat=flat, kept_off=off, cons=r["cons"],
                            sensi_filtered=r["sensi_filtered"][:3], C=r["C"], n_protected=r["n_protected"], **meta)
        print(name, "BESO", len(r["kept_ids"]), "removals ->", r["ELS"].shape[0], "els",
              round(time.time() - t, 1), "s", flush=True)

        for which, fn in (("eso_stiff", rr.run_eso_stiff), ("eso_stress", rr.run_eso_stress)):
            t = time.time()
            r = fn(L, H, nx, ny, dirs, pos, 50, 0.005, 0.05, 0.5)
            flat, off = pack_ragged(r["els_ids"])
            np.savez_compressed(os.path.join(OUT, "%s_%s.npz" % (which, name)), niter=50, RR=0.005, ER=0.05,
                                volfrac=0.5, ELS=r["ELS"], nodes=r["nodes"], ids_flat=flat, ids_off=off,
                                cons=r["cons"], C=r["C"], **meta)
            print(name, which, [len(a) for a in r["els_ids"]], "->", r["ELS"].shape[0],
                  round(time.time() - t, 1), "s", flush=True)


if __name__ == "__main__":
    main()
