"""Golden fixtures at the sizes BASELINE.json quotes (configs[1] and configs[2]), generated with the CPU
restatement T1 (``oracle/t1.py``: numpy assembly + SuperLU ``spsolve``, proven equal to the untouched
reference T0 by ``tests/test_oracle_t1.py``; T0 itself cannot run these sizes -- dense N_e x N_e filter,
O(N_n N_e) Python loops).

ORACLE / TEST INFRASTRUCTURE ONLY.  Run in the build container:

    python oracle/gen_golden_sized.py            # ~10 min, one thread, ~6 GB for the 1024x512 factorisation

  * ``simp_mbb_480x160.npz``  config 2: MBB half beam 480x160, density filter r = 2.5 elements, 6 design
    iterations: compliance / g per iteration, rho after iterations 1 and 6.
  * ``beso_1024x512.npz``     config 3: BESO cantilever 1024x512, the first two removals (three direct solves):
    kept sets as packed bits, thresholds, compliance, protected counts, pinned-DOF flags.
"""
import os
import sys
import time
import unittest.mock as mock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import t1  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")


def mbb_model(nx, ny):
    """MBB half beam: symmetry rollers on the left edge, one vertical support bottom-right, load top-left."""
    nodes, mats, els, loads, _ = t1.beam(L=nx, H=ny, nx=nx, ny=ny, dirs=np.array([[0, -1]]), positions=np.array([[ny, 1]]))
    cons = np.zeros((nodes.shape[0], 2))
    col = np.arange(nodes.shape[0]) % (nx + 1)
    cons[col == 0, 0] = -1
    cons[nx, 1] = -1
    load_node = (nx + 1)*ny
    return nodes, els, cons, load_node


def gen_simp_mbb(nx=480, ny=160, niter=6, r_min=2.5):
    nodes, els, cons, load_node = mbb_model(nx, ny)
    nodes_o = nodes.copy()
    nodes_o[:, 3:5] = cons
    loads_o = np.array([[load_node, 0, -1]])
    ref = {"keep_rho": True}
    t = time.time()
    with mock.patch.object(t1, "beam", lambda **k: (nodes_o, np.tile([206.8e9, 0.28, 1.0], (nx*ny, 1)), els, loads_o, None)):
        rho = t1.simp(nx, ny, nx, ny, None, None, niter, 3, r_min=r_min, history=ref)
    np.savez_compressed(os.path.join(OUT, "simp_mbb_%dx%d.npz" % (nx, ny)), nx=nx, ny=ny, niter=niter, r_min=r_min,
                        compliance=np.array(ref["compliance"]), g=np.array(ref["g"]), rho_first=ref["rho"][0],
                        rho_final=rho, load_node=load_node)
    print("simp mbb", nx, ny, niter, "iterations", round(time.time() - t, 1), "s", ref["compliance"], flush=True)


def gen_beso(nx=1024, ny=512, niter=2):
    dirs, pos = np.array([[0, -1]]), np.array([[ny//4 + 1, 1]])        # load off the mirror axis (no exact ties)
    ref = {}
    t = time.time()
    ELS, nodes = t1.beso(nx, ny, nx, ny, dirs, pos, niter, 1e-3, 0.05, 0.5, history=ref)
    kept = np.zeros((niter, nx*ny), dtype=bool)
    for k, ids in enumerate(ref["kept_ids"]):
        kept[k, ids] = True
    np.savez_compressed(os.path.join(OUT, "beso_%dx%d.npz" % (nx, ny)), nx=nx, ny=ny, niter=niter, dirs=dirs,
                        positions=pos, t=1e-3, ER=0.05, volfrac=0.5, kept_bits=np.packbits(kept, axis=1),
                        alpha=np.array(ref["alpha"]), C=np.array(ref["C"]), n_protected=np.array(ref["n_protected"]),
                        cons_bits=np.packbits(np.array(ref["cons"]) != 0, axis=None), n_els_final=ELS.shape[0])
    print("beso", nx, ny, [int(k.sum()) for k in kept], round(time.time() - t, 1), "s", flush=True)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    what = sys.argv[1:] or ["simp", "beso"]
    if "simp" in what:
        gen_simp_mbb()
    if "beso" in what:
        gen_beso()
