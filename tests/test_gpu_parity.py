"""Parity of the CUDA path (through the C ABI / the drop-in package) against the oracle and the goldens.

All tests here need a B200 (``-m gpu``).  Tolerances: integer / index / mask results are bit-exact;
floating-point results use the tolerance BASELINE.json's north_star states (1e-6 relative on compliance
and density per iteration), tighter where the stage is a pure kernel (1e-12..1e-13).
"""
import numpy as np
import pytest
import scipy.sparse as sp
from scipy.sparse.linalg import spsolve

import gmg_proto
import t1
from conftest import golden, unpack_ragged

pytestmark = pytest.mark.gpu

E0, NU = 206.8e9, 0.28


@pytest.fixture(scope="module")
def capi(lib_built):
    from solidspy_opt import _capi
    _capi.device_count()
    return _capi


def _cons(nx, ny, kind, rng=None):
    nn = (nx + 1)*(ny + 1)
    cons = np.zeros((nn, 2), dtype=np.int8)
    col = np.arange(nn) % (nx + 1)
    row = np.arange(nn)//(nx + 1)
    if kind == "cantilever":
        cons[col == 0] = -1
    elif kind == "mbb":                      # symmetry rollers on the left edge + one vertical support
        cons[col == 0, 0] = -1
        cons[nx, 1] = -1
    elif kind == "right_partial":            # beam_2 style
        cons[(col == nx) & ((row > 3*ny//4) | (row < ny//4))] = -1
    elif kind == "scattered":
        cons[col == 0] = -1
        pick = rng.choice(nn, size=max(1, nn//15), replace=False)
        cons[pick, rng.integers(0, 2, pick.size)] = -1
    return cons


def _assembled(nx, ny, scale, ke, cons):
    K = gmg_proto.full_K(nx, ny, scale, ke)
    free = (cons.reshape(-1) == 0).astype(float)
    M = sp.diags(free)
    return (M @ K @ M).tocsr(), free


@pytest.mark.parametrize("nx,ny,kind", [(60, 60, "cantilever"), (31, 5, "mbb"), (32, 9, "right_partial"),
                                        (33, 17, "scattered"), (129, 40, "cantilever"), (1, 1, "mbb"),
                                        (300, 2, "cantilever"), (2, 200, "scattered")])
def test_matvec_matches_assembled_matrix(capi, nx, ny, kind):
    rng = np.random.default_rng(nx*1000 + ny)
    ke = t1.kloc_closed_form(E0, NU)
    cons = _cons(nx, ny, kind, rng)
    scale = 1e-9 + rng.uniform(0, 1, nx*ny)**3
    u = rng.standard_normal(2*(nx + 1)*(ny + 1))
    with capi.Topo(nx, ny, ke) as t:
        t.set_cons(cons)
        t.set(capi.F_SCALE, scale)
        y = t.matvec(u).reshape(-1)
    K, free = _assembled(nx, ny, scale, ke, cons)
    ref = K @ (u*free)
    assert np.abs(y - ref).max() <= 1e-13*np.abs(ref).max()
    assert np.all(y[free == 0] == 0.0)


def test_matvec_general_rectangular_ke(capi):
    rng = np.random.default_rng(5)
    nx, ny = 45, 14
    ke = t1.ke_gauss(np.array([[0, 0], [2.0, 0], [2.0, 0.7], [0, 0.7]]), 3.0, 0.31)
    cons = _cons(nx, ny, "cantilever")
    scale = rng.uniform(0.1, 1, nx*ny)
    u = rng.standard_normal(2*(nx + 1)*(ny + 1))
    with capi.Topo(nx, ny, ke) as t:
        t.set_cons(cons)
        t.set(capi.F_SCALE, scale)
        y = t.matvec(u).reshape(-1)
    K, free = _assembled(nx, ny, scale, ke, cons)
    ref = K @ (u*free)
    assert np.abs(y - ref).max() <= 1e-13*np.abs(ref).max()


@pytest.mark.parametrize("precond", ["jacobi", "gmg"])
@pytest.mark.parametrize("nx,ny,kind", [(60, 60, "cantilever"), (37, 23, "mbb"), (64, 20, "right_partial"),
                                        (50, 30, "scattered")])
def test_solve_matches_direct(capi, precond, nx, ny, kind):
    rng = np.random.default_rng(nx + ny)
    ke = t1.kloc_closed_form(E0, NU)
    cons = _cons(nx, ny, kind, rng)
    rho = rng.uniform(0.05, 1, nx*ny)
    scale = 1e-9 + rho**3
    loads = np.array([[(nx + 1)*(ny//3) + nx, 0.0, -1.0], [(nx + 1)*ny + nx//2, 2.0, 0.5]])
    with capi.Topo(nx, ny, ke) as t:
        t.set_cons(cons)
        t.set(capi.F_SCALE, scale)
        t.set_loads(loads)
        it, rel, _ = t.solve(capi.PRECOND_JACOBI if precond == "jacobi" else capi.PRECOND_GMG, 1e-11, 50000, warm=False)
        u = t.get(capi.F_U).reshape(-1)
        f = t.get(capi.F_LOAD).reshape(-1)
        c = t.compliance()
        assert t.equilibrium_ok()
    K, free = _assembled(nx, ny, scale, ke, cons)
    idx = np.flatnonzero(free)
    uref = np.zeros_like(u)
    uref[idx] = spsolve(K[idx][:, idx].tocsc(), f[idx])
    assert rel <= 1e-11
    assert np.abs(u - uref).max() <= 1e-7*np.abs(uref).max()
    assert abs(c - f @ uref) <= 1e-9*abs(f @ uref)
    assert np.all(u[free == 0] == 0.0)
    if precond == "gmg":
        assert it < 80, "GMG-PCG took %d iterations" % it


def test_simp_stage_kernels_match_oracle(capi):
    rng = np.random.default_rng(11)
    nx, ny = 70, 33
    ke = t1.kloc_closed_form(E0, NU)
    rho = rng.uniform(0, 1, nx*ny)
    rho[rng.integers(0, nx*ny, 40)] = 0.0
    UC = rng.standard_normal(((nx + 1)*(ny + 1), 2))*1e-11
    _, _, els = t1.rect_grid(nx, ny, nx, ny)
    with capi.Topo(nx, ny, ke) as t:
        t.set_cons(_cons(nx, ny, "cantilever"))
        t.set(capi.F_RHO, rho)
        t.set(capi.F_U, UC)
        t.simp_scale(3.0, 1e-9, 1.0)
        assert np.allclose(t.get(capi.F_SCALE), 1e-9 + rho**3*(1 - 1e-9), rtol=1e-15, atol=0)
        obj = t.simp_sensitivity(3.0, 1e-9, 1.0)
        s = t1.element_energy(UC, els, ke)
        dc_ref = (-3*rho**2*(1 - 1e-9))*s
        dc = t.get(capi.F_DC)
        assert np.allclose(dc, dc_ref, rtol=1e-13, atol=1e-13*np.abs(dc_ref).max())
        assert np.isclose(obj, ((1e-9 + rho**3*(1 - 1e-9))*s).sum(), rtol=1e-12)
        for rmin in (4.0, 2.5, 1.5):
            t.set(capi.F_DC, dc_ref)
            t.filter_sensitivity(rmin, 1.0, 1.0)
            ref = t1.sensitivity_filter_simp(nx, ny, 1.0, 1.0, rmin, rho, dc_ref)
            got = t.get(capi.F_DC)
            assert np.allclose(got, ref, rtol=1e-13, atol=1e-300), rmin
        # optimality criteria: same bisection path, same result
        dcf = t1.sensitivity_filter_simp(nx, ny, 1.0, 1.0, 4.0, rho, dc_ref)
        t.set(capi.F_DC, dcf)
        g, change, nsteps = t.oc_update(0.37)
        rho_ref, g_ref, n_ref = t1.optimality_criteria(rho, dcf, 0.37)
        assert nsteps == n_ref
        assert np.allclose(t.get(capi.F_RHO), rho_ref, rtol=1e-14, atol=0)
        assert np.isclose(g, g_ref, rtol=0, atol=1e-9)
        assert np.isclose(change, np.abs(rho_ref - rho).max(), rtol=1e-14)


@pytest.mark.parametrize("case", ["c1", "c2"])
@pytest.mark.parametrize("precond", ["gmg", "jacobi"])
def test_simp_trajectory_matches_reference(capi, case, precond):
    """Config 1 of BASELINE.json (reference tests/test_optimize.py:13-42) + a second case."""
    from solidspy_opt.optimize import SIMP
    g = golden("simp_" + case)
    hist = {"keep_rho": True}
    rho = SIMP(float(g["length"]), float(g["height"]), int(g["nx"]), int(g["ny"]), g["dirs"], g["positions"],
               int(g["niter"]), int(g["penal"]), precond=precond, history=hist, verbose=False)
    assert isinstance(rho, np.ndarray) and rho.size == int(g["nx"])*int(g["ny"])     # the reference's own asserts
    n = len(g["compliance"])
    assert len(hist["compliance"]) == n
    assert np.allclose(hist["compliance"], g["compliance"], rtol=1e-6, atol=0)
    assert np.allclose(hist["objective"], g["compliance"], rtol=1e-6, atol=0)       # u^T K u == f.u
    for k in range(n):
        assert np.abs(hist["rho"][k] - g["rho"][k]).max() <= 1e-6, "iteration %d" % k
    assert np.abs(rho - g["rho_final"]).max() <= 1e-6
    assert np.allclose(hist["g"], g["g"], rtol=0, atol=1e-5)


def _first_singular(g, n):
    C = g["C"]
    return n if np.all(np.isfinite(C)) else int(np.argmax(~np.isfinite(C))) - 1


@pytest.mark.parametrize("case", ["c1", "c2"])
def test_beso_removal_sets_match_reference(capi, case):
    from solidspy_opt.optimize import BESO
    g = golden("beso_" + case)
    hist = {}
    ELS, nodes = BESO(float(g["length"]), float(g["height"]), int(g["nx"]), int(g["ny"]), g["dirs"], g["positions"],
                      int(g["niter"]), float(g["t"]), float(g["ER"]), float(g["volfrac"]), history=hist, verbose=False)
    kept = unpack_ragged(g["kept_flat"], g["kept_off"])
    assert len(hist["kept_ids"]) == len(kept)
    for k, (a, b) in enumerate(zip(hist["kept_ids"], kept)):
        assert np.array_equal(a, b), "removal %d: %d vs %d kept" % (k, a.size, b.size)
    assert ELS.dtype == g["ELS"].dtype and np.array_equal(ELS, g["ELS"])
    assert np.array_equal(nodes, g["nodes"])
    assert np.array_equal(np.array(hist["cons"]), g["cons"])
    assert np.array_equal(hist["n_protected"], g["n_protected"])
    assert np.allclose(hist["C"], g["C"][1:1 + len(hist["C"])], rtol=1e-6)


@pytest.mark.parametrize("case", ["c1", "c2"])
@pytest.mark.parametrize("kind", ["eso_stiff", "eso_stress"])
def test_eso_removal_sets_match_reference(capi, kind, case):
    import solidspy_opt.optimize as opt
    g = golden("%s_%s" % (kind, case))
    fn = opt.ESO_stiff if kind == "eso_stiff" else opt.ESO_stress
    hist = {}
    ELS, nodes = fn(float(g["length"]), float(g["height"]), int(g["nx"]), int(g["ny"]), g["dirs"], g["positions"],
                    int(g["niter"]), float(g["RR"]), float(g["ER"]), float(g["volfrac"]), history=hist, verbose=False)
    ids = unpack_ragged(g["ids_flat"], g["ids_off"])
    n_ok = _first_singular(g, len(ids))
    assert len(hist["els_ids"]) >= n_ok
    for k, (a, b) in enumerate(zip(hist["els_ids"][:n_ok], ids[:n_ok])):
        assert np.array_equal(a, b), "removal %d: %d vs %d present" % (k, a.size, b.size)
    if n_ok == len(ids):
        assert np.array_equal(ELS, g["ELS"])
        assert np.array_equal(nodes, g["nodes"])


def _floating_count(nx, ny, ids, support_nodes):
    """Host check of topo_floating_elements: present elements in node-connected components without a support node."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    ids = np.asarray(ids)
    n0 = ids + ids//nx
    en = np.stack([n0, n0 + 1, n0 + nx + 2, n0 + nx + 1], 1)
    E = len(ids)
    A = coo_matrix((np.ones(4*E), (np.repeat(np.arange(E), 4), en.ravel())), shape=(E, (nx + 1)*(ny + 1))).tocsr()
    nc, lab = connected_components(A @ A.T, directed=False)
    sup = np.zeros((nx + 1)*(ny + 1), bool)
    sup[support_nodes] = True
    ok = np.zeros(nc, bool)
    np.logical_or.at(ok, lab, sup[en].any(axis=1))
    return int(np.count_nonzero(~ok[lab]))


def test_eso_stop_on_island_reproduces_the_reference_nan_path(capi):
    """Reference fixture eso_stress c2: the fifth removal cuts 10 elements loose, SuperLU meets a zero pivot, spsolve
    returns NaN, nothing is removed in that pass and the guard ends the loop (optimize.py:72,84).  With
    ``stop_on_island=True`` the whole history and the returned (ELS, nodes) equal the reference's."""
    import solidspy_opt.optimize as opt
    g = golden("eso_stress_c2")
    nx, ny = int(g["nx"]), int(g["ny"])
    hist = {}
    ELS, nodes = opt.ESO_stress(float(g["length"]), float(g["height"]), nx, ny, g["dirs"], g["positions"], int(g["niter"]),
                                float(g["RR"]), float(g["ER"]), float(g["volfrac"]), history=hist, verbose=False,
                                stop_on_island=True)
    ids = unpack_ragged(g["ids_flat"], g["ids_off"])
    assert len(hist["els_ids"]) == len(ids)
    for k, (a, b) in enumerate(zip(hist["els_ids"], ids)):
        assert np.array_equal(a, b), "removal %d" % k
    assert np.array_equal(np.array(hist["cons"]), g["cons"])
    assert np.array_equal(ELS, g["ELS"]) and np.array_equal(nodes, g["nodes"])
    assert hist["floating"] == [0, 0, 0, 0, 10]


def test_floating_elements_match_host_components(capi):
    """The device flood fill against scipy's connected components on the present-element sets of an ESO run that cuts
    several pieces loose (96x40 stress: the untouched reference carries on through them -- its SuperLU returns a finite
    solution there -- so the default host loop does too and the removal sets keep matching, see the mid-size test)."""
    import solidspy_opt.optimize as opt
    nx, ny = 96, 40
    hist = {"keep_floating": True}
    opt.ESO_stress(nx, ny, nx, ny, np.array([[0, -1]]), np.array([[11, 1]]), 12, 0.005, 0.05, 0.5, history=hist, verbose=False)
    support = np.arange(0, (nx + 1)*(ny + 1), nx + 1)
    want = [_floating_count(nx, ny, ids, support) for ids in hist["els_ids"]]
    assert hist["floating"] == want
    assert max(want) > 0


def test_rank_select_matches_sort(capi):
    rng = np.random.default_rng(2)
    nx, ny = 97, 41
    n = nx*ny
    vals = rng.uniform(0, 1, n)
    vals[rng.integers(0, n, 300)] = 0.0
    vals[rng.integers(0, n, 200)] = vals[7]            # duplicates around a common value
    vals[rng.integers(0, n, 50)] = 1.0
    vals[5] = -0.25
    vals[6] = 5e-324
    srt = np.sort(vals)[::-1]
    with capi.Topo(nx, ny, np.eye(8)) as t:
        t.set(capi.F_SENS, vals)
        for k in (0, 1, 49, 50, 51, n//2, int(n*0.95), n - 302, n - 2, n - 1):
            assert t.rank_select(k) == srt[k], k


def test_helper_level_api(capi):
    """Secondary boundary: reference Utils helper names on small inputs."""
    from solidspy_opt.Utils import (beam, center_els, density_filter, optimality_criteria, sparse_assem,
                                    sensitivity_elsESO, sensitivity_elsBESO, sensitivity_nodes, sensitivity_filter)
    rng = np.random.default_rng(4)
    nx, ny = 24, 10
    nodes, mats, els, loads, BC = beam(L=nx, H=ny, nx=nx, ny=ny, dirs=np.array([[0, -1]]), positions=np.array([[3, 1]]))
    centers = center_els(nodes, els)
    rho = rng.uniform(0, 1, nx*ny)
    dc = -rng.uniform(0, 1, nx*ny)
    assert np.allclose(density_filter(centers, 4.0, rho, dc), t1.sensitivity_filter_simp(nx, ny, 1.0, 1.0, 4.0, rho, dc), rtol=1e-13)
    r_new, gt = optimality_criteria(nx, ny, rho, dc, 0.1)
    r_ref, g_ref, _ = t1.optimality_criteria(rho, dc, 0.1)
    assert np.allclose(r_new, r_ref, rtol=1e-14) and np.isclose(gt, g_ref, atol=1e-10)
    assem_op, bc, neq = t1.dme(nodes[:, -2:], els)
    kloc = t1.kloc_closed_form(mats[0, 0], mats[0, 1])
    mats[:, 2] = 1e-9 + rho**3
    K = sparse_assem(els, mats, neq, assem_op, kloc)
    Kref = t1.assemble(mats[:, 2], kloc, assem_op, neq)
    v = rng.standard_normal(neq)
    assert np.allclose(K.dot(v), Kref.dot(v), rtol=1e-12, atol=1e-12*np.abs(Kref.dot(v)).max())
    assert np.isclose(K.max(), Kref.max(), rtol=1e-14)
    UC = rng.standard_normal((nodes.shape[0], 2))
    mats[:, 2] = 1
    ke = t1.ke_gauss(nodes[els[0, 3:7], 1:3], mats[0, 0], mats[0, 1])
    a = t1.strain_energy_els(UC, els, ke)
    assert np.allclose(sensitivity_elsESO(nodes, mats, els, UC), a/a.max(), rtol=1e-12)
    mask = rng.uniform(0, 1, nx*ny) > 0.3
    am = np.where(mask, a, 0.0)
    assert np.allclose(sensitivity_elsBESO(nodes, mats, els, mask, UC), am/am.max(), rtol=1e-12)
    w4, w2x, w2y, we = t1.beso_weights(nodes, els, nx, ny, 1.0)
    s = rng.uniform(0, 1, nx*ny)
    sn = sensitivity_nodes(nodes, None, centers, s)
    assert np.array_equal(sn, t1.sensitivity_nodes(nx, ny, s, w4, w2x, w2y))      # bit-exact operation order
    assert np.array_equal(sensitivity_filter(nodes, centers, sn, 1.0), t1.sensitivity_filter_beso(nx, ny, sn, we))


@pytest.mark.parametrize("nx,ny,L,H,r_factor", [(24, 10, 24.0, 10.0, 1.0), (24, 10, 24.0, 10.0, 1.6), (17, 13, 34.0, 13.0, 2.3),
                                                (9, 30, 9.0, 45.0, 3.1), (12, 12, 12.0, 12.0, 0.8)])
def test_sensitivity_filter_any_radius(capi, nx, ny, L, H, r_factor):
    """Helper-level API: ``sensitivity_filter`` (reference solver.py:212-243) accepts any ``r_min``; the optimisers only
    ever pass one element width (optimize.py:264).  Radii from 0.8 to 3.1 element widths, non-square elements,
    against the reference's loop restated literally."""
    from solidspy_opt.Utils import beam, center_els, sensitivity_filter
    rng = np.random.default_rng(nx*ny)
    nodes, mats, els, loads, BC = beam(L=L, H=H, nx=nx, ny=ny, dirs=np.array([[0, -1]]), positions=np.array([[2, 1]]))
    centers = center_els(nodes, els)
    sn = rng.uniform(0, 1, nodes.shape[0])
    r_min = r_factor*np.linalg.norm(nodes[0, 1:3] - nodes[1, 1:3])
    got = sensitivity_filter(nodes, centers, sn, r_min)
    ref = t1.sensitivity_filter_any_radius(nodes, centers, sn, r_min)
    assert np.allclose(got, ref, rtol=1e-12, atol=0)


# ---- full-size properties (BASELINE configs 3/4) ------------------------------------------------------
def test_fullsize_matvec_properties(capi):
    nx, ny = 2048, 1024
    rng = np.random.default_rng(0)
    ke = t1.kloc_closed_form(E0, NU)
    nn = (nx + 1)*(ny + 1)
    u, v = rng.standard_normal(2*nn), rng.standard_normal(2*nn)
    with capi.Topo(nx, ny, ke) as t:
        t.set_cons(_cons(nx, ny, "cantilever"))
        t.set(capi.F_SCALE, 1e-9 + rng.uniform(0, 1, nx*ny)**3)
        Ku, Kv = t.matvec(u).reshape(-1), t.matvec(v).reshape(-1)
        Kuv = t.matvec(2.0*u - 3.0*v).reshape(-1)
        # linearity, symmetry, positive definiteness, rigid translation in the null space of the free-free K
        assert np.abs(Kuv - (2*Ku - 3*Kv)).max() <= 1e-12*np.abs(Ku).max()
        assert abs(v @ Ku - u @ Kv) <= 1e-11*abs(v @ Ku)
        assert u @ Ku > 0
        t.set_cons(np.zeros((nn, 2), dtype=np.int8))
        ones = np.tile([1.0, 0.0], nn)
        assert np.abs(t.matvec(ones)).max() <= 1e-9*np.abs(Ku).max()


def test_fullsize_simp_iteration(capi):
    """2048x1024 headline mesh: two design iterations run, PCG converges, volume is conserved by OC and
    the energy form of the objective equals f.u."""
    from solidspy_opt.optimize import SIMP
    hist = {}
    nx, ny = 2048, 1024
    rho = SIMP(nx, ny, nx, ny, np.array([[0, -1]]), np.array([[ny//4, 1]]), 2, 3, history=hist, verbose=False)
    assert rho.shape == (nx*ny,) and np.all((rho >= 0) & (rho <= 1))
    assert max(hist["relres"]) <= 1e-9
    assert np.allclose(hist["objective"], hist["compliance"], rtol=1e-8)
    assert abs(rho.mean() - 0.5) < 2e-3
    assert hist["compliance"][1] < hist["compliance"][0]


def test_simp_from_model_mbb_config2(capi):
    """BASELINE config 2 shape (MBB half beam, density-filter radius 2.5 elements) through the dict entry point
    (SolidsPy README.rst:134-141 model format), against the oracle at a size it finishes in seconds."""
    from solidspy_opt.optimize import SIMP_from_model
    nx, ny = 120, 40
    nodes, mats, els, loads, _ = t1.beam(L=nx, H=ny, nx=nx, ny=ny, dirs=np.array([[0, -1]]), positions=np.array([[ny, 1]]))
    cons = np.zeros((nodes.shape[0], 2))
    col = np.arange(nodes.shape[0]) % (nx + 1)
    cons[col == 0, 0] = -1                       # symmetry plane: ux = 0
    cons[nx, 1] = -1                             # roller at the bottom-right corner: uy = 0
    load_node = (nx + 1)*ny                      # top-left corner, pushing down
    data = {"nodes": nodes[:, :3], "cons": cons, "elements": els, "mats": np.array([[206.8e9, 0.28]]),
            "loads": np.array([[load_node, 0.0, -1.0]])}
    hist = {"keep_rho": True}
    rho = SIMP_from_model(data, 12, 3, r_min=2.5, history=hist, verbose=False)
    nodes_o = nodes.copy()
    nodes_o[:, 3:5] = cons
    ref = {}
    import unittest.mock as mock
    loads_o = np.array([[load_node, 0, -1]])
    with mock.patch.object(t1, "beam", lambda **k: (nodes_o, np.tile([206.8e9, 0.28, 1.0], (nx*ny, 1)), els, loads_o, None)):
        rho_ref = t1.simp(nx, ny, nx, ny, None, None, 12, 3, r_min=2.5, history=ref)
    assert len(hist["compliance"]) == len(ref["compliance"]) == 12
    assert np.allclose(hist["compliance"], ref["compliance"], rtol=1e-6)
    for k in range(12):
        assert np.abs(hist["rho"][k] - ref["rho"][k]).max() <= 1e-6, k
    assert np.abs(rho - rho_ref).max() <= 1e-6
    # the same model on 4 row slabs (the multi-GPU algorithm; all slabs on this GPU)
    hist4 = {}
    rho4 = SIMP_from_model(data, 12, 3, r_min=2.5, history=hist4, verbose=False, slabs=4)
    assert np.allclose(hist4["compliance"], ref["compliance"], rtol=1e-6)
    assert np.abs(rho4 - rho_ref).max() <= 1e-6


def test_fullsize_beso_config3(capi):
    """BASELINE config 3 (BESO cantilever 1024x512, sensitivity filtering, device rank-select): properties that do
    not need the CPU oracle at this size -- kept counts follow the evolution ratio exactly, protected elements
    survive, removed nodes are pinned, the rank-select threshold splits the sensitivities at the requested rank."""
    from solidspy_opt.optimize import BESO
    nx, ny = 1024, 512
    hist = {"keep_sens": True}
    ELS, nodes = BESO(nx, ny, nx, ny, np.array([[0, -1]]), np.array([[ny//4, 1]]), 4, 1e-3, 0.05, 0.5, history=hist,
                      verbose=False)
    ne = nx*ny
    n_prev = ne
    for k, kept in enumerate(hist["kept_ids"]):
        target = int(n_prev*(np.full(n_prev, 1.0).sum()*(1 - 0.05))/np.full(n_prev, 1.0).sum())
        assert kept.size - hist["n_protected"][k] <= target <= kept.size, (k, kept.size, target)
        s = hist["sens"][k]
        assert np.count_nonzero(s > hist["alpha"][k]) == kept.size - hist["n_protected"][k]
        assert np.sort(s)[::-1][target] == hist["alpha"][k]                 # device radix select == full sort
        n_prev = kept.size
    assert ELS.shape[1] == 7 and ELS.shape[0] == hist["kept_ids"][-2].size
    used = np.zeros(nodes.shape[0], bool)
    last = hist["kept_ids"][-1]
    nrow = last + last//nx
    for off in (0, 1, nx + 1, nx + 2):
        used[nrow + off] = True
    assert np.all(nodes[~used, 3:5] == -1)
    load_node = (nx + 1)*(ny//4) - 1
    assert np.all(nodes[load_node, 3:5] == 0) and np.all(nodes[::nx + 1, 3:5] == -1)
    assert max(hist["relres"]) <= 1e-12


@pytest.mark.parametrize("nx,ny", [(1, 1), (2, 1), (3, 3), (4, 2), (5, 7), (16, 9), (33, 32)])
def test_simp_tiny_and_odd_meshes(capi, nx, ny):
    """Hierarchies with 1..6 levels, odd sizes (trailing single child), strips narrower than a warp."""
    from solidspy_opt.optimize import SIMP
    dirs, pos = np.array([[0, -1]]), np.array([[max(ny//4, 1), 1]])
    hist, ref = {"keep_rho": True}, {}
    rho = SIMP(nx, ny, nx, ny, dirs, pos, 4, 3, history=hist, verbose=False)
    rho_ref = t1.simp(nx, ny, nx, ny, dirs, pos, 4, 3, history=ref)
    assert len(hist["compliance"]) == len(ref["compliance"])
    assert np.allclose(hist["compliance"], ref["compliance"], rtol=1e-6)
    assert np.abs(rho - rho_ref).max() <= 1e-6


def test_simp_beam2_supports(capi):
    """`beam(..., n=2)` supports (reference beams.py:133-136): right edge clamped above H/4 and below -H/4."""
    from solidspy_opt.optimize import SIMP
    nx, ny = 48, 32
    dirs, pos = np.array([[0, -1], [1, 0]]), np.array([[16, 40], [30, 20]])
    hist, ref = {}, {}
    rho = SIMP(nx, ny, nx, ny, dirs, pos, 6, 3, n=2, history=hist, verbose=False)
    rho_ref = t1.simp(nx, ny, nx, ny, dirs, pos, 6, 3, n=2, history=ref)
    assert np.allclose(hist["compliance"], ref["compliance"], rtol=1e-6)
    assert np.abs(rho - rho_ref).max() <= 1e-6


def test_from_model_variants_match_signature_variants(capi):
    """The dict entry points (SolidsPy README model format) give exactly what the positional entry points give."""
    import solidspy_opt.optimize as opt
    g = golden("beso_c2")
    L, H, nx, ny = float(g["length"]), float(g["height"]), int(g["nx"]), int(g["ny"])
    nodes, mats, els, loads, BC = t1.beam(L=L, H=H, nx=nx, ny=ny, dirs=g["dirs"], positions=g["positions"])
    data = {"nodes": nodes[:, :3], "cons": nodes[:, 3:5], "elements": els, "mats": np.array([[206.8e9, 0.28]]), "loads": loads}
    ELS, nd = opt.BESO_from_model(data, int(g["niter"]), float(g["t"]), float(g["ER"]), float(g["volfrac"]), verbose=False)
    assert np.array_equal(ELS, g["ELS"]) and np.array_equal(nd, g["nodes"])
    for kind, fn in (("eso_stiff", opt.ESO_stiff_from_model), ("eso_stress", opt.ESO_stress_from_model)):
        gg = golden(kind + "_c1")
        n1, m1, e1, l1, b1 = t1.beam(L=60, H=60, nx=60, ny=60, dirs=gg["dirs"], positions=gg["positions"])
        d1 = {"nodes": n1[:, :3], "cons": n1[:, 3:5], "elements": e1, "mats": np.array([[206.8e9, 0.28]]), "loads": l1}
        ELS, nd = fn(d1, 50, 0.005, 0.05, 0.5, verbose=False)
        assert np.array_equal(ELS, gg["ELS"]) and np.array_equal(nd, gg["nodes"])


def test_slab_matvec_and_pcg_on_two_gpus(capi):
    """Row-slab building block over NCCL (needs >= 2 GPUs on the box; skipped on the single-GPU test run)."""
    import subprocess, sys, os, json
    if capi.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    env = dict(os.environ, DIST_BIG="0")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(root, "tests", "dist_gpu_check.py")],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    assert json.loads(line)["ok"]


def test_slab_simp_on_two_gpus(capi):
    """The full row-slab design iteration over NCCL against the single-GPU path (needs >= 2 GPUs on the box)."""
    import subprocess, sys, os, json
    if capi.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29519", os.path.join(root, "tests", "dist_simp_check.py")],
                       capture_output=True, text=True, env=dict(os.environ, DIST_BIG=""), timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    assert json.loads(line)["ok"]


@pytest.mark.parametrize("knob", ["TOPO_LVL_TILED", "TOPO_SETUP_FAST"])
def test_round2_kernel_variants_are_bit_identical(capi, monkeypatch, knob):
    """The tile level kernels (transfers inside) against the separate transfer kernels, and the branch-free Galerkin
    set-up kernels against the general ones: same operations in the same order, so the trajectories must agree BIT FOR
    BIT -- compliance history, PCG iteration counts and the final densities (mesh with odd level sizes on purpose)."""
    nx, ny = 300, 140
    d, p = np.array([[0, -1]]), np.array([[35, 1]])
    runs = []
    for value in ("1", "0"):
        monkeypatch.setenv(knob, value)                 # read once per handle in topo_create
        from solidspy_opt.optimize import SIMP
        hist = {}
        rho = SIMP(float(nx), float(ny), nx, ny, d, p, 6, 3, history=hist, verbose=False)
        runs.append((hist["compliance"], hist["cg_iters"], rho))
    assert runs[0][1] == runs[1][1]
    assert runs[0][0] == runs[1][0]
    assert np.array_equal(runs[0][2], runs[1][2])


def test_simp_step_with_host_buffers_equals_device_resident_path(capi):
    """topo_simp_step_host (the e2e path of the bench: density upload, step, download overlapping the equilibrium guard on
    the side stream) against the device-resident step: identical results, and the downloaded densities are the device's."""
    nx, ny = 200, 96
    ke = t1.kloc_closed_form(E0, NU)
    pr = capi.SimpParams(penal=3.0, emin=1e-9, emax=1.0, rmin=2.5, dx=1.0, dy=1.0, move=0.2, rtol=1e-10,
                         precond=capi.PRECOND_GMG, maxit=2000, warm=1)
    out = []
    for host in (False, True):
        with capi.Topo(nx, ny, ke) as t:
            t.set_cons(_cons(nx, ny, "cantilever"))
            t.set_loads(np.array([[(nx + 1)*(ny//4) - 1, 0.0, -1.0]]))
            t.fill(capi.F_RHO, 0.5)
            a, b = np.full(nx*ny, 0.5), np.empty(nx*ny)
            g, hist = 0.0, []
            for _ in range(4):
                if host:
                    g, r = t.simp_step(pr, g, rho_in=a, rho_out=b)
                    assert np.array_equal(b, t.get(capi.F_RHO))
                    a, b = b, a
                else:
                    g, r = t.simp_step(pr, g)
                hist.append((r.compliance, r.cg_iters, r.oc_steps, r.change))
            out.append((hist, t.get(capi.F_RHO)))
    assert out[0][0] == out[1][0]
    assert np.array_equal(out[0][1], out[1][1])


def test_handle_lifecycle_and_errors(capi):
    """Repeated create/destroy does not leak device memory; bad arguments return error codes, not crashes."""
    import ctypes as C
    import torch
    free0 = torch.cuda.mem_get_info()[0]
    for _ in range(20):
        with capi.Topo(256, 128, t1.kloc_closed_form(E0, NU)) as t:
            t.set_cons(_cons(256, 128, "cantilever"))
            t.set_loads(np.array([[257*32 + 256, 0.0, -1.0]]))
            t.fill(capi.F_RHO, 0.5)
            t.simp_scale(3.0, 1e-9, 1.0)
            t.solve(capi.PRECOND_GMG, 1e-8, 500, warm=False)
    free1 = torch.cuda.mem_get_info()[0]
    assert free0 - free1 < 64 << 20                      # nothing but allocator noise left behind
    with capi.Topo(8, 4, t1.kloc_closed_form(E0, NU)) as t:
        with pytest.raises(capi.TopoError) as ei:
            t.solve()                                    # BC flags / loads not set
        assert ei.value.code == capi.ESTATE
        t.set_cons(_cons(8, 4, "cantilever"))
        with pytest.raises(capi.TopoError):
            t.set_loads(np.array([[10**6, 0.0, 1.0]]))  # node id out of range
        with pytest.raises(capi.TopoError):
            t.rank_select(8*4)                           # rank outside [0, N_e)
        with pytest.raises(capi.TopoError):
            t.filter_sensitivity(100.0, 1.0, 1.0)        # radius wider than the stencil table
        # an unloaded structure: zero right-hand side converges immediately to u = 0
        t.set_loads(np.zeros((0, 3)))
        it, rel, code = t.solve(capi.PRECOND_GMG, 1e-10, 10, warm=False)
        assert it == 0 and code == capi.OK and np.all(t.get(capi.F_U) == 0)
    # maxit exhausted -> TOPO_ENOCONV, results still readable
    with capi.Topo(64, 32, t1.kloc_closed_form(E0, NU)) as t:
        t.set_cons(_cons(64, 32, "cantilever"))
        t.set_loads(np.array([[65*8 + 64, 0.0, -1.0]]))
        it, rel, code = t.solve(capi.PRECOND_JACOBI, 1e-12, 5, warm=False, raise_on_noconv=False)
        assert code == capi.ENOCONV and it == 5 and rel > 1e-12


# ---- row-slab decomposition: several slabs on ONE GPU exercise every kernel / segment of the multi-GPU path ----
def _slab_case(nx, ny, world, K, rho=None, seed=3, transport="p2p"):
    """(SlabSimp, single-GPU Topo) on the same cantilever with the same (random) density."""
    from solidspy_opt import _capi
    from solidspy_opt.Utils.beams import beam
    from solidspy_opt.distributed import SlabSimp
    from solidspy_opt.optimize import kloc_simp
    nodes, mats, els, loads, _ = beam(L=float(nx), H=float(ny), nx=nx, ny=ny, dirs=np.array([[0, -1]]),
                                      positions=np.array([[ny//2 + 1, 1]]), n=1)     # mid-height of the free end
    ke = kloc_simp(mats[0, 0], mats[0, 1])
    if rho is None:
        rho = np.random.default_rng(seed).uniform(0.05, 1.0, nx*ny)
    sim = SlabSimp(nx, ny, ke, nodes[:, 3:5], loads, rmin=2.5, world=world, K=K, rtol=1e-10, transport=transport)
    assert sim.p2p == (transport == "p2p"), sim.p2p_error
    sim.set_rho(rho)
    topo = _capi.Topo(nx, ny, ke)
    topo.set_cons(nodes[:, 3:5])
    topo.set_loads(loads)
    topo.set(_capi.F_RHO, rho)
    return sim, topo


@pytest.mark.parametrize("nx,ny,world,tail_max", [(128, 64, 2, 10000), (96, 64, 4, 500), (256, 128, 4, 600),
                                                  (130, 96, 3, 10000), (512, 256, 8, 300)])
@pytest.mark.parametrize("transport", ["p2p", "copies"])
def test_slab_solve_equals_single_gpu(capi, monkeypatch, nx, ny, world, tail_max, transport):
    """Same preconditioner, same Krylov iteration: the slab solve reproduces topo_solve on the whole mesh.
    ``tail_max`` (nodes of the first tail level) moves the first global level K: 1, 2, 3, 1, 5.  ``transport``: rows,
    level-K pieces and scalars through the peer-memory kernels, or moved by the host between segments (the NCCL
    path of a multi-process run; device copies here)."""
    import torch
    from solidspy_opt import _capi
    monkeypatch.setenv("TOPO_TAIL_MAX", str(tail_max))
    sim, topo = _slab_case(nx, ny, world, 0, transport=transport)
    try:
        topo.simp_scale(3.0, 1e-9, 1.0)
        it1, rel1, _ = topo.solve(_capi.PRECOND_GMG, 1e-10, 2000, warm=False)
        u1 = topo.get(_capi.F_U).reshape(ny + 1, nx + 1, 2)
        with torch.cuda.stream(sim.stream):
            for t in sim.slabs:
                t.simp_scale(3.0, 1e-9, 1.0)
            sim.setup()
            it2, rel2 = sim.solve(warm=False)
        u2 = sim.local_u()
        assert rel2 <= 1e-10
        assert abs(it1 - it2) <= 1, (it1, it2)
        assert np.abs(u1 - u2).max() <= 1e-7*np.abs(u1).max()
        # interface rows are bit-identical on both neighbours (accumulated form)
        for a, b in zip(sim.slabs[:-1], sim.slabs[1:]):
            ua = a.get(_capi.F_U).reshape(sim.nyl + 1, nx + 1, 2)[-1]
            ub = b.get(_capi.F_U).reshape(sim.nyl + 1, nx + 1, 2)[0]
            assert np.array_equal(ua, ub)
    finally:
        sim.close()
        topo.close()


@pytest.mark.parametrize("world,K,transport", [(2, 0, "p2p"), (4, 2, "p2p"), (4, 2, "copies")])
def test_slab_design_iterations_equal_single_gpu(capi, world, K, transport):
    """Five full design iterations (solve, sensitivity, filter with ghost rows, OC with all-reduced sums)."""
    from solidspy_opt import _capi
    nx, ny = 96, 64
    sim, topo = _slab_case(nx, ny, world, K, rho=np.full(nx*ny, 0.5), transport=transport)
    try:
        params = _capi.SimpParams(penal=3.0, emin=1e-9, emax=1.0, rmin=2.5, dx=1.0, dy=1.0, move=0.2, rtol=1e-10,
                                  precond=_capi.PRECOND_GMG, maxit=2000, warm=1)
        g1 = g2 = 0.0
        for k in range(5):
            g1, res = topo.simp_step(params, g1)
            g2, r2 = sim.step(g2)
            assert r2["compliance"] == pytest.approx(res.compliance, rel=1e-8)
            assert r2["objective"] == pytest.approx(res.objective, rel=1e-8)
            assert r2["oc_steps"] == res.oc_steps
            assert g2 == pytest.approx(g1, abs=1e-6)
            assert np.abs(sim.local_rho().reshape(-1) - topo.get(_capi.F_RHO)).max() <= 1e-8, "iteration %d" % k
    finally:
        sim.close()
        topo.close()


@pytest.mark.parametrize("case,slabs", [("c1", 2), ("c2", 2)])
def test_slab_simp_trajectory_matches_reference(capi, case, slabs):
    """The reference goldens through the slab path (what runs one slab per GPU under torch.distributed)."""
    from solidspy_opt.optimize import SIMP
    g = golden("simp_" + case)
    hist = {"keep_rho": True}
    rho = SIMP(float(g["length"]), float(g["height"]), int(g["nx"]), int(g["ny"]), g["dirs"], g["positions"],
               int(g["niter"]), int(g["penal"]), history=hist, verbose=False, slabs=slabs)
    n = len(g["compliance"])
    assert len(hist["compliance"]) == n
    assert np.allclose(hist["compliance"], g["compliance"], rtol=1e-6, atol=0)
    for k in range(n):
        assert np.abs(hist["rho"][k] - g["rho"][k]).max() <= 1e-6, "iteration %d" % k
    assert np.abs(rho - g["rho_final"]).max() <= 1e-6
    assert np.allclose(hist["g"], g["g"], rtol=0, atol=1e-5)


@pytest.mark.parametrize("nx,ny,world,kind", [(480, 160, 5, "mbb"), (120, 64, 4, "right_partial"), (96, 64, 2, "scattered")])
def test_slab_design_iterations_other_supports(capi, nx, ny, world, kind):
    """Row slabs with constraints that are NOT the same on every slab (MBB rollers + one support in slab 0 -- the
    BASELINE config-2 mesh -- beam_2-style partial clamping, scattered pinned DOFs) and a load on an interface row."""
    from solidspy_opt import _capi
    from solidspy_opt.distributed import SlabSimp
    from solidspy_opt.optimize import kloc_simp
    cons = _cons(nx, ny, kind, np.random.default_rng(5))
    nyl = ny//world
    loads = np.array([[(nx + 1)*nyl + nx//2, 0.0, -1.0], [(nx + 1)*ny + 3*nx//4, 0.5, -0.25]])   # first one on an interface row
    ke = kloc_simp(206.8e9, 0.28)
    sim = SlabSimp(nx, ny, ke, cons, loads, rmin=2.5, world=world, rtol=1e-10)
    topo = _capi.Topo(nx, ny, ke)
    try:
        topo.set_cons(cons)
        topo.set_loads(loads)
        topo.fill(_capi.F_RHO, 0.5)
        params = _capi.SimpParams(penal=3.0, emin=1e-9, emax=1.0, rmin=2.5, dx=1.0, dy=1.0, move=0.2, rtol=1e-10,
                                  precond=_capi.PRECOND_GMG, maxit=5000, warm=1)
        g1 = g2 = 0.0
        for k in range(4):
            g1, res = topo.simp_step(params, g1)
            g2, r2 = sim.step(g2)
            assert abs(r2["cg_iters"] - res.cg_iters) <= 1, (k, r2["cg_iters"], res.cg_iters)
            assert r2["compliance"] == pytest.approx(res.compliance, rel=1e-8)
            assert r2["oc_steps"] == res.oc_steps
            assert np.abs(sim.local_rho().reshape(-1) - topo.get(_capi.F_RHO)).max() <= 1e-7, "iteration %d" % k
    finally:
        sim.close()
        topo.close()


@pytest.mark.parametrize("nx,ny,niter", [(128, 64, 10), (200, 60, 8)])
def test_beso_removal_sets_match_oracle_midsize(capi, nx, ny, niter):
    """BESO beyond the reference fixtures: removal sets against the CPU restatement (direct solver) on meshes where
    a loosely converged PCG or a different summation order would reorder near-equal sensitivities."""
    from solidspy_opt.optimize import BESO
    # load at quarter height: a load on the mid-axis makes the problem mirror symmetric, the sensitivities of mirror
    # elements exactly equal in exact arithmetic, and which of the two is removed a matter of solver round-off (the
    # reference itself is not reproducible across LAPACK/SuperLU builds there)
    dirs, pos = np.array([[0, -1]]), np.array([[ny//4 + 1, 1]])
    hist, ref = {}, {}
    ELS, nodes = BESO(nx, ny, nx, ny, dirs, pos, niter, 1e-3, 0.05, 0.5, history=hist, verbose=False)
    ELS_ref, nodes_ref = t1.beso(nx, ny, nx, ny, dirs, pos, niter, 1e-3, 0.05, 0.5, history=ref)
    assert len(hist["kept_ids"]) == len(ref["kept_ids"])
    for k, (a, b) in enumerate(zip(hist["kept_ids"], ref["kept_ids"])):
        assert np.array_equal(a, b), "removal %d: %d vs %d kept, %d differ" % (k, a.size, b.size, np.setxor1d(a, b).size)
    assert np.array_equal(ELS, ELS_ref)
    assert np.array_equal(nodes, nodes_ref)


@pytest.mark.parametrize("kind,nx,ny,niter", [("eso_stiff", 96, 40, 30), ("eso_stress", 96, 40, 12), ("eso_stress", 150, 50, 12)])
def test_eso_removal_sets_match_oracle_midsize(capi, kind, nx, ny, niter):
    """ESO beyond the reference fixtures, against the CPU restatement (direct solver); load off the mirror axis."""
    import solidspy_opt.optimize as opt
    fn, fn_ref = (opt.ESO_stiff, t1.eso_stiff) if kind == "eso_stiff" else (opt.ESO_stress, t1.eso_stress)
    dirs, pos = np.array([[0, -1]]), np.array([[ny//4 + 1, 1]])
    hist, ref = {}, {}
    ELS, nodes = fn(nx, ny, nx, ny, dirs, pos, niter, 0.005, 0.05, 0.5, history=hist, verbose=False)
    with np.errstate(all="ignore"):
        ELS_ref, nodes_ref = fn_ref(nx, ny, nx, ny, dirs, pos, niter, 0.005, 0.05, 0.5, history=ref)
    C = np.array(ref["C"])
    n_ok = len(ref["els_ids"]) if np.all(np.isfinite(C)) else int(np.argmax(~np.isfinite(C))) - 1
    assert n_ok >= 2 and len(hist["els_ids"]) >= n_ok
    for k, (a, b) in enumerate(zip(hist["els_ids"][:n_ok], ref["els_ids"][:n_ok])):
        assert np.array_equal(a, b), "removal %d: %d vs %d present, %d differ" % (k, a.size, b.size, np.setxor1d(a, b).size)
    if n_ok == len(ref["els_ids"]):
        assert np.array_equal(ELS, ELS_ref)
        assert np.array_equal(nodes, nodes_ref)


@pytest.mark.parametrize("nx,ny,L,H,penal,r_factor,n,pos", [
    (48, 24, 48.0, 24.0, 2, 4, 1, [[7, 1]]),            # penal = 2 (numpy's ** 2 fast path, optimize.py:444)
    (40, 30, 40.0, 30.0, 4, 3, 1, [[20, 1], [5, 1]]),   # penal = 4, two loads, r = 3 elements
    (36, 18, 72.0, 36.0, 3, 2, 1, [[10, 3]]),           # 2 x 2 elements, load off the edge column
    (45, 15, 45.0, 15.0, 3, 5, 2, [[8, 1]]),            # beam_2 supports, r = 5 elements
    (33, 21, 33.0, 21.0, 1, 4, 1, [[11, 1]]),           # penal = 1 (no penalisation)
])
def test_simp_parameter_sweep_matches_oracle(capi, nx, ny, L, H, penal, r_factor, n, pos):
    """SIMP outside the fixtures' parameters: penalisation exponents, filter radii, element sizes, support kinds,
    several loads -- 8 design iterations each against the CPU restatement (direct solver)."""
    from solidspy_opt.optimize import SIMP
    pos = np.array(pos)
    dirs = np.tile([0, -1], (pos.shape[0], 1))
    hist, ref = {"keep_rho": True}, {}
    rho = SIMP(L, H, nx, ny, dirs, pos, 8, penal, n=n, r_factor=r_factor, history=hist, verbose=False)
    rho_ref = t1.simp(L, H, nx, ny, dirs, pos, 8, penal, n=n, r_factor=float(r_factor), history=ref)
    assert len(hist["compliance"]) == len(ref["compliance"])
    assert np.allclose(hist["compliance"], ref["compliance"], rtol=1e-6)
    for k in range(len(ref["compliance"])):
        assert np.abs(hist["rho"][k] - ref["rho"][k]).max() <= 1e-6, k
    assert np.abs(rho - rho_ref).max() <= 1e-6
    # ... and on two row slabs where the mesh allows it (slab height even)
    if ny % 4 == 0:
        hist2 = {}
        rho2 = SIMP(L, H, nx, ny, dirs, pos, 8, penal, n=n, r_factor=r_factor, history=hist2, verbose=False, slabs=2)
        assert np.allclose(hist2["compliance"], ref["compliance"], rtol=1e-6)
        assert np.abs(rho2 - rho_ref).max() <= 1e-6


@pytest.mark.parametrize("nx,ny,ER,volfrac,pos", [(40, 24, 0.02, 0.6, [[7, 1]]), (50, 20, 0.08, 0.4, [[6, 1], [15, 1]]),
                                                  (64, 32, 0.05, 0.5, [[9, 2]])])
def test_beso_parameter_sweep_matches_oracle(capi, nx, ny, ER, volfrac, pos):
    """BESO with other evolution ratios / volume targets / loads: bit-identical kept sets against the CPU restatement."""
    from solidspy_opt.optimize import BESO
    pos = np.array(pos)
    dirs = np.tile([0, -1], (pos.shape[0], 1))
    hist, ref = {}, {}
    ELS, nodes = BESO(nx, ny, nx, ny, dirs, pos, 12, 1e-3, ER, volfrac, history=hist, verbose=False)
    ELS_ref, nodes_ref = t1.beso(nx, ny, nx, ny, dirs, pos, 12, 1e-3, ER, volfrac, history=ref)
    assert len(hist["kept_ids"]) == len(ref["kept_ids"])
    for k, (a, b) in enumerate(zip(hist["kept_ids"], ref["kept_ids"])):
        assert np.array_equal(a, b), "removal %d: %d differ" % (k, np.setxor1d(a, b).size)
    assert np.array_equal(ELS, ELS_ref) and np.array_equal(nodes, nodes_ref)


# ---- plot=True outputs: the fields the reference hands to its plots, and the plot calls themselves ----
def _compat_pos():
    import os
    import sys
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle", "compat")
    if p not in sys.path:
        sys.path.insert(0, p)
    import solidspy.postprocesor as pos
    return pos


@pytest.mark.parametrize("kind", ["eso_stress", "eso_stiff", "beso"])
def test_plot_fields_match_strain_nodes(capi, kind):
    """UCI/E_nodesI/S_nodesI and UC/E_nodes/S_nodes (optimize.py:62-63,85-86,162-163,182-183,261-262,299-300) against
    the oracle's strain_nodes fed with the same displacements: ESO averages over present elements (orphan nodes
    NaN), BESO over the full mesh."""
    import solidspy_opt.optimize as opt
    pos = _compat_pos()
    nx, ny = 30, 20
    dirs, positions = np.array([[0, -1]]), np.array([[5, 1]])
    hist = {"keep_fields": True}
    if kind == "beso":
        ELS, nodes = opt.BESO(30.0, 20.0, nx, ny, dirs, positions, 4, 0.001, 0.05, 0.5, history=hist, verbose=False)
    else:
        fn = opt.ESO_stress if kind == "eso_stress" else opt.ESO_stiff
        ELS, nodes = fn(30.0, 20.0, nx, ny, dirs, positions, 4, 0.02, 0.03, 0.5, history=hist, verbose=False)
    f = hist["fields"]
    from solidspy_opt.Utils.beams import beam
    els_full = beam(L=30.0, H=20.0, nx=nx, ny=ny, dirs=dirs, positions=positions, n=1)[2]
    mats = np.array([[E0, NU]])
    nodes0 = nodes.copy()
    E_i, S_i = pos.strain_nodes(nodes0, els_full, mats, f["UCI"])
    # 1e-12 of the field's magnitude (components that cancel to ~0 carry the rounding of the larger terms)
    tol_e, tol_s = 1e-12*np.abs(E_i).max(), 1e-12*np.abs(S_i).max()
    assert np.allclose(f["E_nodesI"], E_i, rtol=1e-12, atol=tol_e)
    assert np.allclose(f["S_nodesI"], S_i, rtol=1e-12, atol=tol_s)
    els_last = els_full if kind == "beso" else ELS
    assert kind == "beso" or ELS.shape[0] < els_full.shape[0]          # something was removed
    E_l, S_l = pos.strain_nodes(nodes0, els_last, mats, f["UC"])
    assert np.array_equal(np.isnan(f["E_nodes"]), np.isnan(E_l))
    assert np.allclose(f["E_nodes"], E_l, rtol=1e-12, atol=tol_e, equal_nan=True)
    assert np.allclose(f["S_nodes"], S_l, rtol=1e-12, atol=tol_s, equal_nan=True)
    clamped = nodes[:, 1] == nodes[:, 1].min()                           # the left edge (orphan nodes are pinned later)
    assert f["UC"].shape == (nodes.shape[0], 2) and np.all(f["UC"][clamped] == 0.0)


def test_plot_true_draws_the_reference_figures(capi, monkeypatch):
    """plot=True (optimize.py:101-109): two fields_plot calls (8 contour figures each) + the design picture, on the
    triangulations of the initial and of the final element set.  matplotlib is not installed: a recording stand-in."""
    import solidspy_opt.optimize as opt
    from test_host import fake_matplotlib
    calls = fake_matplotlib(monkeypatch)
    nx, ny = 24, 12
    ELS, nodes = opt.ESO_stiff(24.0, 12.0, nx, ny, np.array([[0, -1]]), np.array([[3, 1]]), 3, 0.01, 0.02, 0.5,
                               plot=True, verbose=False)
    tris = [c for c in calls if c[0] == "Triangulation"]
    cont = [c for c in calls if c[0] == "tricontourf"]
    assert len(cont) == 17 and len(tris) == 2*3 + 1
    assert tris[0][1] == 2*nx*ny and tris[-1][1] == 2*ELS.shape[0]
