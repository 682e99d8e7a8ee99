"""Host-side logic of the drop-in package (no GPU): mesh/numbering bit-exactness, bookkeeping helpers against
brute-force restatements of the reference loops, C-ABI export list, and the loud failure without a device."""
import ctypes
import os
import re

import numpy as np
import pytest

import t1
from solidspy_opt import _capi
from solidspy_opt.Utils import beams, solver

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.mark.parametrize("L,H,nx,ny", [(60, 60, 60, 60), (60, 30, 40, 20), (7.3, 2.1, 13, 5), (1, 1, 1, 1), (10, 3, 64, 7)])
@pytest.mark.parametrize("n", [1, 2])
def test_beam_is_bit_exact(L, H, nx, ny, n):
    """Connectivity, DOF flags, load rows, BC ids: identical arrays (dtype, shape, bits) to the mgrid-based oracle."""
    dirs, pos = np.array([[0, -1], [2, 3]]), np.array([[3, 1], [2, 2]])
    a = t1.beam(L=L, H=H, nx=nx, ny=ny, dirs=dirs, positions=pos, n=n)
    b = beams.beam(L=L, H=H, nx=nx, ny=ny, dirs=dirs, positions=pos, n=n)
    for p, q in zip(a, b):
        assert p.dtype == q.dtype and p.shape == q.shape and np.array_equal(p, q)


def test_beam_matches_reference_fixture_facts():
    # load node 914 at (30, -16); 61 clamped nodes 0, 61, 122, ...
    nodes, mats, els, loads, BC = beams.beam(L=60, H=60, nx=60, ny=60, dirs=np.array([[0, -1]]), positions=np.array([[15, 1]]))
    assert loads.dtype.kind == "i" and list(loads[0]) == [914, 0, -1]
    assert tuple(nodes[914, 1:3]) == (30.0, -16.0)
    assert np.array_equal(BC, np.arange(61)*61.0)
    assert np.array_equal(mats[0], [206.8e9, 0.28, 1.0])
    assert beams.beam(n=3) is None                      # the reference's `match` falls through


def _brute_protect_els(els, nels, loads, BC):           # reference solver.py:26-33, literal loops
    mask = np.zeros(nels, dtype=bool)
    for p in np.hstack((loads[:, 0], BC)).astype(int):
        idx = np.argwhere(els[:, -4:] == p)[:, 0]
        mask[els[idx, 0]] = True
    return mask


def _brute_del_node(nodes, els, loads, BC):             # reference solver.py:53-58
    prot = np.hstack((loads[:, 0], BC)).astype(int)
    for n in nodes[:, 0]:
        if n not in els[:, -4:]:
            nodes[int(n), -2:] = -1
        elif n not in prot and n in els[:, -4:]:
            nodes[int(n), -2:] = 0


def test_bookkeeping_helpers_match_reference_loops():
    rng = np.random.default_rng(0)
    nx, ny = 9, 6
    nodes, mats, els, loads, BC = beams.beam(L=nx, H=ny, nx=nx, ny=ny, dirs=np.array([[0, -1]]), positions=np.array([[2, 1]]))
    for _ in range(5):
        mask = rng.uniform(size=nx*ny) > 0.4
        assert np.array_equal(solver.protect_els(els[~mask], nx*ny, loads, BC), _brute_protect_els(els[~mask], nx*ny, loads, BC))
        assert np.array_equal(solver.protect_elsESO(els[mask], loads, BC), ~_brute_protect_els(els[mask], nx*ny, loads, BC)[els[mask, 0]])
        a, b = nodes.copy(), nodes.copy()
        a[rng.integers(0, a.shape[0], 7), 3:] = -1
        b[:] = a
        solver.del_node(a, els[mask], loads, BC)
        _brute_del_node(b, els[mask], loads, BC)
        assert np.array_equal(a, b)
        c = nodes.copy()
        solver.del_nodeESO(c, els[mask])
        used = np.isin(np.arange(nodes.shape[0]), els[mask][:, -4:])
        assert np.array_equal(c[:, 3:] == -1, (~used)[:, None] | (nodes[:, 3:] == -1))
    adj = solver.adjacency_nodes(nodes, els)
    for n in (0, 5, nx + 1, 2*(nx + 1) + 3, nodes.shape[0] - 1):
        assert np.array_equal(adj[n], np.argwhere(els[:, -4:] == n)[:, 0])
    assert np.array_equal(solver.center_els(nodes, els), t1.center_els(nodes, els))
    assert np.array_equal(solver.volume(els, 9.0, 6.0, nx, ny), t1.volume(els, 9.0, 6.0, nx, ny))
    En, Sn = rng.standard_normal((nodes.shape[0], 3)), rng.standard_normal((nodes.shape[0], 3))
    Ee, Se = solver.strain_els(els, En, Sn)
    assert np.allclose(Se[7], Sn[els[7, 3:]].sum(0)/4) and np.allclose(Ee[0], En[els[0, 3:]].sum(0)/4)


def test_element_matrices():
    assert np.array_equal(solver.kloc_simp(206.8e9, 0.28), t1.kloc_closed_form(206.8e9, 0.28))
    coord = np.array([[0, 0], [1.5, 0], [1.5, 0.7], [0, 0.7]])
    assert np.allclose(solver.ke_quad4(coord, 3.0, 0.3), t1.ke_gauss(coord, 3.0, 0.3), rtol=1e-14)
    nodes, mats, els, loads, BC = beams.beam(L=30, H=15, nx=20, ny=10, dirs=np.array([[0, -1]]), positions=np.array([[2, 1]]))
    for a, b in zip(solver.beso_weights(nodes, els, 20, 10, 1.5), t1.beso_weights(nodes, els, 20, 10, 1.5)):
        assert np.array_equal(a, b)


def test_grid_introspection_rejects_unstructured_input():
    nodes, mats, els, loads, BC = beams.beam(L=8, H=4, nx=8, ny=4, dirs=np.array([[0, -1]]), positions=np.array([[1, 1]]))
    assert solver.grid_shape_from_els(els) == (8, 4)
    assert solver.grid_shape_from_nodes(nodes)[:2] == (8, 4)
    bad = els.copy()
    bad[3, 4] += 1
    with pytest.raises(ValueError):
        solver.grid_shape_from_els(bad)


@pytest.mark.parametrize("L,H,nx,ny,n", [(60, 60, 60, 60, 1), (60.0, 30.0, 40, 20, 2), (7.5, 3.0, 13, 5, 1), (10, 10, 1, 1, 2)])
def test_beam_bc_equals_beam(L, H, nx, ny, n):
    """The lean set-up of SIMP() carries exactly beam()'s BC flags, loads and node spacing."""
    from solidspy_opt.Utils.beams import beam, beam_bc
    dirs, pos = np.array([[0, -2], [1, 0]]), np.array([[max(ny//3, 1), 1], [ny, 1]])
    nodes, mats, els, loads, BC = beam(L=L, H=H, nx=nx, ny=ny, dirs=dirs, positions=pos, n=n)
    cons, loads2, spacing = beam_bc(L, H, nx, ny, dirs, pos, n)
    assert cons.dtype == np.int8 and np.array_equal(cons, nodes[:, 3:5])
    assert loads2.dtype == loads.dtype and np.array_equal(loads2, loads)
    assert spacing == np.linalg.norm(nodes[0, 1:3] - nodes[1, 1:3])
    assert (mats[0, 0], mats[0, 1]) == (206.8e9, 0.28)


def test_library_exports_every_declared_symbol(lib_built):
    """Every function include/topo_b200.h declares is exported by the shared library (no compute calls)."""
    header = open(os.path.join(ROOT, "include", "topo_b200.h")).read()
    declared = set(re.findall(r"\b(topo_[a-z0-9_]+)\s*\(", header)) - {"topo_t"}
    lib = ctypes.CDLL(lib_built)
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    assert declared >= set(_capi.SYMBOLS)
    assert lib.topo_abi_version() == 1


def test_no_cpu_fallback(lib_built):
    """Without a CUDA device the product path must fail loudly, never compute on the host."""
    try:
        n = _capi.device_count()
    except _capi.TopoError as e:
        assert e.code == _capi.ENODEV
        with pytest.raises(_capi.TopoError):
            _capi.Topo(4, 4, t1.kloc_closed_form(1.0, 0.3))
        from solidspy_opt.optimize import SIMP
        with pytest.raises(_capi.TopoError):
            SIMP(4, 4, 4, 4, np.array([[0, -1]]), np.array([[1, 1]]), 2, 3, verbose=False)
    else:
        assert n >= 1


def test_create_rejects_non_rectangular_element_matrix(lib_built):
    ke = np.arange(64, dtype=float)
    with pytest.raises(_capi.TopoError) as ei:
        _capi.Topo(4, 4, ke)
    assert ei.value.code in (_capi.EINVAL, _capi.ENODEV)


def test_slab_path_has_no_cpu_fallback_and_validates_its_arguments(lib_built):
    """The row-slab entry point (multi-GPU algorithm) fails loudly without a device and rejects meshes it cannot split."""
    from solidspy_opt.distributed import SlabSimp
    from solidspy_opt.optimize import SIMP
    ke = t1.kloc_closed_form(1.0, 0.3)
    cons = np.zeros((5*8, 2), dtype=np.int8)
    with pytest.raises(ValueError):                                       # 7 element rows cannot form 2 equal slabs
        SlabSimp(4, 7, ke, cons, np.zeros((0, 3)), world=2)
    try:
        _capi.device_count()
    except _capi.TopoError:
        with pytest.raises(Exception):
            SIMP(8, 8, 8, 8, np.array([[0, -1]]), np.array([[1, 1]]), 2, 3, verbose=False, slabs=2)


def test_header_documents_what_every_slab_segment_hands_to_the_host():
    """Each segment call of the row-slab API names the movement the host has to perform next (exchange kind /
    gather / all-reduce) -- the contract a binding in another host language relies on."""
    header = open(os.path.join(ROOT, "include", "topo_b200.h")).read()
    for fn, needs in [("topo_dist_pcg_matvec", "TOPO_X_AP"), ("topo_dist_pcg_update", "TOPO_X_LEVEL_B"),
                      ("topo_dist_vcycle_down", "TOPO_G_BK"), ("topo_dist_vcycle_up", "TOPO_X_LEVEL_X2"),
                      ("topo_dist_pcg_smooth", "TOPO_S_RR_RZ"), ("topo_dist_solve_begin", "TOPO_S_BB"),
                      ("topo_dist_filter_pack", "TOPO_X_FILTER")]:
        line = [l for l in header.splitlines() if fn + "(" in l]
        i = header.index(fn + "(")
        assert needs in header[max(0, i - 400):i + 400], (fn, needs)


# ---- plot plumbing: no matplotlib in this image, so a recording stand-in ------------------------------
def fake_matplotlib(monkeypatch):
    """Installs stand-ins for matplotlib / matplotlib.pyplot / matplotlib.tri that record the calls the plot
    helpers make; returns the call list."""
    import sys
    import types
    calls = []

    class Triangulation:
        def __init__(self, x, y, triangles):
            self.x, self.y, self.triangles = np.asarray(x), np.asarray(y), np.asarray(triangles)
            assert self.triangles.min() >= 0 and self.triangles.max() < self.x.size
            calls.append(("Triangulation", self.triangles.shape[0]))

        def set_mask(self, mask):
            calls.append(("set_mask", int(np.count_nonzero(mask))))

    def rec(name):
        def f(*a, **k):
            calls.append((name,) + tuple(type(x).__name__ for x in a))
            return types.SimpleNamespace(show=lambda: None)
        return f

    plt = types.ModuleType("matplotlib.pyplot")
    for n in ("figure", "tricontourf", "colorbar", "title", "axis", "ion"):
        setattr(plt, n, rec(n))
    tri = types.ModuleType("matplotlib.tri")
    tri.Triangulation = Triangulation
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot, mpl.tri = plt, tri
    monkeypatch.setitem(sys.modules, "matplotlib", mpl)
    monkeypatch.setitem(sys.modules, "matplotlib.pyplot", plt)
    monkeypatch.setitem(sys.modules, "matplotlib.tri", tri)
    return calls


def test_mesh2tri_and_plot_calls(monkeypatch):
    """Triangles of the quad4 mesh ((n0,n1,n2), (n2,n3,n0) per element) and the figure sequence of the reference's
    plot branch (optimize.py:101-109): 2 x (2 displacement + 3 strain + 3 stress) contour figures + the design."""
    from solidspy_opt import plots
    from solidspy_opt.Utils.beams import beam
    nodes, mats, els, loads, BC = beam(L=6.0, H=4.0, nx=6, ny=4, dirs=np.array([[0, -1]]),
                                       positions=np.array([[1, 1]]), n=1)
    x, y, tri = plots.mesh2tri(nodes, els)
    assert tri.shape == (48, 3) and np.array_equal(tri[0], els[0, [3, 4, 5]]) and np.array_equal(tri[1], els[0, [5, 6, 3]])
    # every triangle is counter-clockwise with half the element's area
    area = 0.5*((x[tri[:, 1]] - x[tri[:, 0]])*(y[tri[:, 2]] - y[tri[:, 0]]) -
                (x[tri[:, 2]] - x[tri[:, 0]])*(y[tri[:, 1]] - y[tri[:, 0]]))
    assert np.allclose(area, 0.5)
    calls = fake_matplotlib(monkeypatch)
    nn = nodes.shape[0]
    rng = np.random.default_rng(0)
    fields = {k: rng.normal(size=(nn, 2 if k.startswith("UC") else 3))
              for k in ("UCI", "E_nodesI", "S_nodesI", "UC", "E_nodes", "S_nodes")}
    fields["E_nodes"][nn - 1] = np.nan                             # a node without a contributing element
    ELS = els[1:]
    plots.eso_plots(els, ELS, nodes, fields)
    assert sum(c[0] == "tricontourf" for c in calls) == 17
    assert [c[1] for c in calls if c[0] == "Triangulation"] == [48]*3 + [46]*4
    masks = [c[1] for c in calls if c[0] == "set_mask"]
    assert len(masks) == 16 and sum(m > 0 for m in masks) == 3      # the three strain figures of the final design
