"""The oracle's ``solidspy`` stand-in against the known answers the reference tree holds."""
import os
import sys

import numpy as np
from scipy.sparse.linalg import spsolve

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "oracle", "compat"))
import solidspy.assemutil as ass          # noqa: E402
import solidspy.postprocesor as pos       # noqa: E402
import solidspy.preprocesor as pre        # noqa: E402
import solidspy.uelutil as uel            # noqa: E402

# inputs of reference examples/square-4_elements/{nodes,eles,mater,loads}.txt (4 elements, 9 nodes)
NODES = np.array([[0, 0, 0, 0, -1], [1, 2, 0, 0, -1], [2, 2, 2, 0, 0], [3, 0, 2, 0, 0], [4, 1, 0, -1, -1],
                  [5, 2, 1, 0, 0], [6, 1, 2, 0, 0], [7, 0, 1, 0, 0], [8, 1, 1, 0, 0]], dtype=float)
ELS = np.array([[0, 1, 0, 0, 4, 8, 7], [1, 1, 0, 4, 1, 5, 8], [2, 1, 0, 7, 8, 6, 3], [3, 1, 0, 8, 5, 2, 6]])
MATS = np.array([[1.0, 0.3]])
LOADS = np.array([[3, 0.0, 1.0], [6, 0.0, 2.0], [2, 0.0, 1.0]])
# reference docs/tutorials/square_example.rst:187 (rounded to 3 decimals there) and :150 (14 equations)
EXPECTED = np.array([0.6, -0.6, -0.6, 4., 0.6, 4., -0.6, 2., -0., 4., 0.6, 2., -0., 2.])


def test_square_example_known_answer():
    assem_op, bc, neq = ass.DME(NODES[:, -2:], ELS)
    assert neq == 14
    K, _ = ass.assembler(ELS, MATS, NODES[:, :3], neq, assem_op)
    f = ass.loadasem(LOADS, bc, neq)
    u = spsolve(K, f)
    assert np.array_equal(np.round(u, 3) + 0.0, EXPECTED + 0.0)
    UC = pos.complete_disp(bc, NODES, u)
    assert UC.shape == (9, 2) and UC[4, 0] == 0 and UC[4, 1] == 0 and UC[0, 1] == 0
    mag = np.sqrt(u[0::2]**2 + u[1::2]**2)
    # square_example.rst:197
    assert np.array_equal(np.round(mag, 3), np.array([0.849, 4.045, 4.045, 2.088, 4., 2.088, 2.]))


def test_rect_grid_numbering():
    x, y, els = pre.rect_grid(60, 60, 60, 60)
    assert els.shape == (3600, 7) and els.dtype.kind == "i"
    assert list(els[0]) == [0, 1, 0, 0, 1, 62, 61]
    assert list(els[60]) == [60, 1, 0, 61, 62, 123, 122]
    assert x[0] == -30 and x[60] == 30 and y[0] == -30 and y[61] == -29


def test_closed_form_ke_equals_gauss_on_squares():
    import t1
    for h in (1.0, 0.37, 12.5):
        coord = np.array([[0, 0], [h, 0], [h, h], [0, h]], dtype=float)
        k, _ = uel.elast_quad4(coord, (206.8e9, 0.28, 1.0))
        kc = t1.kloc_closed_form(206.8e9, 0.28)
        assert np.abs(k - kc).max() <= 1e-15*np.abs(kc).max()*8
        assert np.abs(k - t1.ke_gauss(coord, 206.8e9, 0.28)).max() <= 1e-15*np.abs(kc).max()*8


def test_strain_nodes_constant_strain_patch():
    # u = (a x, b y): strains are constant, so nodal averaging must return them exactly
    x, y, els = pre.rect_grid(4, 2, 4, 2)
    nodes = np.zeros((x.size, 5))
    nodes[:, 0] = np.arange(x.size); nodes[:, 1] = x; nodes[:, 2] = y
    UC = np.stack([0.01*x, -0.02*y], 1)
    E_n, S_n = pos.strain_nodes(nodes, els, np.array([[10.0, 0.25]]), UC)
    assert np.allclose(E_n, [0.01, -0.02, 0.0], atol=1e-15)
    c = 10/(1 - 0.25**2)
    assert np.allclose(S_n[:, 0], c*(0.01 - 0.25*0.02)) and np.allclose(S_n[:, 2], 0, atol=1e-14)


def test_beam_convergence_error_table():
    """reference examples/beam_convergence: the notebook's mesh-refinement loop (rect_grid, DME, assembler, loadasem,
    solve, complete_disp, scipy griddata to the next mesh) run on the stand-in reproduces error_vs_h.txt: iterations
    2-3 to all six printed digits, 4-6 to 4e-4 (there the difference is how scipy's griddata breaks the ties of the
    degenerate Delaunay triangulation of a regular grid, not the finite-element arithmetic)."""
    from scipy.interpolate import griddata
    published = {2: 0.0489864, 3: 0.107604, 4: 0.0427641, 5: 0.0277539, 6: 0.01599}     # error_vs_h.txt rows 2-6
    P, E, nu, L, h = -50, 1000, 0.3, 24, 8
    mats = np.array([[E, nu]])
    u_i = v_i = None
    for cont in range(1, 7):
        nx, ny = 3*2**(cont - 1), 2**(cont - 1)
        x, y, els = pre.rect_grid(L, h, nx, ny)
        nodes = np.zeros(((nx + 1)*(ny + 1), 5))
        nodes[:, 0] = range((nx + 1)*(ny + 1))
        nodes[:, 1] = x
        nodes[:, 2] = y
        nodes[x == L/2, 3] = -1
        nodes[nx*(ny//2 + 1) - 1, 4] = -1
        loads = np.zeros((ny + 1, 3))
        loads[:, 0] = nodes[x == -L/2, 0]
        loads[:, 2] = P/ny
        assem_op, bc, neq = ass.DME(nodes[:, -2:], els)
        K, _ = ass.assembler(els, mats, nodes[:, :3], neq, assem_op)
        u = spsolve(K, ass.loadasem(loads, bc, neq))
        UC = pos.complete_disp(bc, nodes, u)
        if cont > 1:
            err = np.linalg.norm(np.column_stack([u_i, v_i]) - UC)/np.linalg.norm(UC)
            if cont <= 3:
                assert "%g" % err == "%g" % published[cont], (cont, err)
            else:
                assert abs(err - published[cont]) <= 1e-3*published[cont], (cont, err)
        xn, yn, _ = pre.rect_grid(L, h, 2*nx, 2*ny)
        u_i = griddata((x, y), UC[:, 0], (xn, yn))
        v_i = griddata((x, y), UC[:, 1], (xn, yn))
