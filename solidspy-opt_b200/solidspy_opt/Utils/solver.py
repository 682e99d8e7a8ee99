"""Helper functions with the names and signatures of the reference ``Utils/solver.py``.

Two kinds:
  * set-up / bookkeeping helpers (``protect_els``, ``del_node``, ``volume``, ``adjacency_nodes``,
    ``center_els``, ``strain_els`` ...) are vectorised numpy on the host -- they are not on the
    per-iteration device path and must reproduce the reference's integer results bit for bit;
  * numerical stages (``sparse_assem``, ``optimality_criteria``, ``density_filter``,
    ``sensitivity_elsESO/BESO``, ``sensitivity_nodes``, ``sensitivity_filter``) run on the GPU through
    the C ABI (``_capi``).  There is no CPU fallback: without the CUDA library they raise.
Structured-grid requirement: element/node arrays must use the ``rect_grid`` numbering (what ``beam``
produces); anything else raises ``ValueError``.
"""
import numpy as np

from solidspy_opt import _capi


# ------------------------------------------------------------------------------------------------
# structured-grid introspection
# ------------------------------------------------------------------------------------------------
def grid_shape_from_els(els):
    """(nx, ny) of a FULL rect_grid connectivity array; raises if the numbering is not rect_grid's."""
    els = np.asarray(els)
    ne = els.shape[0]
    nx = int(els[0, 6] - els[0, 3] - 1)
    if nx < 1 or ne % nx:
        raise ValueError("elements are not a structured rect_grid mesh")
    ny = ne//nx
    e = np.arange(ne, dtype=np.int64)
    n = e + e//nx
    ok = (np.array_equal(els[:, 3], n) and np.array_equal(els[:, 4], n + 1)
          and np.array_equal(els[:, 5], n + nx + 2) and np.array_equal(els[:, 6], n + nx + 1))
    if not ok:
        raise ValueError("elements are not numbered like solidspy rect_grid")
    return nx, ny


def grid_shape_from_nodes(nodes):
    """(nx, ny, dx, dy) from a full node array [id, x, y, ...] in rect_grid order."""
    nodes = np.asarray(nodes)
    x, y = nodes[:, 1], nodes[:, 2]
    npr = int(np.count_nonzero(y == y[0]))
    nn = nodes.shape[0]
    if npr < 2 or nn % npr:
        raise ValueError("nodes are not a structured rect_grid lattice")
    nx, ny = npr - 1, nn//npr - 1
    return nx, ny, float(x[1] - x[0]), float(y[npr] - y[0]) if ny > 0 else 0.0


# ------------------------------------------------------------------------------------------------
# bookkeeping helpers (host)
# ------------------------------------------------------------------------------------------------
def _protected(loads, BC):
    return np.hstack((np.asarray(loads)[:, 0], np.asarray(BC))).astype(np.int64)


def protect_els(els, nels, loads, BC):
    """Mask (by element id) of the elements of ``els`` touching a load or support node (ref. :6-33)."""
    els = np.asarray(els)
    hit = np.isin(els[:, -4:], _protected(loads, BC)).any(axis=1)
    mask_els = np.zeros(nels, dtype=bool)
    mask_els[els[hit, 0]] = True
    return mask_els


def protect_elsESO(els, loads, BC):
    """True where the element may be deleted, i.e. touches no load/support node (ref. :318-343)."""
    els = np.asarray(els)
    return ~np.isin(els[:, -4:], _protected(loads, BC)).any(axis=1)


def _used_nodes(nn, els):
    used = np.zeros(nn, dtype=bool)
    used[np.asarray(els)[:, -4:].ravel()] = True
    return used


def del_node(nodes, els, loads, BC):
    """Pin nodes that belong to no element of ``els``; free used, unprotected ones (ref. :35-58)."""
    used = _used_nodes(nodes.shape[0], els)
    prot = np.zeros(nodes.shape[0], dtype=bool)
    prot[_protected(loads, BC)] = True
    nodes[~used, -2:] = -1
    nodes[used & ~prot, -2:] = 0


def del_nodeESO(nodes, els):
    """Pin nodes that belong to no element of ``els`` (ref. :345-362)."""
    nodes[~_used_nodes(nodes.shape[0], els), -2:] = -1


def volume(els, length, height, nx, ny):
    """Per-element volume array (ref. :61-88; the dx/dy names are swapped there, the product is not)."""
    dy = length/nx
    dx = height/ny
    return dx*dy*np.ones(np.asarray(els).shape[0])


def adjacency_nodes(nodes, els):
    """List (one entry per node) of the ascending element ROW indices adjacent to it (ref. :131-151)."""
    els = np.asarray(els)
    nn = np.asarray(nodes).shape[0]
    con = els[:, -4:]
    rows = np.repeat(np.arange(els.shape[0]), 4)
    order = np.lexsort((rows, con.ravel()))          # by node, then by element row
    srt_nodes = con.ravel()[order]
    srt_rows = rows[order]
    bounds = np.searchsorted(srt_nodes, np.arange(nn + 1))
    return [srt_rows[bounds[i]:bounds[i + 1]] for i in range(nn)]


def center_els(nodes, els):
    """Element centres indexed by element id (the second, winning definition: ref. :479-501)."""
    nodes = np.asarray(nodes)
    els = np.asarray(els)
    n = nodes[els[:, -4:], 1:3]
    centers = np.zeros((els.shape[0], 2))
    centers[els[:, 0].astype(np.int64), 0] = n[:, 1, 0] + (n[:, 0, 0] - n[:, 1, 0])/2
    centers[els[:, 0].astype(np.int64), 1] = n[:, 2, 1] + (n[:, 0, 1] - n[:, 2, 1])/2
    return centers


def strain_els(els, E_nodes, S_nodes):
    """Element strains/stresses as the mean of the 4 nodal values, connectivity order (ref. :281-316)."""
    con = np.asarray(els)[:, 3:7]
    En, Sn = np.asarray(E_nodes)[con], np.asarray(S_nodes)[con]
    E_els = (En[:, 0] + En[:, 1] + En[:, 2] + En[:, 3])/4
    S_els = (Sn[:, 0] + Sn[:, 1] + Sn[:, 2] + Sn[:, 3])/4
    return E_els, S_els


# ------------------------------------------------------------------------------------------------
# element matrix helpers (host, tiny)
# ------------------------------------------------------------------------------------------------
def kloc_simp(E, nu):
    """Closed-form quad4 matrix of ``SIMP`` for square elements (reference optimize.py:408-416)."""
    k = np.array([1/2 - nu/6, 1/8 + nu/8, -1/4 - nu/12, -1/8 + 3*nu/8,
                  -1/4 + nu/12, -1/8 - nu/8, nu/6, 1/8 - 3*nu/8])
    pattern = np.array([[0, 1, 2, 3, 4, 5, 6, 7], [1, 0, 7, 6, 5, 4, 3, 2], [2, 7, 0, 5, 6, 3, 4, 1],
                        [3, 6, 5, 0, 7, 2, 1, 4], [4, 5, 6, 7, 0, 1, 2, 3], [5, 4, 3, 2, 1, 0, 7, 6],
                        [6, 3, 4, 1, 2, 7, 0, 5], [7, 2, 1, 4, 3, 6, 5, 0]])
    return E/(1 - nu**2)*k[pattern]


def ke_quad4(coord, E, nu):
    """Plane-stress bilinear quad stiffness, 2x2 Gauss, unit thickness (the ``elast_quad4`` contract)."""
    coord = np.asarray(coord, dtype=float)
    Cm = E/(1 - nu**2)*np.array([[1, nu, 0], [nu, 1, 0], [0, 0, (1 - nu)/2]])
    ri = np.array([-1.0, 1.0, 1.0, -1.0])
    si = np.array([-1.0, -1.0, 1.0, 1.0])
    gp = 1/np.sqrt(3.0)
    ke = np.zeros((8, 8))
    for r in (-gp, gp):
        for s in (-gp, gp):
            dN = 0.25*np.array([ri*(1 + s*si), si*(1 + r*ri)])
            J = dN @ coord
            det = J[0, 0]*J[1, 1] - J[0, 1]*J[1, 0]
            dNx = np.array([[J[1, 1], -J[0, 1]], [-J[1, 0], J[0, 0]]])/det @ dN
            B = np.zeros((3, 8))
            B[0, 0::2] = dNx[0]
            B[1, 1::2] = dNx[1]
            B[2, 0::2] = dNx[1]
            B[2, 1::2] = dNx[0]
            ke += det*(B.T @ Cm @ B)
    return ke


def beso_weights(nodes, els, nx, ny, r_min):
    """Floating-point weights of ``sensitivity_nodes`` / ``sensitivity_filter`` evaluated with the
    reference's own expressions (ref. :203, :237) on representative nodes / element 0."""
    nodes = np.asarray(nodes)
    els = np.asarray(els)

    def centre(e):
        n = nodes[els[e, -4:], 1:3]
        return np.array([n[1, 0] + (n[0, 0] - n[1, 0])/2, n[2, 1] + (n[0, 1] - n[2, 1])/2])

    def wn(node, adj):
        c = np.array([centre(e) for e in adj])
        r = np.linalg.norm(c - nodes[node, 1:3], axis=1)
        return 1/(len(adj) - 1)*(1 - r/r.sum())
    w4 = wn(nx + 2, [0, 1, nx, nx + 1]) if (nx > 1 and ny > 1) else np.full(4, 0.25)
    w2x = wn(1, [0, 1]) if nx > 1 else np.full(2, 0.5)
    w2y = wn(nx + 1, [0, nx]) if ny > 1 else np.full(2, 0.5)
    c0 = centre(0)
    # nodes within r_min of element 0's centre can only lie in the first few rows/columns
    rows = min(ny + 1, int(np.ceil(r_min/max(abs(nodes[min(nx + 1, nodes.shape[0] - 1), 2] - nodes[0, 2]), 1e-300))) + 2)
    cand = np.arange(min(nodes.shape[0], rows*(nx + 1)))
    r = np.linalg.norm(nodes[cand, 1:3] - c0, axis=1)
    om = r < r_min
    if not np.array_equal(cand[om], np.sort(els[0, -4:])):
        return w4, w2x, w2y, None                 # a wider (or narrower) radius: no 4-corner weights, see sensitivity_filter
    we = 1/(om.sum() - 1)*(1 - r[om]/r[om].sum())
    return w4, w2x, w2y, we


# ------------------------------------------------------------------------------------------------
# GPU-backed numerical stages
# ------------------------------------------------------------------------------------------------
class StiffnessOperator:
    """What ``sparse_assem`` returns here: K(scale) as a matrix-free device operator.

    Supports the two things the reference does with ``stiff_mat`` besides handing it to the solver
    (optimize.py:455): ``.dot(disp)`` on the reduced (free-DOF) vector and ``.max()``.
    """

    def __init__(self, topo, bc_free_index, neq):
        self.topo, self._free, self.shape = topo, bc_free_index, (neq, neq)

    def _expand(self, v):
        full = np.zeros(2*self.topo.nn)
        full[self._free] = np.asarray(v, dtype=float)
        return full

    def dot(self, v):
        return self.topo.matvec(self._expand(v)).reshape(-1)[self._free]

    def max(self):
        scale = self.topo.get(_capi.F_SCALE).reshape(self.topo.ny, self.topo.nx)
        pad = np.zeros((self.topo.ny + 2, self.topo.nx + 2))
        pad[1:-1, 1:-1] = scale
        ke = self._ke
        dx = pad[1:, 1:]*ke[0, 0] + pad[1:, :-1]*ke[2, 2] + pad[:-1, :-1]*ke[4, 4] + pad[:-1, 1:]*ke[6, 6]
        dy = pad[1:, 1:]*ke[1, 1] + pad[1:, :-1]*ke[3, 3] + pad[:-1, :-1]*ke[5, 5] + pad[:-1, 1:]*ke[7, 7]
        d = np.stack([dx, dy], -1).reshape(-1)
        return float(d[self._free].max())


def sparse_assem(elements, mats, neq, assem_op, kloc):
    """Reference signature (:364-410).  Returns a matrix-free device operator instead of a CSR matrix:
    K = sum_e mats[e, 2] * kloc is applied on the fly by the stage-1 kernel."""
    elements = np.asarray(elements)
    nx, ny = grid_shape_from_els(elements)
    topo = _capi.Topo(nx, ny, np.asarray(kloc, dtype=float))
    nn = topo.nn
    # recover the BC flags from the assembly operator: -1 entries are constrained DOFs
    cons = np.zeros((nn, 2), dtype=np.int8)
    eq = np.full(2*nn, -1, dtype=np.int64)
    dof = np.stack([2*elements[:, 3:7], 2*elements[:, 3:7] + 1], axis=2).reshape(elements.shape[0], 8)
    eq[dof.ravel()] = np.asarray(assem_op)[:, :8].ravel()
    cons.reshape(-1)[eq == -1] = -1
    topo.set_cons(cons)
    topo.set(_capi.F_SCALE, np.asarray(mats)[elements[:, 0], 2])
    free = np.flatnonzero(eq != -1)
    order = np.argsort(eq[free], kind="stable")
    op = StiffnessOperator(topo, free[order], neq)
    op._ke = np.asarray(kloc, dtype=float)
    return op


_scratch = {}


def _scratch_topo(nx, ny):
    key = (int(nx), int(ny))
    if key not in _scratch:
        _scratch.clear()
        _scratch[key] = _capi.Topo(nx, ny, np.eye(8))
    return _scratch[key]


def optimality_criteria(nelx, nely, rho, d_c, g):
    """Reference signature (:412-448): OC update with the multiplier bisection done on the device."""
    t = _scratch_topo(nelx, nely)
    t.set(_capi.F_RHO, rho)
    t.set(_capi.F_DC, d_c)
    gt, _, _ = t.oc_update(float(g))
    return t.get(_capi.F_RHO), gt


def density_filter(centers, r_min, rho, d_rho):
    """Reference signature (:452-477): the (sensitivity) filter as a device stencil.  ``centers`` must be
    the centres of a structured grid in element-id order."""
    centers = np.asarray(centers)
    ne = centers.shape[0]
    nx = int(np.count_nonzero(centers[:, 1] == centers[0, 1]))
    if nx < 1 or ne % nx:
        raise ValueError("centers do not form a structured grid")
    ny = ne//nx
    dx = float(centers[1, 0] - centers[0, 0]) if nx > 1 else float(r_min)
    dy = float(centers[nx, 1] - centers[0, 1]) if ny > 1 else float(r_min)
    t = _scratch_topo(nx, ny)
    t.set(_capi.F_RHO, rho)
    t.set(_capi.F_DC, d_rho)
    t.filter_sensitivity(float(r_min), dx, dy)
    return t.get(_capi.F_DC)


def _energy(nodes, mats, els, mask, UC):
    nodes = np.asarray(nodes)
    els = np.asarray(els)
    nx, ny, _, _ = grid_shape_from_nodes(nodes)
    e0 = int(els[0, 0])
    # element matrix from the first element handed in (all elements share it on a uniform grid)
    ke = ke_quad4(nodes[els[0, -4:], 1:3], *np.asarray(mats)[els[0, 2], :2])
    t = _capi.Topo(nx, ny, ke)
    keep = np.zeros(nx*ny, dtype=np.uint8)
    ids = els[:, 0].astype(np.int64)
    keep[ids] = 1 if mask is None else np.asarray(mask, dtype=np.uint8)
    t.set(_capi.F_KEEP, keep)
    t.set(_capi.F_U, UC)
    t.energy_sensitivity()
    out = t.get(_capi.F_SENS)[ids]
    t.close()
    del e0
    return out


def sensitivity_elsBESO(nodes, mats, els, mask, UC):
    """Reference signature (:90-129): 1/2 u_e^T K_e u_e (0 where ``mask`` is False) divided by its max."""
    return _energy(nodes, mats, els, mask, UC)


def sensitivity_elsESO(nodes, mats, els, UC):
    """Reference signature (:245-279)."""
    return _energy(nodes, mats, els, None, UC)


def _unit_els(nx, ny):
    e = np.arange(nx*ny, dtype=np.int64)
    n = e + e//nx
    return np.stack([e, e*0 + 1, e*0, n, n + 1, n + nx + 2, n + nx + 1], axis=1)


def sensitivity_nodes(nodes, adj_nodes, centers, sensi_els):
    """Reference signature (:177-210): nodal value = weighted mean of the adjacent elements' numbers
    (device kernel; weights evaluated with the reference's expression)."""
    nodes = np.asarray(nodes)
    nx, ny, _, _ = grid_shape_from_nodes(nodes)
    w4, w2x, w2y, _ = beso_weights(nodes, _unit_els(nx, ny), nx, ny, np.linalg.norm(nodes[0, 1:3] - nodes[1, 1:3]))
    t = _scratch_topo(nx, ny)
    t.set(_capi.F_SENS, sensi_els)
    t.beso_nodes(w4, w2x, w2y)
    return t.get(_capi.F_NODAL)


def sensitivity_filter(nodes, centers, sensi_nodes, r_min):
    """Reference signature (:212-243), any ``r_min`` (device kernels).  The radius the optimisers use (one element
    width, optimize.py:264: exactly the 4 corner nodes) runs the kernel whose operation order reproduces the reference
    bit for bit; other radii run the general stencil (nodes within ``r_min`` of the element centre, cut off at the mesh
    boundary; 1e-13 relative to the reference's summation order)."""
    nodes = np.asarray(nodes)
    nx, ny, dx, dy = grid_shape_from_nodes(nodes)
    _, _, _, we = beso_weights(nodes, _unit_els(nx, ny), nx, ny, r_min)
    t = _scratch_topo(nx, ny)
    t.set(_capi.F_NODAL, sensi_nodes)
    if we is not None:
        t.beso_elements(we)
    else:
        t.beso_elements_radius(r_min, dx, dy)
    return t.get(_capi.F_SENS)
