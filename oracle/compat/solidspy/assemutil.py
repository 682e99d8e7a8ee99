"""``solidspy.assemutil`` subset: equation numbering, assembly, load vector.

Call sites in the reference: ``optimize.py:56-58,79-81,156-158,176-178,255-257,293-295,417,434``.
"""
import numpy as np
from scipy.sparse import coo_matrix
from . import femutil as fem
from . import uelutil as uel


def eqcounter(cons, ndof_node=2):
    """Number the free DOFs consecutively (node-major, dof-minor); constrained DOFs stay -1."""
    bc_array = cons.astype(int).copy()
    free = bc_array == 0
    neq = int(free.sum())
    bc_array[free] = np.arange(neq)          # boolean indexing walks C order = node-major, dof-minor
    return neq, bc_array


def DME(cons, elements, ndof_node=2, ndof_el_max=18, ndof_el=None):
    """Assembly operator ``assem_op[e, :ndof] = bc_array[nodes_of(e)].flatten()``."""
    nels = elements.shape[0]
    assem_op = np.zeros([nels, ndof_el_max], dtype=int)
    neq, bc_array = eqcounter(cons, ndof_node=ndof_node)
    for ele in range(nels):
        if ndof_el is None:
            ndof, _, _ = fem.eletype(elements[ele, 1])
        else:
            ndof = ndof_el(elements[ele, 1])
        assem_op[ele, :ndof] = bc_array[elements[ele, 3:]].flatten()
    return assem_op, bc_array, neq


def assembler(elements, mats, nodes, neq, assem_op, sparse=True, uel=None):
    """Global stiffness and mass matrices; per-element ``elast_quad4`` scattered through ``assem_op``."""
    rows, cols, kv, mv = [], [], [], []
    cache = {}
    for ele in range(elements.shape[0]):
        params = mats[elements[ele, 2], :]
        coord = nodes[elements[ele, 3:], 1:3]
        # identical (geometry, material) pairs give identical matrices: reuse them (pure memoisation)
        key = (tuple((coord - coord[0]).ravel()), tuple(params))
        if key not in cache:
            cache[key] = _uel_quad4(coord, params) if uel is None else uel(coord, params)
        k, m = cache[key]
        dme = assem_op[ele, :8]
        ok = dme != -1
        rr, cc = np.meshgrid(dme[ok], dme[ok], indexing="ij")
        rows.append(rr.ravel()); cols.append(cc.ravel())
        kv.append(k[np.ix_(ok, ok)].ravel()); mv.append(m[np.ix_(ok, ok)].ravel())
    rows = np.concatenate(rows); cols = np.concatenate(cols)
    K = coo_matrix((np.concatenate(kv), (rows, cols)), shape=(neq, neq)).tocsr()
    M = coo_matrix((np.concatenate(mv), (rows, cols)), shape=(neq, neq)).tocsr()
    return K, M


def _uel_quad4(coord, params):
    return uel.elast_quad4(coord, params)


def loadasem(loads, bc_array, neq, ndof_node=2):
    """``rhs[eq(node, d)] = loads[i, d+1]`` for free DOFs (assignment, later rows overwrite)."""
    rhs = np.zeros([neq])
    for i in range(loads.shape[0]):
        node = int(loads[i, 0])
        for d in range(ndof_node):
            eq = bc_array[node, d]
            if eq != -1:
                rhs[eq] = loads[i, d + 1]
    return rhs
