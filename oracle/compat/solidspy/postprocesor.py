"""``solidspy.postprocesor`` subset (call sites: reference ``optimize.py:62-63,85-86,...,438``)."""
import numpy as np
from . import femutil as fem


def complete_disp(bc_array, nodes, sol, ndof_node=2):
    """Scatter the reduced solution into (nnodes, 2); constrained DOFs are zero."""
    out = np.zeros(bc_array.shape, dtype=float)
    free = bc_array != -1
    out[free] = np.asarray(sol)[bc_array[free]]
    return out


def strain_nodes(nodes, elements, mats, sol_complete):
    """Nodal strains/stresses: corner-point strains of every element averaged per node.

    Nodes that belong to no element divide 0/0 (NaN) -- the reference silences that warning
    (``optimize.py:12``).
    """
    nnodes = nodes.shape[0]
    E_nodes = np.zeros((nnodes, 3))
    S_nodes = np.zeros((nnodes, 3))
    count = np.zeros(nnodes, dtype=int)
    con = elements[:, 3:]
    for el in range(elements.shape[0]):
        young, poisson = mats[int(elements[el, 2]), :]
        shear = young/(2*(1 + poisson))
        f1 = young/(1 - poisson**2)
        f2 = poisson*young/(1 - poisson**2)
        coord = nodes[con[el], 1:3]
        ul = sol_complete[con[el]].reshape(-1)
        eps, _ = fem.str_el4(coord, ul)
        for c, node in enumerate(con[el]):
            E_nodes[node] += eps[c]
            S_nodes[node, 0] += f1*eps[c, 0] + f2*eps[c, 1]
            S_nodes[node, 1] += f2*eps[c, 0] + f1*eps[c, 1]
            S_nodes[node, 2] += shear*eps[c, 2]
            count[node] += 1
    with np.errstate(divide="ignore", invalid="ignore"):
        E_nodes = E_nodes/count[:, None]
        S_nodes = S_nodes/count[:, None]
    return E_nodes, S_nodes


def fields_plot(*args, **kwargs):      # plotting is out of scope
    raise NotImplementedError("plotting is not part of the oracle")


def mesh2tri(*args, **kwargs):
    raise NotImplementedError("plotting is not part of the oracle")
