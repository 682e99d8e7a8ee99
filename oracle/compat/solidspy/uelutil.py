"""``solidspy.uelutil`` subset (called at reference Utils/solver.py:119,269 and by ``assembler``)."""
import numpy as np
from . import femutil as fem
from . import gaussutil as gau


def elast_quad4(coord, params):
    """Stiffness and mass matrix (8x8 each) of a plane-stress bilinear quad, 2x2 Gauss, unit thickness.

    ``params = (E, nu[, density])``; DOFs interleaved ``[ux0, uy0, ..., ux3, uy3]``.
    """
    C = fem.umat(params[:2])
    dens = 1.0 if len(params) == 2 else params[-1]
    k = np.zeros((8, 8))
    m = np.zeros((8, 8))
    gpts, gwts = gau.gauss_nd(2)
    for (r, s), w in zip(gpts, gwts):
        H, B, det = fem.elast_diff_2d(r, s, coord, fem.shape_quad4)
        k += det*w*(B.T @ C @ B)
        m += dens*det*w*(H.T @ H)
    return k, m
