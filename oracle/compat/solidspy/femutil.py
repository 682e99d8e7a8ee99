"""Element-level helpers for the bilinear quadrilateral (the only element type on the hot path)."""
import numpy as np

# natural coordinates of the 4 nodes, counter-clockwise from (-1, -1)
_RI = np.array([-1.0, 1.0, 1.0, -1.0])
_SI = np.array([-1.0, -1.0, 1.0, 1.0])


def eletype(iet):
    """(ndof, nnodes, ngpts) for an element-type id; only quad4 (1) is needed here."""
    table = {1: (8, 4, 4), 2: (12, 6, 7), 3: (6, 3, 3), 4: (18, 9, 9), 5: (4, 2, 3), 6: (4, 2, 3)}
    return table[int(iet)]


def umat(params):
    """Plane-stress constitutive matrix from (E, nu)."""
    E, nu = params[0], params[1]
    c = E/(1.0 - nu**2)
    return c*np.array([[1.0, nu, 0.0], [nu, 1.0, 0.0], [0.0, 0.0, (1.0 - nu)/2.0]])


def shape_quad4(r, s):
    """Shape functions N (4,) and natural derivatives dN (2,4) at (r, s)."""
    N = 0.25*(1 + r*_RI)*(1 + s*_SI)
    dN = 0.25*np.array([_RI*(1 + s*_SI), _SI*(1 + r*_RI)])
    return N, dN


def elast_diff_2d(r, s, coord, element=shape_quad4):
    """Interpolation matrix H (2,8), strain-displacement matrix B (3,8) and det(J) at (r, s)."""
    N, dNdr = element(r, s)
    jac = dNdr @ coord
    det = jac[0, 0]*jac[1, 1] - jac[0, 1]*jac[1, 0]
    inv = np.array([[jac[1, 1], -jac[0, 1]], [-jac[1, 0], jac[0, 0]]])/det
    dNdx = inv @ dNdr
    H = np.zeros((2, 8))
    H[0, 0::2] = N
    H[1, 1::2] = N
    B = np.zeros((3, 8))
    B[0, 0::2] = dNdx[0]
    B[1, 1::2] = dNdx[1]
    B[2, 0::2] = dNdx[1]
    B[2, 1::2] = dNdx[0]
    return H, B, det


def str_el4(coord, ul):
    """Strains (4,3) at the four corners (natural coords +-1) of a quad4, plus their positions."""
    eps = np.zeros((4, 3))
    xl = np.zeros((4, 2))
    for i in range(4):
        H, B, _ = elast_diff_2d(_RI[i], _SI[i], coord)
        eps[i] = B @ ul
        xl[i] = (H[0, 0::2] @ coord[:, 0], H[0, 0::2] @ coord[:, 1])
    return eps, xl
