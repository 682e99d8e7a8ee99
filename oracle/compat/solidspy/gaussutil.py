"""Gauss-Legendre points on [-1,1]^d (2-point rule is what quad4 uses)."""
import itertools
import numpy as np


def gauss_1d(npts):
    return np.polynomial.legendre.leggauss(npts)


def gauss_nd(npts, ndim=2):
    pts, wts = gauss_1d(npts)
    gpts = np.array(list(itertools.product(pts, repeat=ndim)))
    gwts = np.array([np.prod(w) for w in itertools.product(wts, repeat=ndim)])
    return gpts, gwts
