"""``solidspy.preprocesor`` subset: structured grid generation (called at reference beams.py:73,126)."""
import numpy as np


def rect_grid(length, height, nx, ny, eletype=None):
    """Nodes of a ``(nx+1) x (ny+1)`` lattice and the quad4 connectivity.

    Row-major, x fastest, y from -height/2 upwards.  ``els[e] = [e, 1, 0, n, n+1, n+nx+2, n+nx+1]``
    with ``e = row*nx + col`` and ``n = e + row`` (counter-clockwise from the lower-left node).
    Coordinates come from ``numpy.mgrid`` with a complex step (``start + i*(stop-start)/(n-1)``).
    """
    y, x = np.mgrid[-height/2:height/2:(ny + 1)*1j, -length/2:length/2:(nx + 1)*1j]
    e = np.arange(nx*ny)
    n = e + e//nx
    els = np.zeros((nx*ny, 7), dtype=int)
    els[:, 0] = e
    els[:, 1] = 1
    els[:, 3] = n
    els[:, 4] = n + 1
    els[:, 5] = n + nx + 2
    els[:, 6] = n + nx + 1
    return x.flatten(), y.flatten(), els
