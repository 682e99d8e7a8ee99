"""Structured quad4 beam models (host side, setup only -- not on the per-iteration path).

Produces the arrays of the reference ``Utils/beams.py`` (``beam`` :4, ``beam_1`` :39, ``beam_2`` :92)
bit for bit -- connectivity, node ids, DOF flags, load-node formula -- without Python loops, so that
meshes of 10^6..10^7 elements are generated in milliseconds.  ``rect_grid`` restates the numbering
contract of ``solidspy.preprocesor.rect_grid``.
"""
import numpy as np


def rect_grid(length, height, nx, ny):
    """Coordinates (x fastest, y upwards from -height/2) and quad4 connectivity of an nx-by-ny grid.

    Element ``e = row*nx + col`` has nodes ``[n, n+1, n+nx+2, n+nx+1]`` with ``n = e + row``.
    """
    # numpy.mgrid with a complex step evaluates start + i*(stop-start)/(n-1); reproduce it exactly
    sy = (height/2 - (-height/2))/float(ny) if ny + 1 != 1 else 1
    sx = (length/2 - (-length/2))/float(nx) if nx + 1 != 1 else 1
    yy = np.arange(ny + 1, dtype=float)*sy + (-height/2)
    xx = np.arange(nx + 1, dtype=float)*sx + (-length/2)
    x = np.tile(xx, ny + 1)
    y = np.repeat(yy, nx + 1)
    e = np.arange(nx*ny, dtype=np.int64)
    n = e + e//nx
    els = np.empty((nx*ny, 7), dtype=np.int64)
    els[:, 0] = e
    els[:, 1] = 1
    els[:, 2] = 0
    els[:, 3] = n
    els[:, 4] = n + 1
    els[:, 5] = n + nx + 2
    els[:, 6] = n + nx + 1
    return x, y, els


def _model(L, H, E, v, nx, ny, dirs, positions, support):
    dirs = np.asarray(dirs)
    positions = np.asarray(positions)
    x, y, els = rect_grid(L, H, nx, ny)
    mats = np.empty((els.shape[0], 3))
    mats[:] = [E, v, 1]
    nn = (nx + 1)*(ny + 1)
    nodes = np.zeros((nn, 5))
    nodes[:, 0] = np.arange(nn)
    nodes[:, 1] = x
    nodes[:, 2] = y
    mask = support(x, y)
    nodes[mask, 3:] = -1
    loads = np.zeros((dirs.shape[0], 3), dtype=int)         # integer loads, as in the reference (:83)
    if dirs.shape[0]:
        loads[:, 0] = nx*positions[:, 0] + (positions[:, 0] - positions[:, 1])
        loads[:, 1] = dirs[:, 0]
        loads[:, 2] = dirs[:, 1]
    BC = nodes[mask, 0]
    return nodes, mats, els, loads, BC


_SUPPORT = {
    1: lambda L, H: (lambda x, y: x == -L/2),                                                   # beams.py:80-81
    2: lambda L, H: (lambda x, y: ((x == L/2) & (y > H/4)) | ((x == L/2) & (y < -H/4))),          # beams.py:133-136
}


def beam_bc(L, H, nx, ny, dirs, positions, n=1):
    """What the SIMP loop needs of ``beam()`` -- BC flags ``nodes[:, 3:5]`` as int8, ``loads`` and the spacing
    ``|nodes[0] - nodes[1]|`` its filter radius is built from (optimize.py:404) -- from the two 1-D coordinate vectors,
    without the node / element / material arrays (252 MB and 0.4 s at 2048x1024; identical values, ``tests/test_host.py``)."""
    dirs = np.asarray(dirs)
    positions = np.asarray(positions)
    sy = (H/2 - (-H/2))/float(ny) if ny + 1 != 1 else 1
    sx = (L/2 - (-L/2))/float(nx) if nx + 1 != 1 else 1
    yy = np.arange(ny + 1, dtype=float)*sy + (-H/2)
    xx = np.arange(nx + 1, dtype=float)*sx + (-L/2)
    if n not in _SUPPORT:
        raise ValueError("beam selector n must be 1 or 2")
    mask = _SUPPORT[n](L, H)(xx[None, :], yy[:, None])
    mask = np.broadcast_to(mask, (ny + 1, nx + 1))
    cons = np.zeros((ny + 1, nx + 1, 2), dtype=np.int8)
    cons[mask] = -1
    loads = np.zeros((dirs.shape[0], 3), dtype=int)
    if dirs.shape[0]:
        loads[:, 0] = nx*positions[:, 0] + (positions[:, 0] - positions[:, 1])
        loads[:, 1] = dirs[:, 0]
        loads[:, 2] = dirs[:, 1]
    spacing = float(np.linalg.norm(np.array([xx[0], yy[0]]) - np.array([xx[1] if nx >= 1 else xx[0], yy[0]])))
    return cons.reshape(-1, 2), loads, spacing


def beam_1(L=10, H=10, E=206.8e9, v=0.28, nx=20, ny=20, dirs=np.array([]), positions=np.array([])):
    """Cantilever: every node on x == -L/2 fully clamped (reference beams.py:80-81)."""
    return _model(L, H, E, v, nx, ny, dirs, positions, lambda x, y: x == -L/2)


def beam_2(L=10, H=10, E=206.8e9, v=0.28, nx=20, ny=20, dirs=np.array([]), positions=np.array([])):
    """Right edge clamped above H/4 and below -H/4 (reference beams.py:133-136)."""
    return _model(L, H, E, v, nx, ny, dirs, positions,
                  lambda x, y: ((x == L/2) & (y > H/4)) | ((x == L/2) & (y < -H/4)))


def beam(L=10, H=10, E=206.8e9, v=0.28, nx=20, ny=20, dirs=np.array([]), positions=np.array([]), n=1):
    """Selector with the reference's signature (beams.py:4); returns nodes, mats, els, loads, BC."""
    if n == 1:
        return beam_1(L, H, E, v, nx, ny, dirs, positions)
    if n == 2:
        return beam_2(L, H, E, v, nx, ny, dirs, positions)
    return None   # the reference's ``match`` falls through for any other selector
