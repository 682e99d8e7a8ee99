"""``plot=True`` outputs of the optimisers (reference optimize.py:101-109, 200-208, 346-354, 458-463).

The reference draws with ``solidspy.postprocesor.fields_plot`` / ``mesh2tri`` and matplotlib; both are optional
here and imported lazily, so the accelerated path never depends on them.  The FIELDS that are drawn (complete
displacements, nodal strains and stresses before and after the optimisation) come from the GPU
(``Topo.strain_nodes``); this module only triangulates and calls matplotlib.
"""
import numpy as np


def mesh2tri(nodes, elements):
    """Triangle list of a quad4 mesh as ``pos.mesh2tri`` builds it: every quadrilateral ``[n0, n1, n2, n3]``
    contributes ``(n0, n1, n2)`` and ``(n2, n3, n0)``.  Returns ``(x, y, triangles)``; the reference wraps them in a
    ``matplotlib.tri.Triangulation``."""
    con = np.asarray(elements)[:, 3:7].astype(np.int64)
    tri = np.empty((2*con.shape[0], 3), dtype=np.int64)
    tri[0::2] = con[:, [0, 1, 2]]
    tri[1::2] = con[:, [2, 3, 0]]
    return nodes[:, 1], nodes[:, 2], tri


def _triangulation(nodes, elements):
    from matplotlib.tri import Triangulation              # optional dependency, only for plot=True
    x, y, tri = mesh2tri(nodes, elements)
    return Triangulation(x, y, tri)


def plot_node_field(field, nodes, elements, titles, figtitle):
    """One filled-contour figure per column of ``field`` on the triangulated mesh (``pos.plot_node_field``)."""
    import matplotlib.pyplot as plt
    tri = _triangulation(nodes, elements)
    field = np.asarray(field, dtype=float)
    if field.ndim == 1:
        field = field[:, None]
    figs = []
    for k in range(field.shape[1]):
        fig = plt.figure(figtitle[k] if k < len(figtitle) else None)
        # orphan nodes of an ESO design carry NaN (0/0 in strain_nodes); matplotlib masks triangles that touch them
        vals = field[:, k]
        bad = ~np.isfinite(vals)
        tri.set_mask(bad[tri.triangles].any(axis=1) if bad.any() else None)
        vals = np.where(bad, 0.0, vals)
        cs = plt.tricontourf(tri, vals, 12)
        plt.colorbar(cs)
        plt.title(titles[k] if k < len(titles) else "")
        plt.axis("image")
        figs.append(fig)
    return figs


def fields_plot(elements, nodes, disp, E_nodes=None, S_nodes=None):
    """Displacement, strain and stress contour figures (``pos.fields_plot``)."""
    figs = plot_node_field(disp, nodes, elements, [r"$u_x$", r"$u_y$"],
                           ["Horizontal displacement", "Vertical displacement"])
    if E_nodes is not None:
        figs += plot_node_field(E_nodes, nodes, elements,
                                [r"$\epsilon_{xx}$", r"$\epsilon_{yy}$", r"$\gamma_{xy}$"],
                                ["Strain epsilon-xx", "Strain epsilon-yy", "Strain gamma-xy"])
    if S_nodes is not None:
        figs += plot_node_field(S_nodes, nodes, elements,
                                [r"$\sigma_{xx}$", r"$\sigma_{yy}$", r"$\tau_{xy}$"],
                                ["Stress sigma-xx", "Stress sigma-yy", "Stress tau-xy"])
    return figs


def design_plot(nodes, ELS, n_nodes):
    """The black/white picture of the remaining elements (optimize.py:105-109)."""
    import matplotlib.pyplot as plt
    fill_plot = np.ones(n_nodes)
    plt.figure()
    plt.tricontourf(_triangulation(nodes, ELS), fill_plot, cmap='binary')
    plt.axis("image")


def eso_plots(elsI, ELS, nodes, fields):
    """The three plot statements shared by ESO_stress, ESO_stiff and BESO (:101-109, :200-208, :346-354)."""
    fields_plot(elsI, nodes, fields["UCI"], E_nodes=fields["E_nodesI"], S_nodes=fields["S_nodesI"])   # initial mesh
    fields_plot(ELS, nodes, fields["UC"], E_nodes=fields["E_nodes"], S_nodes=fields["S_nodes"])       # optimised mesh
    design_plot(nodes, ELS, fields["E_nodes"].shape[0])
