"""The four optimisers of SolidsPy-Opt with the reference's signatures, driven on a B200.

Entry points (argument order identical to reference ``src/solidspy_opt/optimize.py``):

    SIMP      (length, height, nx, ny, dirs, positions, niter, penal, plot=False)            :360
    BESO      (length, height, nx, ny, dirs, positions, niter, t, ER, volfrac, plot=False)   :215
    ESO_stiff (length, height, nx, ny, dirs, positions, niter, RR, ER, volfrac, plot=False)  :116
    ESO_stress(length, height, nx, ny, dirs, positions, niter, RR, ER, volfrac, plot=False)  :16

The host loop stays in Python; every numerical stage of an iteration is a call into
``libtopo_b200.so`` (matrix-free K(rho)u, GMG-preconditioned CG instead of ``spsolve``, sensitivity,
filter, OC / rank-select removal).  Keyword-only extras (``precond``, ``rtol``, ``device``,
``history`` ...) default to the reference's behaviour; ``*_from_model`` variants take the
``{"nodes","cons","elements","mats","loads"}`` dict of the SolidsPy README (reference README.rst:134-141).
"""
import time

import numpy as np

from solidspy_opt import _capi
from solidspy_opt.Utils.beams import beam, beam_bc
from solidspy_opt.Utils.solver import (beso_weights, grid_shape_from_els, grid_shape_from_nodes, ke_quad4,
                                       kloc_simp, volume)

np.seterr(divide='ignore', invalid='ignore')      # reference optimize.py:12

_PRECOND = {"jacobi": _capi.PRECOND_JACOBI, "gmg": _capi.PRECOND_GMG}
DEFAULT_RTOL_SIMP = 1e-9      # design point 1e-9..1e-10 (1e-8 is the loosest that keeps the 1e-6 trajectory bar)
DEFAULT_RTOL_ESO = 1e-12      # removal sets come from strict comparisons: solve tighter
DEFAULT_MAXIT = 100000


def _els_rows(nx, ids):
    """Rows ``[id, 1, 0, n, n+1, n+nx+2, n+nx+1]`` of the given element ids (rect_grid numbering)."""
    ids = np.asarray(ids, dtype=np.int64)
    n = ids + ids//nx
    out = np.empty((ids.size, 7), dtype=np.int64)
    out[:, 0], out[:, 1], out[:, 2] = ids, 1, 0
    out[:, 3], out[:, 4], out[:, 5], out[:, 6] = n, n + 1, n + nx + 2, n + nx + 1
    return out


def _plot_density(rho, nx, ny):
    import matplotlib.pyplot as plt                       # optional dependency, only for plot=True
    from matplotlib import colors
    plt.ion()
    fig, ax = plt.subplots()
    ax.imshow(-rho.reshape(nx, ny), cmap='gray', interpolation='none', norm=colors.Normalize(vmin=-1, vmax=0))
    ax.set_title('Predicted')
    fig.show()


def _simp_loop(topo, nx, ny, dx, dy, r_min, niter, penal, precond, rtol, maxit, history, verbose, rho0=None):
    """Body of reference optimize.py:419-456 on device-resident state."""
    Emin, Emax = 1e-9, 1.0                                  # :391-392
    change, g = 10.0, 0.0                                   # :397-398
    if rho0 is None:
        topo.fill(_capi.F_RHO, 0.5)                         # :399
    else:
        topo.set(_capi.F_RHO, rho0)
    params = _capi.SimpParams(penal=float(penal), emin=Emin, emax=Emax, rmin=float(r_min), dx=float(dx),
                              dy=float(dy), move=0.2, rtol=float(rtol), precond=_PRECOND[precond],
                              maxit=int(maxit), warm=1)
    for _ in range(niter):
        if change < 0.01:                                   # :424
            if verbose:
                print('Convergence reached')
            break
        g, res = topo.simp_step(params, g)
        change = res.change
        if history is not None:
            history.setdefault("compliance", []).append(res.compliance)
            history.setdefault("objective", []).append(res.objective)
            history.setdefault("g", []).append(g)
            history.setdefault("change", []).append(res.change)
            history.setdefault("cg_iters", []).append(res.cg_iters)
            history.setdefault("relres", []).append(res.relres)
            history.setdefault("oc_steps", []).append(res.oc_steps)
            if history.get("keep_rho", False):
                history.setdefault("rho", []).append(topo.get(_capi.F_RHO))
            if history.get("keep_dc", False):
                history.setdefault("dc", []).append(topo.get(_capi.F_DC))
        if res.status == _capi.ENOCONV or not res.equilibrium_ok:    # :455 (PCG failure = no equilibrium)
            break
    return topo.get(_capi.F_RHO)


def _simp_slab_loop(nx, ny, kloc, cons, loads, dx, dy, r_min, niter, penal, rtol, maxit, history, verbose, slabs,
                    dist, device):
    """Body of reference optimize.py:419-456 on a mesh split into row slabs (one GPU per slab with ``dist`` =
    torch.distributed, or ``slabs`` slabs on one GPU).  Returns this process' part of rho (all of it when
    ``dist`` is None)."""
    from .distributed import SlabSimp
    sim = SlabSimp(nx, ny, kloc, cons, loads, volfrac=0.5, penal=penal, emin=1e-9, emax=1.0, rmin=r_min, dx=dx, dy=dy,
                   move=0.2, rtol=rtol, maxit=maxit, world=slabs, dist=dist, device=device)
    try:
        change, g = 10.0, 0.0
        for _ in range(niter):
            if change < 0.01:
                if verbose:
                    print('Convergence reached')
                break
            g, res = sim.step(g)
            change = res["change"]
            if history is not None:
                for k in ("compliance", "objective", "change", "cg_iters", "relres", "oc_steps"):
                    history.setdefault(k, []).append(res[k])
                history.setdefault("g", []).append(g)
                if history.get("keep_rho", False):
                    history.setdefault("rho", []).append(sim.local_rho().reshape(-1))
        return sim.local_rho().reshape(-1)
    finally:
        sim.close()


def SIMP(length, height, nx, ny, dirs, positions, niter, penal, plot=False, *, n=1, r_min=None, r_factor=4,
         precond="gmg", rtol=DEFAULT_RTOL_SIMP, maxit=DEFAULT_MAXIT, device=0, history=None, verbose=True,
         slabs=None, dist=None):
    """SIMP compliance minimisation of a rectangular beam (reference optimize.py:360-465).

    Returns ``rho`` (nx*ny,) float64.  Keyword-only extras: ``n`` support selector of ``beam``
    (reference hard-codes 1, :394), ``r_min``/``r_factor`` filter radius (reference: 4 element
    widths, :404), solver controls, ``history`` dict to collect per-iteration scalars.
    ``dist`` = an initialised ``torch.distributed`` (NCCL, one process per GPU): the mesh is split into row slabs,
    one per rank, and the return value is this rank's rows of rho; ``slabs`` = N runs the same slab algorithm with
    N slabs on one GPU.
    """
    # beam() (:394) reduced to what the loop uses: BC flags, loads, node spacing (the node / element / material arrays
    # the reference also builds there are 252 MB and 0.4 s of host work at 2048x1024 and are never read again)
    cons, loads, spacing = beam_bc(length, height, nx, ny, dirs, positions, n)
    if r_min is None:
        r_min = spacing*r_factor                                                    # :404
    kloc = kloc_simp(206.8e9, 0.28)                                                 # beam()'s material defaults; :406-416
    if slabs is not None or dist is not None:
        rho = _simp_slab_loop(nx, ny, kloc, cons, loads, length/nx, height/ny, r_min, niter, penal, rtol, maxit,
                              history, verbose, slabs, dist, device)
        if plot and dist is None:
            _plot_density(rho, nx, ny)
        return rho
    with _capi.Topo(nx, ny, kloc, device=device) as topo:
        topo.set_cons(cons)
        topo.set_loads(loads)
        rho = _simp_loop(topo, nx, ny, length/nx, height/ny, r_min, niter, penal, precond, rtol, maxit,
                         history, verbose)
    if plot:
        _plot_density(rho, nx, ny)
    return rho


def _model_grid(data):
    nodes = np.asarray(data["nodes"], dtype=float)
    els = np.asarray(data["elements"]).astype(np.int64)
    nx, ny = grid_shape_from_els(els)
    nx2, ny2, dx, dy = grid_shape_from_nodes(nodes)
    if (nx, ny) != (nx2, ny2):
        raise ValueError("nodes and elements describe different grids")
    return nodes, els, nx, ny, dx, dy


def SIMP_from_model(data, niter, penal, *, r_min=None, r_factor=4, precond="gmg", rtol=DEFAULT_RTOL_SIMP,
                    maxit=DEFAULT_MAXIT, device=0, history=None, verbose=True, rho0=None, slabs=None, dist=None):
    """SIMP on a model given as the SolidsPy dict ``{"nodes","cons","elements","mats","loads"}``
    (reference README.rst:134-141), restricted to structured quad4 grids with square elements.
    ``mats[0] = (E, nu)``; ``loads`` rows ``[node, fx, fy]``; ``cons`` (N_n, 2) with -1 = constrained."""
    nodes, els, nx, ny, dx, dy = _model_grid(data)
    mats = np.asarray(data["mats"], dtype=float)
    if not np.isclose(dx, dy):
        raise ValueError("SIMP's closed-form element matrix needs square elements (dx=%g, dy=%g)" % (dx, dy))
    if r_min is None:
        r_min = np.linalg.norm(nodes[0, 1:3] - nodes[1, 1:3])*r_factor
    kloc = kloc_simp(mats[0, 0], mats[0, 1])
    if slabs is not None or dist is not None:                 # row slabs (see SIMP): this process' rows of rho
        if rho0 is not None:
            raise ValueError("rho0 is not supported on row slabs")
        return _simp_slab_loop(nx, ny, kloc, np.asarray(data["cons"]), np.asarray(data["loads"], dtype=float), dx, dy,
                               r_min, niter, penal, rtol, maxit, history, verbose, slabs, dist, device)
    with _capi.Topo(nx, ny, kloc, device=device) as topo:
        topo.set_cons(np.asarray(data["cons"]))
        topo.set_loads(np.asarray(data["loads"], dtype=float))
        return _simp_loop(topo, nx, ny, dx, dy, r_min, niter, penal, precond, rtol, maxit, history, verbose,
                          rho0=rho0)


# ------------------------------------------------------------------------------------------------
def _setup_eso(length, height, nx, ny, dirs, positions, device, model=None):
    """Arrays + device handle of the ESO/BESO set-up (reference optimize.py:52,152,251).  ``model`` = an already
    built ``(nodes, mats, els, loads, BC)`` tuple (the ``*_from_model`` entry points)."""
    if model is None:
        nodes, mats, els, loads, BC = beam(L=length, H=height, nx=nx, ny=ny, dirs=dirs, positions=positions, n=1)
    else:
        nodes, mats, els, loads, BC = model
    ke = ke_quad4(nodes[els[0, 3:7], 1:3], mats[0, 0], mats[0, 1])
    topo = _capi.Topo(nx, ny, ke, device=device)
    topo.set_cons(nodes[:, 3:5])
    topo.set_loads(loads)
    protect = np.hstack((loads[:, 0], BC)).astype(np.int64)
    return nodes, mats, els, loads, BC, topo, protect


def _model_arrays(data):
    """(length, height, nx, ny, (nodes5, mats, els, loads, BC)) from the SolidsPy dict model."""
    nodes3, els, nx, ny, dx, dy = _model_grid(data)
    cons = np.asarray(data["cons"], dtype=float).reshape(-1, 2)
    nodes = np.zeros((nodes3.shape[0], 5))
    nodes[:, :3] = nodes3[:, :3]
    nodes[:, 3:5] = cons
    m = np.asarray(data["mats"], dtype=float).reshape(-1)
    mats = np.tile([m[0], m[1], 1.0], (els.shape[0], 1))
    loads = np.asarray(data["loads"])
    BC = nodes[(cons == -1).all(axis=1), 0]          # what beam() returns: the fully clamped nodes
    return nx*dx, ny*dy, nx, ny, (nodes, mats, els, loads, BC)


GMG_ATTEMPT_CAP = int(__import__("os").environ.get("TOPO_GMG_ATTEMPT_CAP", "400"))      # PCG iterations granted to the multigrid preconditioner before the Jacobi restart (ESO/BESO)


def _solve(topo, precond, rtol, maxit, stats):
    """K u = f on the current element set (replaces spsolve, optimize.py:61,84,161,181,260,298).  ESO designs are
    hard 0/1: once thin or corner-connected members appear, the coarse multigrid levels glue parts the fine mesh
    separates and multigrid-PCG can stall (measured: 96x40 ESO_stress, fifth solve: 1e-5 after 3000 iterations where
    Jacobi-PCG needs 913).  So the multigrid attempt is capped and a stalled solve restarts, on the GPU, with the
    Jacobi preconditioner, which converges on any SPD system."""
    p = _PRECOND[precond]
    gmg = p == _capi.PRECOND_GMG
    it, rel, code = topo.solve(p, rtol, min(maxit, GMG_ATTEMPT_CAP) if gmg else maxit, warm=True, raise_on_noconv=False)
    if gmg and code == _capi.ENOCONV:
        it2, rel, code = topo.solve(_capi.PRECOND_JACOBI, rtol, maxit, warm=bool(rel < 1.0), raise_on_noconv=False)
        it += it2
        if stats is not None:
            stats["jacobi_restarts"] = stats.get("jacobi_restarts", 0) + 1
    if stats is not None:
        stats.setdefault("cg_iters", []).append(it)
        stats.setdefault("relres", []).append(rel)
    ok = (code == _capi.OK) and topo.equilibrium_ok()        # the guard of :72,:172,:286 for the NEXT pass
    return ok


def BESO(length, height, nx, ny, dirs, positions, niter, t, ER, volfrac, plot=False, *, precond="gmg",
         rtol=DEFAULT_RTOL_ESO, maxit=DEFAULT_MAXIT, device=0, history=None, verbose=True, _model=None):
    """Bi-directional ESO (reference optimize.py:215-356), quirks included: the full
    mesh is solved every iteration and removal acts by pinning the DOFs of orphaned nodes.
    Returns ``(ELS, nodes)``."""
    nodes, mats, els, loads, BC, topo, protect = _setup_eso(length, height, nx, ny, dirs, positions, device, _model)
    fields = _want_fields(plot, history)
    with topo:
        eq_ok = _solve(topo, precond, rtol, maxit, None)                           # :255-261
        _store_fields(fields, topo, mats, length/nx, height/ny, "I", present_only=False)   # :261-262
        r_min = np.linalg.norm(nodes[0, 1:3] - nodes[1, 1:3])*1                    # :264
        w4, w2x, w2y, we = beso_weights(nodes, els, nx, ny, r_min)
        vol_el = volume(els[:1], length, height, nx, ny)[0]
        nels = els.shape[0]
        V_opt = volume(els, length, height, nx, ny).sum()*volfrac                  # :268-269
        V_full = volume(els, length, height, nx, ny).sum()
        ELS_mask = None
        mask = np.ones(nels, dtype=bool)                                           # :273
        C_h = np.zeros(niter)
        error = 1000
        for i in range(niter):
            if history is not None:
                history.setdefault("wall_time", []).append(time.perf_counter())    # start of every design iteration
            if verbose:
                print("Number of elements: {}".format(nels))                       # :279
            n_kept = int(np.count_nonzero(mask))
            V = np.full(n_kept, vol_el).sum()                                      # Vi[mask].sum(), :283
            if not eq_ok or V_full < V_opt:                                        # :286
                break
            ELS_mask = mask                                                        # :290
            eq_ok = _solve(topo, precond, rtol, maxit, history)                    # :293-299
            _store_fields(fields, topo, mats, length/nx, height/ny, "", present_only=False)    # :299-300 (full `els`)
            C = 0.5*topo.compliance()                                              # :334 (same f, u)
            topo.energy_sensitivity()                                              # :304
            topo.beso_filter(w4, w2x, w2y, we, average_with_prev=(i > 0))          # :305-311
            V_r = False
            if V <= V_opt:                                                         # :315
                break
            V_k = V*(1 + ER) if V < V_opt else V*(1 - ER)                          # :320
            els_k = n_kept*V_k/V                                                   # :324
            alpha_del = topo.rank_select(int(els_k))                               # :323,:325
            n_new, n_prot = topo.beso_apply(alpha_del, protect)                    # :328-331
            mask = topo.get(_capi.F_KEEP).astype(bool)
            C_h[i] = C
            if history is not None and not history.get("timing_only", False):
                history.setdefault("kept_ids", []).append(np.flatnonzero(mask))
                history.setdefault("alpha", []).append(alpha_del)
                history.setdefault("C", []).append(C)
                history.setdefault("n_protected", []).append(n_prot)
                history.setdefault("cons", []).append(topo.get_cons())
                if history.get("keep_sens", False):
                    history.setdefault("sens", []).append(topo.get(_capi.F_SENS))
            if i > 10:
                error = C_h[i - 5:].sum() - C_h[i - 10:-5].sum()/C_h[i - 5:].sum()   # :336, as written
            if error <= t and V_r:                                                 # :339 (unreachable)
                if verbose:
                    print("convergence")
                break
            topo.beso_save_history()                                               # :344
        nodes[:, 3:5] = topo.get_cons()
    ELS = None if ELS_mask is None else _els_rows(nx, np.flatnonzero(ELS_mask))
    _finish_fields(fields, plot, history, els, ELS, nodes)                         # :346-354
    return ELS, nodes


def _want_fields(plot, history):
    """Dict that collects the plotted fields (reference: UCI, E_nodesI, S_nodesI, UC, E_nodes, S_nodes), or None.
    Asked for by ``plot=True`` or ``history["keep_fields"]`` (the fields without matplotlib)."""
    return {} if (plot or (history is not None and history.get("keep_fields", False))) else None


def _store_fields(fields, topo, mats, dx, dy, suffix, present_only):
    """Complete displacements and nodal strains / stresses of the solve just finished (pos.complete_disp +
    pos.strain_nodes after every spsolve of the reference) -- computed only when somebody will look at them."""
    if fields is None:
        return
    fields["UC" + suffix] = topo.get(_capi.F_U).reshape(-1, 2)
    fields["E_nodes" + suffix], fields["S_nodes" + suffix] = topo.strain_nodes(mats[0, 0], mats[0, 1], dx, dy,
                                                                               present_only=present_only)


def _finish_fields(fields, plot, history, elsI, ELS, nodes):
    if fields is None:
        return
    if history is not None:
        history["fields"] = fields
    if plot:
        if ELS is None or "UC" not in fields:
            # the reference would fail here with an undefined `UC` (its loop never solved); nothing to draw
            raise RuntimeError("plot=True: the optimisation loop stopped before its first solve, no fields to plot")
        from . import plots
        plots.eso_plots(elsI, ELS, nodes, fields)


def _eso(kind, length, height, nx, ny, dirs, positions, niter, RR, ER, volfrac, precond, rtol, maxit, device,
         history, verbose, model=None, plot=False, stop_on_island=False):
    """Shared body of ESO_stiff / ESO_stress.

    ``stop_on_island``: what happens once a removal cuts a piece of the structure loose (a group of present elements no
    longer tied to a support) is NOT defined by the reference's algorithm but by its direct solver: the stiffness
    matrix is singular, and ``spsolve`` either meets an exactly zero pivot -- the solution is NaN, nothing is removed
    in that pass and the guard (:72 / :172) ends the loop, as in the reference fixture eso_stress c2 -- or rounding
    leaves a tiny pivot, a finite solution comes back and the loop carries on (the untouched reference on 96x40 and
    60x24, ``tests/test_oracle_t1.py``).  The PCG solve here always returns the finite solution of the consistent
    system (default, False); True reproduces the NaN path: the floating part is detected on the device
    (``topo_floating_elements``) and the loop ends exactly as the reference's does after a NaN solve."""
    nodes, mats, els, loads, BC, topo, protect = _setup_eso(length, height, nx, ny, dirs, positions, device, model)
    fields = _want_fields(plot, history)
    support = np.flatnonzero((nodes[:, 3:5] == -1).any(axis=1)).astype(np.int64)
    want_floating = stop_on_island or (history is not None and history.get("keep_floating", False))
    island = False
    with topo:
        eq_ok = _solve(topo, precond, rtol, maxit, None)                           # :56-61 / :156-161
        _store_fields(fields, topo, mats, length/nx, height/ny, "I", present_only=True)    # :62-63 / :162-163
        if kind == "stiff":
            niter, RR, ER = 200, 0.005, 0.05                                       # :165-167 overrides the caller
        vol_el = volume(els[:1], length, height, nx, ny)[0]
        V_opt = volume(els, length, height, nx, ny).sum()*volfrac
        mask = np.ones(els.shape[0], dtype=bool)
        ELS_mask = None
        for _ in range(niter):
            n_present = int(np.count_nonzero(mask))
            if kind == "stress" and verbose:
                print("Number of elements: {}".format(n_present))                  # :69
            if not eq_ok or np.full(n_present, vol_el).sum() < V_opt:              # :72 / :172
                if kind == "stress" and verbose:
                    print('hollaa')                                                # :73
                break
            ELS_mask = mask                                                        # :76 / :192
            if island and stop_on_island:
                # the reference's solve returns NaN on this (singular) system: every comparison with NaN is False, so
                # nothing is removed, and the guard ends the loop at the top of the next pass
                eq_ok = False
                if history is not None:
                    history.setdefault("els_ids", []).append(np.flatnonzero(mask))
                    history.setdefault("cons", []).append(topo.get_cons())
                RR += ER
                continue
            eq_ok = _solve(topo, precond, rtol, maxit, history)
            _store_fields(fields, topo, mats, length/nx, height/ny, "", present_only=True)     # :85-86 / :182-183
            if kind == "stiff":
                if verbose:
                    print("Number of elements: {}".format(n_present))              # :185
                topo.energy_sensitivity()                                          # :188
            else:
                topo.vonmises_sensitivity(mats[0, 0], mats[0, 1], length/nx, height/ny)   # :86-92
            topo.eso_apply(RR, protect)                                            # :93-97 / :189-196
            mask = topo.get(_capi.F_KEEP).astype(bool)
            if want_floating:
                n_float = topo.floating_elements(support)
                island = n_float > 0
                if history is not None:
                    history.setdefault("floating", []).append(n_float)
            if history is not None:
                history.setdefault("els_ids", []).append(np.flatnonzero(mask))
                history.setdefault("cons", []).append(topo.get_cons())
            RR += ER                                                               # :99 / :198
        nodes[:, 3:5] = topo.get_cons()
    ELS = None if ELS_mask is None else _els_rows(nx, np.flatnonzero(ELS_mask))
    _finish_fields(fields, plot, history, els, ELS, nodes)                         # :101-109 / :200-208
    return ELS, nodes


def ESO_stiff(length, height, nx, ny, dirs, positions, niter, RR, ER, volfrac, plot=False, *, precond="gmg",
              rtol=DEFAULT_RTOL_ESO, maxit=DEFAULT_MAXIT, device=0, history=None, verbose=True, stop_on_island=False):
    """Stiffness-based ESO (reference optimize.py:116-210).  As in the reference, ``niter``, ``RR`` and
    ``ER`` are overridden by 200, 0.005, 0.05 (:165-167).  Returns ``(ELS, nodes)``.  ``stop_on_island``: see ``_eso``."""
    return _eso("stiff", length, height, nx, ny, dirs, positions, niter, RR, ER, volfrac, precond, rtol, maxit,
                device, history, verbose, plot=plot, stop_on_island=stop_on_island)


def ESO_stress(length, height, nx, ny, dirs, positions, niter, RR, ER, volfrac, plot=False, *, precond="gmg",
               rtol=DEFAULT_RTOL_ESO, maxit=DEFAULT_MAXIT, device=0, history=None, verbose=True, stop_on_island=False):
    """Stress-based ESO (reference optimize.py:16-111): von Mises of element-averaged nodal stresses.
    Returns ``(ELS, nodes)``.  ``stop_on_island``: see ``_eso``."""
    return _eso("stress", length, height, nx, ny, dirs, positions, niter, RR, ER, volfrac, precond, rtol, maxit,
                device, history, verbose, plot=plot, stop_on_island=stop_on_island)


def BESO_from_model(data, niter, t, ER, volfrac, **kw):
    """``BESO`` on a model given as the SolidsPy dict ``{"nodes","cons","elements","mats","loads"}`` (structured
    quad4 grid).  Protected nodes are the loaded nodes and the fully clamped ones, as ``beam()`` defines ``BC``."""
    L, H, nx, ny, model = _model_arrays(data)
    return BESO(L, H, nx, ny, None, None, niter, t, ER, volfrac, _model=model, **kw)


def ESO_stiff_from_model(data, niter, RR, ER, volfrac, *, precond="gmg", rtol=DEFAULT_RTOL_ESO, maxit=DEFAULT_MAXIT,
                         device=0, history=None, verbose=True, stop_on_island=False):
    """``ESO_stiff`` on a dict model (niter/RR/ER are overridden exactly as in the reference, :165-167)."""
    L, H, nx, ny, model = _model_arrays(data)
    return _eso("stiff", L, H, nx, ny, None, None, niter, RR, ER, volfrac, precond, rtol, maxit, device, history,
                verbose, model, stop_on_island=stop_on_island)


def ESO_stress_from_model(data, niter, RR, ER, volfrac, *, precond="gmg", rtol=DEFAULT_RTOL_ESO, maxit=DEFAULT_MAXIT,
                          device=0, history=None, verbose=True, stop_on_island=False):
    """``ESO_stress`` on a dict model."""
    L, H, nx, ny, model = _model_arrays(data)
    return _eso("stress", L, H, nx, ny, None, None, niter, RR, ER, volfrac, precond, rtol, maxit, device, history,
                verbose, model, stop_on_island=stop_on_island)
