"""T0 oracle: execute the UNTOUCHED reference source and record what it computes per iteration.

ORACLE / TEST INFRASTRUCTURE ONLY.  Never imported by the product path; only runs where
``/root/reference`` is mounted (this container), never on the GPU box.  Its outputs are frozen as
fixtures under ``tests/golden/`` by ``oracle/gen_golden.py``.

The reference (``/root/reference/src/solidspy_opt/optimize.py``) returns only the final ``rho`` (SIMP,
``:465``) or ``(ELS, nodes)`` (``:111,210,356``) and stores no history, so the per-iteration quantities
the parity bar is defined on (compliance, density, removal sets) are captured by wrapping the *names the
reference loop looks up in its own module namespace* (``spsolve``, ``optimality_criteria``,
``density_filter``, ``del_node`` ...).  The reference files themselves are not modified or copied.
"""
from __future__ import annotations

import contextlib
import importlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_SRC = os.environ.get("SOLIDSPY_OPT_REFERENCE", "/root/reference/src")


def load_reference():
    """Import the reference package on top of the compat ``solidspy`` and the ``matplotlib`` stub."""
    if not os.path.isdir(os.path.join(REFERENCE_SRC, "solidspy_opt")):
        raise RuntimeError("reference source not mounted at %s" % REFERENCE_SRC)
    for p in (os.path.join(HERE, "compat"), REFERENCE_SRC):
        if p not in sys.path:
            sys.path.insert(0, p)
    mod = importlib.import_module("solidspy_opt.optimize")
    if not os.path.abspath(mod.__file__).startswith(os.path.abspath(REFERENCE_SRC)):
        raise RuntimeError("solidspy_opt resolved to %s, not the reference" % mod.__file__)
    return mod


class _Tap:
    """Replace ``module.name`` by a wrapper that calls the original and hands (args, result) to ``fn``."""

    def __init__(self, module, name, fn):
        self.module, self.name, self.fn = module, name, fn
        self.orig = getattr(module, name)

    def __enter__(self):
        orig, fn = self.orig, self.fn

        def wrapper(*a, **k):
            out = orig(*a, **k)
            fn(a, k, out)
            return out
        setattr(self.module, self.name, wrapper)
        return self

    def __exit__(self, *exc):
        setattr(self.module, self.name, self.orig)


def _quiet(call, *a, **k):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        out = call(*a, **k)
    return out, buf.getvalue()


def run_simp(length, height, nx, ny, dirs, positions, niter, penal):
    """Reference ``SIMP`` (optimize.py:360) with per-iteration capture.

    Returns dict: rho_final, and per-iteration lists compliance (f.u, :440), rho (after OC, :449),
    g (:449), dc_filtered (:445), disp-norm, neq.
    """
    opt = load_reference()
    rec = {"compliance": [], "rho": [], "g": [], "dc": [], "u_absmax": [], "rhs": None}

    def on_loadasem(a, k, out):
        rec["rhs"] = out.copy()

    def on_solve(a, k, out):
        rec["compliance"].append(float(rec["rhs"].dot(out)))
        rec["u_absmax"].append(float(np.abs(out).max()))

    def on_filter(a, k, out):
        rec["dc"].append(np.array(out, dtype=float))

    def on_oc(a, k, out):
        rec["rho"].append(np.array(out[0], dtype=float))
        rec["g"].append(float(out[1]))

    with _Tap(opt.ass, "loadasem", on_loadasem), _Tap(opt, "spsolve", on_solve), \
            _Tap(opt, "density_filter", on_filter), _Tap(opt, "optimality_criteria", on_oc):
        rho, _ = _quiet(opt.SIMP, length, height, nx, ny, np.asarray(dirs), np.asarray(positions),
                        niter, penal, plot=False)
    return {
        "rho_final": np.asarray(rho),
        "compliance": np.array(rec["compliance"]),
        "rho": np.array(rec["rho"]),
        "g": np.array(rec["g"]),
        "dc": np.array(rec["dc"]),
        "u_absmax": np.array(rec["u_absmax"]),
    }


def run_beso(length, height, nx, ny, dirs, positions, niter, t, ER, volfrac):
    """Reference ``BESO`` (optimize.py:215) with capture of masks / sensitivities / pinned nodes."""
    opt = load_reference()
    rec = {"mask": [], "sensi": [], "cons": [], "C": [], "rhs": None, "alpha": [], "n_protected": []}

    def on_loadasem(a, k, out):
        rec["rhs"] = out.copy()

    def on_solve(a, k, out):
        rec["C"].append(0.5*float(rec["rhs"].dot(out)))

    def on_filter(a, k, out):           # sensitivity_filter output (before history average, :306)
        rec["sensi"].append(np.array(out, dtype=float))

    def on_protect(a, k, out):          # protect_els(els[~mask], nels, loads, BC) (:329)
        rec["n_protected"].append(int(np.count_nonzero(out)))

    def on_del_node(a, k, out):         # del_node(nodes, els[mask], loads, BC) (:331)
        nodes, els_kept = a[0], a[1]
        rec["mask"].append(np.array(els_kept[:, 0], dtype=np.int64))
        rec["cons"].append(np.array(nodes[:, 3:5], dtype=np.int8))

    with _Tap(opt.ass, "loadasem", on_loadasem), _Tap(opt, "spsolve", on_solve), \
            _Tap(opt, "sensitivity_filter", on_filter), _Tap(opt, "protect_els", on_protect), \
            _Tap(opt, "del_node", on_del_node):
        (ELS, nodes), _ = _quiet(opt.BESO, length, height, nx, ny, np.asarray(dirs), np.asarray(positions),
                                 niter, t, ER, volfrac, plot=False)
    return {
        "ELS": np.asarray(ELS), "nodes": np.asarray(nodes),
        "kept_ids": rec["mask"],                 # list of int arrays: kept element ids after each removal
        "cons": np.array(rec["cons"]),           # BC flags after each del_node
        "sensi_filtered": np.array(rec["sensi"]),
        "C": np.array(rec["C"]),                 # includes the setup solve (:260) as entry 0
        "n_protected": np.array(rec["n_protected"]),
    }


def _run_eso(which, length, height, nx, ny, dirs, positions, niter, RR, ER, volfrac):
    opt = load_reference()
    rec = {"els_ids": [], "cons": [], "C": [], "rhs": None}

    def on_loadasem(a, k, out):
        rec["rhs"] = out.copy()

    def on_solve(a, k, out):
        rec["C"].append(0.5*float(rec["rhs"].dot(out)))

    def on_del(a, k, out):              # del_nodeESO(nodes, els) after np.delete (:97,:196)
        rec["els_ids"].append(np.array(a[1][:, 0], dtype=np.int64))
        rec["cons"].append(np.array(a[0][:, 3:5], dtype=np.int8))

    fn = getattr(opt, which)
    with _Tap(opt.ass, "loadasem", on_loadasem), _Tap(opt, "spsolve", on_solve), \
            _Tap(opt, "del_nodeESO", on_del):
        (ELS, nodes), _ = _quiet(fn, length, height, nx, ny, np.asarray(dirs), np.asarray(positions),
                                 niter, RR, ER, volfrac, plot=False)
    return {"ELS": np.asarray(ELS), "nodes": np.asarray(nodes), "els_ids": rec["els_ids"],
            "cons": np.array(rec["cons"]), "C": np.array(rec["C"])}


def run_eso_stiff(*a):
    """Reference ``ESO_stiff`` (optimize.py:116)."""
    return _run_eso("ESO_stiff", *a)


def run_eso_stress(*a):
    """Reference ``ESO_stress`` (optimize.py:16)."""
    return _run_eso("ESO_stress", *a)
