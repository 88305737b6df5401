"""Import shim that lets the UNMODIFIED reference package run in the authoring container.

Only `tests/golden/make_golden.py` uses this file, and only in the authoring
container: `/root/reference` does not exist on the GPU box, so nothing that runs
there may import this module.

Why a shim is needed: the reference imports gurobipy / cplex /
mosek at module top (`solver_caller/caller.py:7-11`, `solving.py:8-10`); none is
installed offline.  Empty stub modules satisfy the imports; the hot path never
touches them.  `np.Inf` (used at `net_manager.py:148`) was removed in NumPy 2.
"""
import contextlib
import sys
import types

import numpy as np

REFERENCE_SRC = "/root/reference/src"


def install():
    """Register the stubs and put the reference on sys.path.  Idempotent."""
    if "gurobipy" not in sys.modules:
        g = types.ModuleType("gurobipy")
        g.GRB = type("GRB", (), {})
        g.Model = type("Model", (), {})
        sys.modules["gurobipy"] = g
    if "cplex" not in sys.modules:
        c = types.ModuleType("cplex")
        c.Cplex = type("Cplex", (), {})
        sys.modules["cplex"] = c
    if "mosek" not in sys.modules:
        m = types.ModuleType("mosek")
        mf = types.ModuleType("mosek.fusion")
        mf.Model = type("Model", (), {})
        m.fusion = mf
        sys.modules["mosek"] = m
        sys.modules["mosek.fusion"] = mf
    if not hasattr(np, "Inf"):
        np.Inf = np.inf
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)


@contextlib.contextmanager
def stable_argsort():
    """Force `np.argsort` calls that pass no `kind` to use kind='stable'.

    BASELINE.json north_star: "ties broken by arc index, with the reference run
    with a stable sort" (`net_manager.py:184,379` call argsort with no kind).
    """
    orig = np.argsort

    def patched(a, axis=-1, kind=None, order=None, **kw):
        return orig(a, axis=axis, kind=kind or "stable", order=order, **kw)

    np.argsort = patched
    try:
        yield
    finally:
        np.argsort = orig
