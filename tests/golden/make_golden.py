#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ by running the UNMODIFIED reference.

Run in the authoring container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference ships no tests, golden vectors or fixtures, so
parity is pinned on outputs of the reference itself: its functions are imported
through `_ref_shim` (stub solver modules, `np.Inf`), `np.argsort` is forced
stable while `get_sorted_flows` runs (north_star), and the restricted-master LP
re-solve goes through a HiGHS `SolverCaller` patched into
`solver_caller/solving.py:13-29` because no vendor solver is installed.

Small cases store inputs and outputs in full (`*.npz`); large cases store the
seed, a digest of the regenerated inputs, digests of the large outputs and the
small outputs in full.
"""
import importlib.util
import json
import os
import sys

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_shim  # noqa: E402
import cases  # noqa: E402

_ref_shim.install()

from smart_crossover.formats import MinCostFlow, OptTransport  # noqa: E402  (reference)
from smart_crossover.network_methods import algorithms as ref_alg  # noqa: E402
from smart_crossover.network_methods.net_manager import MCFManagerStd, OTManager  # noqa: E402
from smart_crossover.network_methods.tree_BI import (max_weight_spanning_tree,  # noqa: E402
                                                     push_tree_to_bfs)
from smart_crossover.solver_caller import solving as ref_solving  # noqa: E402
from smart_crossover.solver_caller.caller import SolverSettings  # noqa: E402

# HiGHS adapter from the product tree, loaded by path so that its
# `from smart_crossover.output import ...` binds to the reference's classes here.
_spec = importlib.util.spec_from_file_location(
    "sx_highs_for_golden",
    os.path.join(HERE, "..", "..", "smart-crossover_b200", "smart_crossover", "solver_caller", "highs.py"))
_highs = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_highs)
ref_solving.generate_solver_caller = lambda solver="GRB", solver_settings=None: _highs.HgsCaller(solver_settings)

QUIET = SolverSettings(log_console=0)


def highs_interior_point(ot):
    """Barrier point without crossover = the `x_bar` the reference's scripts feed in
    (scripts/run_network_crossover.py:117-121)."""
    out = ref_solving.solve_mcf(ot.to_MCF(), solver="HGS", method="barrier",
                                settings=SolverSettings(crossover="off", log_console=0))
    return np.maximum(out.x, 0.0)


def run_ot_reference(s, d, M, x, full_pipeline=True, e2e=True, K=64):
    """Every hot-path output of the reference for one OT instance."""
    S, D = M.shape
    ot = OptTransport(s.copy(), d.copy(), M.copy())
    mgr = OTManager(ot)
    with _ref_shim.stable_argsort():
        queue, scores = mgr.get_sorted_flows(x)
    out = {"scores": scores, "queue": queue.astype(np.int64)}
    tree = max_weight_spanning_tree(ot, scores)
    out["tree"] = tree.astype(np.int64)
    if not full_pipeline or tree.size != S + D - 1:
        return out
    mgr.get_mcf()
    A = mgr.mcf.A.tocsc()
    B = A[:-1, :][:, tree]
    out["tree_flows"] = sp.linalg.spsolve(B, mgr.mcf.b[:-1])
    vbasis, push_iter = push_tree_to_bfs(mgr, tree)
    out["vbasis_tree"] = vbasis.astype(np.int64)
    out["push_iter"] = np.int64(push_iter)
    # a5: potentials of the (pre-push) spanning tree, SURVEY.md section 8c definition (3)
    y = np.zeros(S + D)
    y[:-1] = sp.linalg.spsolve(B.T.tocsc(), mgr.mcf.c[tree])
    out["y_tree"] = y
    out["rc_tree"] = mgr.get_reduced_cost_for_original_OT(y)
    out["optimal_tree"] = np.bool_(mgr.check_optimality_condition(x, y))
    # a perturbed dual so that a sizeable set of arcs violates
    rng = np.random.default_rng(12345)
    y2 = y + 0.05 * np.abs(M).max() * (rng.random(S + D) - 0.5)
    out["y_pert"] = y2
    out["rc_pert"] = mgr.get_reduced_cost_for_original_OT(y2)
    out["optimal_pert"] = np.bool_(mgr.check_optimality_condition(x, y2))
    if e2e:
        for method in ("tnet", "cnet_ot"):
            res = ref_alg.network_crossover(x=x.copy(), ot=OptTransport(s.copy(), d.copy(), M.copy()),
                                            method=method, solver="HGS", solver_settings=QUIET)
            out[f"{method}_obj"] = np.float64(res.obj_val)
            out[f"{method}_iters"] = np.int64(res.iter_count)
            out[f"{method}_basic"] = np.flatnonzero(res.basis.vbasis == 0).astype(np.int64)
    return out


def build_mcf(tail, head, b, c, u):
    E, N = tail.size, b.size
    A = sp.lil_matrix((N, E), dtype=int)      # scripts/min2mcf.py:31-37
    A[head, np.arange(E)] = -1
    A[tail, np.arange(E)] = 1
    return MinCostFlow(A=A.tocsr(), b=b.copy(), c=c.copy(), u=u.copy())


def run_mcf_reference(tail, head, b, c, u, x, e2e=True):
    mcf = build_mcf(tail, head, b, c, u)
    mgr = MCFManagerStd(mcf)
    with _ref_shim.stable_argsort():
        queue, ind = mgr.get_sorted_flows(x)
    out = {"scores": ind, "queue": queue.astype(np.int64)}
    rng = np.random.default_rng(777)
    y = rng.random(b.size) * 100.0
    vbasis = -np.ones(c.size, dtype=np.int64)
    vbasis[rng.random(c.size) < 0.1] = -2
    vbasis[rng.random(c.size) < 0.05] = 0
    from smart_crossover.output import Basis
    mgr.set_basis(Basis(vbasis, -np.ones(b.size)))
    out["y"] = y
    out["vbasis"] = vbasis
    out["rc"] = mgr.get_reduced_cost_for_original_mcf(y)
    out["optimal"] = np.bool_(mgr.check_optimality_condition(x, y))
    if e2e:
        res = ref_alg.network_crossover(x=x.copy(), mcf=build_mcf(tail, head, b, c, u),
                                        method="cnet_mcf", solver="HGS", solver_settings=QUIET)
        out["cnet_mcf_obj"] = np.float64(res.obj_val)
        out["cnet_mcf_iters"] = np.int64(res.iter_count)
        direct = ref_solving.solve_mcf(build_mcf(tail, head, b, c, u), solver="HGS", settings=QUIET)
        out["direct_obj"] = np.float64(direct.obj_val)
    return out


def save_full(name, inputs, outputs):
    np.savez_compressed(os.path.join(HERE, name + ".npz"),
                        **{"in_" + k: v for k, v in inputs.items()},
                        **{"out_" + k: v for k, v in outputs.items()})
    print("wrote", name, {k: np.shape(v) for k, v in outputs.items()})


def main():
    # ---- small OT cases, stored in full ---------------------------------
    s, d, M = cases.ot_points(40, 40, 20260001)          # C1
    x = highs_interior_point(OptTransport(s, d, M))
    x = np.maximum(x, 1e-14)                             # strictly positive
    save_full("ot_c1_40x40", dict(s=s, d=d, M=M, x=x), run_ot_reference(s, d, M, x))

    s, d, M = cases.ot_points(12, 9, 11)                 # non-square, odd D, tie-heavy flow
    x = cases.product_flow(s, d)
    save_full("ot_ties_12x9", dict(s=s, d=d, M=M, x=x), run_ot_reference(s, d, M, x))

    s, d, M = cases.ot_points(30, 51, 12)                # rational-kernel interior flow, odd D
    x = cases.interior_flow(s, d, M, 12, 0.33)
    save_full("ot_rational_30x51", dict(s=s, d=d, M=M, x=x), run_ot_reference(s, d, M, x))

    s, d, M = cases.ot_grid(6, 13)                       # 36x36 integer cost: tied reduced costs
    x = cases.interior_flow(s, d, M, 13, 12.0)
    save_full("ot_grid_36x36", dict(s=s, d=d, M=M, x=x), run_ot_reference(s, d, M, x))

    # zero-weight quirk: tree arcs of weight 0 vanish from the reference's result
    s = np.array([0.5, 0.25, 0.25]); d = np.array([0.25, 0.25, 0.5])
    M = np.arange(9, dtype=np.float64).reshape(3, 3)
    x = np.array([0.25, 0, 0, 0, 0.25, 0, 0, 0, 0.25], dtype=np.float64)
    save_full("ot_zero_3x3", dict(s=s, d=d, M=M, x=x),
              run_ot_reference(s, d, M, x, full_pipeline=False))

    # ---- small MCF cases, stored in full ----------------------------------
    tail, head, b, c, u = cases.netgen_like(200, 1500, 21, n_supply=8)
    x = cases.mcf_interior_flow(u, 21)
    save_full("mcf_small_200", dict(tail=tail, head=head, b=b, c=c, u=u, x=x),
              run_mcf_reference(tail, head, b, c, u, x))

    tail, head, b, c, u = cases.netgen_like(60, 400, 22, n_supply=4)
    rng = np.random.default_rng(22)
    x = np.floor(rng.random(u.size) * (u + 1.0))         # integer flows: ties, x == u, x == 0
    x[::17] = -1.0                                       # out of bounds -> x_hat = 0
    x[5::23] = u[5::23] + 2.0
    save_full("mcf_ties_60", dict(tail=tail, head=head, b=b, c=c, u=u, x=x),
              run_mcf_reference(tail, head, b, c, u, x, e2e=False))

    # ---- large cases: digests only -----------------------------------------
    meta = {}
    s, d, M = cases.ot_grid(28, 20260002)                # C2: 784 x 784
    x = cases.interior_flow(s, d, M, 20260002, 260.0)
    out = run_ot_reference(s, d, M, x, e2e=False)
    meta["ot_c2_784"] = {"inputs": cases.digest(s, d, M, x),
                         "digests": {k: cases.digest(v) for k, v in out.items()}}
    cnt = int((out["rc_pert"] < -1e-6).sum())
    o = np.argsort(out["rc_pert"], kind="stable")[:256]
    small = {k: out[k] for k in ("tree", "y_tree", "y_pert", "push_iter", "optimal_tree", "optimal_pert")}
    small["count_pert"] = np.int64(cnt)
    small["topk_ids_pert"] = o[out["rc_pert"][o] < -1e-6].astype(np.int64)
    small["min_rc_pert"] = np.float64(out["rc_pert"].min())
    small["basic_tree"] = np.flatnonzero(out["vbasis_tree"] == 0).astype(np.int64)
    save_full("ot_c2_784_small", {}, small)

    tail, head, b, c, u = cases.netgen_like(20000, 200000, 23)
    x = cases.mcf_interior_flow(u, 23)
    out = run_mcf_reference(tail, head, b, c, u, x, e2e=False)
    meta["mcf_mid_20k"] = {"inputs": cases.digest(tail, head, b, c, u, x),
                           "digests": {k: cases.digest(v) for k, v in out.items()}}
    save_full("mcf_mid_20k_small", {}, {"y": out["y"], "optimal": out["optimal"],
                                         "count": np.int64((out["rc"] < -1e-6).sum())})

    with open(os.path.join(HERE, "digests.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print("wrote digests.json")


if __name__ == "__main__":
    main()
