"""End-to-end drop-in check on a B200: `network_crossover` (tnet / cnet_ot / cnet_mcf) through the
device-backed managers + HiGHS, against the objective and basis the reference produced for the same
inputs (golden fixtures), and the manager-level calls against the reference's own outputs."""
import numpy as np
import pytest
import scipy.sparse as sp

from golden_util import MCF_FULL, OT_FULL, Fixture

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module", autouse=True)
def _need_cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def _quiet():
    from smart_crossover.solver_caller.caller import SolverSettings
    return SolverSettings(log_console=0)


@pytest.mark.parametrize("name", OT_FULL)
@pytest.mark.parametrize("method", ["tnet", "cnet_ot"])
def test_network_crossover_ot(name, method):
    from smart_crossover.formats import OptTransport
    from smart_crossover.network_methods.algorithms import network_crossover
    fx = Fixture(name)
    ot = OptTransport(fx.inp["s"].copy(), fx.inp["d"].copy(), fx.inp["M"].copy())
    out = network_crossover(x=fx.inp["x"].copy(), ot=ot, method=method, solver="HGS", solver_settings=_quiet())
    ref_obj = float(fx.out[f"{method}_obj"])
    assert abs(out.obj_val - ref_obj) <= 1e-9 * max(1.0, abs(ref_obj))
    if name != "ot_ties_12x9":
        # generic inputs: the optimal basis is unique and the simplex path is reproducible.  The
        # tie-heavy instance (x = s d^T) is degenerate: HiGHS may stop at another optimal basis /
        # take a different pivot count on another host, so only the objective is pinned there.
        assert np.array_equal(np.flatnonzero(out.basis.vbasis == 0), fx.out[f"{method}_basic"])
        assert out.iter_count == int(fx.out[f"{method}_iters"])


def test_network_crossover_mcf():
    from smart_crossover.formats import MinCostFlow
    from smart_crossover.network_methods.algorithms import network_crossover
    fx = Fixture("mcf_small_200")
    tail, head, b, c, u, x = (fx.inp[k] for k in ("tail", "head", "b", "c", "u", "x"))
    E, N = c.size, b.size
    A = sp.lil_matrix((N, E), dtype=int)
    A[head, np.arange(E)] = -1
    A[tail, np.arange(E)] = 1
    mcf = MinCostFlow(A=A.tocsr(), b=b.copy(), c=c.copy(), u=u.copy())
    out = network_crossover(x=x.copy(), mcf=mcf, method="cnet_mcf", solver="HGS", solver_settings=_quiet())
    ref = float(fx.out["cnet_mcf_obj"])
    assert abs(out.obj_val - ref) <= 1e-9 * abs(ref)
    assert abs(out.obj_val - float(fx.out["direct_obj"])) <= 1e-9 * abs(ref)
    assert out.iter_count == int(fx.out["cnet_mcf_iters"])
    assert mcf.c.max() == 1.0            # documented side effect: the caller's costs are rescaled


@pytest.mark.parametrize("name", OT_FULL)
def test_manager_calls_match_the_reference(name):
    from smart_crossover.formats import OptTransport
    from smart_crossover.network_methods.net_manager import OTManager
    from smart_crossover.network_methods.tree_BI import (max_weight_spanning_tree, tree_basis_identify,
                                                         tree_potentials)
    fx = Fixture(name)
    ot = OptTransport(fx.inp["s"], fx.inp["d"], fx.inp["M"])
    mgr = OTManager(ot)
    queue, scores = mgr.get_sorted_flows(fx.inp["x"])
    assert queue.dtype == np.int64 and np.array_equal(queue, fx.out["queue"])
    assert scores.tobytes() == fx.out["scores"].tobytes()
    assert np.array_equal(max_weight_spanning_tree(ot, scores), fx.out["tree"])
    basis, push_iter = tree_basis_identify(mgr, scores)
    assert push_iter == int(fx.out["push_iter"]) and np.array_equal(basis.vbasis, fx.out["vbasis_tree"])
    assert basis.cbasis[-1] == 0 and (basis.cbasis[:-1] == -1).all()
    y = tree_potentials(ot, fx.out["tree"])
    np.testing.assert_allclose(y, fx.out["y_tree"], rtol=1e-9, atol=1e-9 * np.abs(fx.inp["M"]).max())
    for tag in ("tree", "pert"):
        rc = mgr.get_reduced_cost_for_original_OT(fx.out["y_" + tag])
        assert rc.tobytes() == fx.out["rc_" + tag].tobytes()
        assert mgr.check_optimality_condition(fx.inp["x"], fx.out["y_" + tag]) == bool(fx.out["optimal_" + tag])


def test_extended_ot_pricing_matches_explicit_extension():
    """cnet_ot prices the (S+1) x (D+1) big-M problem without copying M to the device: compare with
    the reference formula applied to the explicitly extended matrix."""
    from smart_crossover.formats import OptTransport
    from smart_crossover.network_methods.net_manager import OTManager
    from oracle import network_oracle as orc
    fx = Fixture("ot_rational_30x51")
    S, D = fx.inp["M"].shape
    mgr = OTManager(OptTransport(fx.inp["s"], fx.inp["d"], fx.inp["M"]))
    mgr.extend_by_bigM(mgr.m * fx.inp["M"].max())
    rng = np.random.default_rng(3)
    y = rng.random(S + D + 2) * 0.3
    rc_ref = orc.reduced_costs_ot(mgr.ot.M, y)
    assert mgr.get_reduced_cost_for_original_OT(y).tobytes() == rc_ref.tobytes()
    cnt, mn, ids, vals = orc.price_summary(rc_ref, K=40)
    res = mgr.price(y, K=40)
    assert (res.n_violating, res.min_rc) == (cnt, mn)
    assert np.array_equal(res.topk_id, ids) and np.array_equal(res.topk_rc, vals)


@pytest.mark.parametrize("name", MCF_FULL)
def test_mcf_manager_calls_match_the_reference(name):
    from smart_crossover.formats import MinCostFlow
    from smart_crossover.network_methods.net_manager import MCFManagerStd
    from smart_crossover.output import Basis
    fx = Fixture(name)
    tail, head, b, c, u, x = (fx.inp[k] for k in ("tail", "head", "b", "c", "u", "x"))
    E, N = c.size, b.size
    A = sp.lil_matrix((N, E), dtype=int)
    A[head, np.arange(E)] = -1
    A[tail, np.arange(E)] = 1
    mgr = MCFManagerStd(MinCostFlow(A=A.tocsr(), b=b.copy(), c=c.copy(), u=u.copy()))
    queue, scores = mgr.get_sorted_flows(x)
    assert np.array_equal(queue, fx.out["queue"]) and scores.tobytes() == fx.out["scores"].tobytes()
    mgr.set_basis(Basis(fx.out["vbasis"], -np.ones(N)))
    assert mgr.get_reduced_cost_for_original_mcf(fx.out["y"]).tobytes() == fx.out["rc"].tobytes()
    assert mgr.check_optimality_condition(x, fx.out["y"]) == bool(fx.out["optimal"])


# ---- BASELINE.json configs[1] (784 x 784) and the 20 000-node MCF: goldens from make_golden_e2e.py ---------
def test_c2_device_tree_flows_through_the_push_phase():
    """784 x 784: tree flows computed on the DEVICE (Euler tour, double-double subtree sums) feed the native
    push loop; the reference gets them from SuperLU (tree_BI.py:74-76).  A rounding difference that flips a
    `< 0` or `== 0` test would change the push count (934) or the basis."""
    from golden_util import c2_inputs
    from smart_crossover.formats import OptTransport
    from smart_crossover.network_methods.net_manager import OTManager
    from smart_crossover.network_methods.tree_BI import push_tree_to_bfs, tree_basis_identify, tree_flows
    s, d, M, x = c2_inputs()
    small = Fixture("ot_c2_784_small")
    ot = OptTransport(s, d, M)
    mgr = OTManager(ot)
    tree = small.out["tree"]
    vbasis, push_iter = push_tree_to_bfs(mgr, tree)                 # device flows -> native push
    assert push_iter == int(small.out["push_iter"]) == 934
    assert np.array_equal(np.flatnonzero(vbasis == 0), small.out["basic_tree"])
    # the same through the public entry: scores -> device tree -> flows -> push
    queue, scores = mgr.get_sorted_flows(x)
    basis, it = tree_basis_identify(mgr, scores)
    assert it == 934 and np.array_equal(np.flatnonzero(basis.vbasis == 0), small.out["basic_tree"])
    flows = tree_flows(ot, tree)
    assert flows.shape == (784 + 784 - 1,) and (flows < 0).sum() > 0      # the push had something to do


@pytest.mark.parametrize("method", ["tnet", "cnet_ot"])
def test_network_crossover_c2_784(method):
    from golden_util import c2_inputs
    from smart_crossover.formats import OptTransport
    from smart_crossover.network_methods.algorithms import network_crossover
    s, d, M, x = c2_inputs()
    fx = Fixture("e2e_c2_784")
    out = network_crossover(x=x.copy(), ot=OptTransport(s.copy(), d.copy(), M.copy()), method=method, solver="HGS",
                            solver_settings=_quiet())
    ref = float(fx.out[f"{method}_obj"])
    assert abs(out.obj_val - ref) <= 1e-9 * abs(ref)
    assert abs(out.obj_val - float(fx.out["direct_obj"])) <= 1e-9 * abs(ref)
    # same restricted masters, same warm-start bases, same solver => same simplex path as the reference run
    assert out.iter_count == int(fx.out[f"{method}_iters"])
    assert np.array_equal(np.flatnonzero(out.basis.vbasis == 0), fx.out[f"{method}_basic"])


def test_network_crossover_mcf_20k():
    """cnet_mcf on 20 000 nodes / 200 000 arcs (the reference needs 87 s here, nearly all of it in HiGHS)."""
    from golden_util import mcf_mid_inputs
    from smart_crossover.formats import MinCostFlow
    from smart_crossover.network_methods.algorithms import network_crossover
    tail, head, b, c, u, x = mcf_mid_inputs()
    E, N = c.size, b.size
    A = sp.csr_matrix((np.concatenate([np.ones(E, dtype=int), -np.ones(E, dtype=int)]),
                       (np.concatenate([tail, head]), np.concatenate([np.arange(E)] * 2))), shape=(N, E))
    fx = Fixture("e2e_mcf_20k")
    out = network_crossover(x=x.copy(), mcf=MinCostFlow(A=A, b=b.copy(), c=c.copy(), u=u.copy()), method="cnet_mcf",
                            solver="HGS", solver_settings=_quiet())
    ref = float(fx.out["cnet_mcf_obj"])
    assert abs(out.obj_val - ref) <= 1e-9 * abs(ref) and abs(out.obj_val - float(fx.out["direct_obj"])) <= 1e-9 * abs(ref)
    assert out.iter_count == int(fx.out["cnet_mcf_iters"])
    assert np.array_equal(np.flatnonzero(out.basis.vbasis == 0), fx.out["cnet_mcf_basic"])


def test_tree_basis_with_zero_weight_tree_arcs():
    """Flow indicators that are exactly zero on arcs the spanning tree needs (an underflowed Sinkhorn plan, a
    sparse warm start).  The reference drops zero-weight tree arcs from `max_weight_spanning_tree`
    (tree_BI.py:56, kept in the public function here too) and then fails in its square solve (:74-76);
    `tree_basis_identify` works on the full device tree and returns a feasible basis."""
    from smart_crossover.formats import OptTransport
    from smart_crossover.network_methods.net_manager import OTManager
    from smart_crossover.network_methods.tree_BI import max_weight_spanning_tree, tree_basis_identify
    fx = Fixture("ot_zero_3x3")
    s, d, M, x = fx.inp["s"], fx.inp["d"], fx.inp["M"], fx.inp["x"]
    S, D = M.shape
    mgr = OTManager(OptTransport(s, d, M))
    queue, scores = mgr.get_sorted_flows(x)
    assert np.array_equal(max_weight_spanning_tree(mgr.ot, scores), fx.out["tree"]) and fx.out["tree"].size < S + D - 1
    basis, push_iter = tree_basis_identify(mgr, scores)
    basic = np.flatnonzero(basis.vbasis == 0)
    assert 1 <= basic.size <= S + D - 1 and basis.cbasis[-1] == 0
    # the basic arcs carry a feasible transport plan: A_B x_B = b has a non-negative solution
    A = np.zeros((S + D, basic.size))
    A[basic // D, np.arange(basic.size)] = -1.0
    A[S + basic % D, np.arange(basic.size)] = 1.0
    b = np.concatenate([-s, d])
    xb, res, *_ = np.linalg.lstsq(A, b, rcond=None)
    assert np.allclose(A @ xb, b, atol=1e-12) and (xb > -1e-12).all()


def test_get_sorted_flows_downloads_through_pinned_staging():
    """Above 64 MB the scores and the queue come back through the chunked pinned-staging download
    (`device.to_host`), the scores while the sort is still running: 3 200 x 3 200 = 82 MB per array, bitwise
    equal to the oracle (net_manager.py:368-379 with a stable argsort)."""
    import cases
    from oracle import network_oracle as orc
    from smart_crossover import device as dev
    from smart_crossover.formats import OptTransport
    from smart_crossover.network_methods.net_manager import OTManager
    S = D = 3200
    s, d, M = cases.ot_points(S, D, 31)
    x = cases.interior_flow(s, d, M, 31, 0.33)
    queue, scores = OTManager(OptTransport(s, d, M)).get_sorted_flows(x)
    ref = orc.ot_flow_scores(x, s, d)
    assert scores.tobytes() == ref.tobytes()
    assert queue.dtype == np.int64 and np.array_equal(queue, orc.stable_queue(ref))
    t = torch.arange(3 * (1 << 22) + 5, dtype=torch.int64, device="cuda")          # 100 MB, not a chunk multiple
    assert np.array_equal(dev.to_host(t), np.arange(3 * (1 << 22) + 5, dtype=np.int64))


def test_staged_upload_matches_plain_copy():
    from smart_crossover import device as dev
    rng = np.random.default_rng(0)
    a = rng.random(3 * (1 << 22) + 7)                                  # 100 MB, not a chunk multiple
    assert np.array_equal(dev.to_device(a).cpu().numpy(), a)
    b = rng.integers(0, 1 << 40, size=(4099, 2051))                    # 2-D int64, 67 MB
    t = dev.to_device(b)
    assert t.shape == b.shape and t.dtype == torch.int64 and np.array_equal(t.cpu().numpy(), b)
    assert dev.to_device(np.arange(10, dtype=np.int64), dtype=torch.int32).dtype == torch.int32
    M = rng.random((3001, 2800))                                       # 67 MB: CostSlabs upload through staging
    slabs = dev.CostSlabs.from_host(M, [torch.cuda.current_device()], border=(9.0, 0.0))
    got = slabs.view(0).cpu().numpy()
    assert np.array_equal(got[:3001, :2800], M) and (got[3001, :2800] == 9.0).all() and (got[:3001, 2800] == 9.0).all()
    assert got[3001, 2800] == 0.0
