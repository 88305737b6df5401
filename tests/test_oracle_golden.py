"""The CPU oracle (oracle/) against the golden vectors produced by the reference itself.

This is what pins the oracle: every fixture in tests/golden/ was
written by tests/golden/make_golden.py running the unmodified reference.
"""
import numpy as np
import pytest
import scipy.sparse as sp

import cases
from golden_util import MCF_FULL, OT_FULL, Fixture, c2_inputs, digests, mcf_mid_inputs
from oracle import network_oracle as orc

RTOL = 1e-9   # north_star: potentials / reduced costs agree to 1e-9 relative


@pytest.mark.parametrize("name", OT_FULL + ["ot_zero_3x3"])
def test_ot_scores_queue_tree_bit_exact(name):
    fx = Fixture(name)
    s, d, M, x = (fx.inp[k] for k in "sdMx")
    F = orc.ot_flow_scores(x, s, d)
    assert F.tobytes() == fx.out["scores"].tobytes()
    assert np.array_equal(orc.stable_queue(F), fx.out["queue"])
    assert np.array_equal(orc.max_weight_spanning_tree(F, *M.shape), fx.out["tree"])


def test_zero_weight_quirk_is_exercised():
    fx = Fixture("ot_zero_3x3")
    S, D = fx.inp["M"].shape
    F = orc.ot_flow_scores(fx.inp["x"], fx.inp["s"], fx.inp["d"])
    full = orc.max_weight_spanning_tree(F, S, D, drop_zero_weight=False)
    assert full.size == S + D - 1 and fx.out["tree"].size < S + D - 1


@pytest.mark.parametrize("name", OT_FULL)
def test_ot_potentials_and_reduced_costs(name):
    fx = Fixture(name)
    M = fx.inp["M"]
    y = orc.ot_tree_potentials(fx.out["tree"], M)
    scale = np.abs(M).max()
    np.testing.assert_allclose(y, fx.out["y_tree"], rtol=RTOL, atol=RTOL * scale)
    # reduced costs: bitwise for the same y
    assert orc.reduced_costs_ot(M, fx.out["y_tree"]).tobytes() == fx.out["rc_tree"].tobytes()
    assert orc.reduced_costs_ot(M, fx.out["y_pert"]).tobytes() == fx.out["rc_pert"].tobytes()
    for tag in ("tree", "pert"):
        cnt, mn, ids, vals = orc.price_summary(fx.out["rc_" + tag], K=32)
        assert (cnt == 0) == bool(fx.out["optimal_" + tag])
        cnt2, mn2, ids2, vals2 = orc.price_dense_ot_blocked(M, fx.out["y_" + tag], K=32, block_rows=7)
        assert (cnt, mn) == (cnt2, mn2) and np.array_equal(ids, ids2) and np.array_equal(vals, vals2)


@pytest.mark.parametrize("name", OT_FULL)
def test_ot_tree_flows_and_push(name):
    fx = Fixture(name)
    s, d, M = fx.inp["s"], fx.inp["d"], fx.inp["M"]
    flows = orc.ot_tree_flows(fx.out["tree"], s, d)
    np.testing.assert_allclose(flows, fx.out["tree_flows"], rtol=1e-9, atol=1e-12)
    # H7: the push phase is fp-fragile; feed it the reference's own tree flows for the exact check
    vbasis, it = orc.push_tree_to_bfs(fx.out["tree"], fx.out["tree_flows"], *M.shape)
    assert it == int(fx.out["push_iter"])
    assert np.array_equal(vbasis, fx.out["vbasis_tree"])


def _A(tail, head, N):
    E = tail.size
    return sp.csr_matrix((np.concatenate([np.ones(E), -np.ones(E)]),
                          (np.concatenate([tail, head]), np.concatenate([np.arange(E)] * 2))),
                         shape=(N, E))


@pytest.mark.parametrize("name", MCF_FULL)
def test_mcf_scores_queue_rc(name):
    fx = Fixture(name)
    tail, head, b, c, u, x = (fx.inp[k] for k in ("tail", "head", "b", "c", "u", "x"))
    A = _A(tail, head, b.size)
    t2, h2 = orc.mcf_endpoints(A)
    assert np.array_equal(t2, tail) and np.array_equal(h2, head)
    ind = orc.mcf_flow_scores(x, u, A)
    assert ind.tobytes() == fx.out["scores"].tobytes()
    assert np.array_equal(orc.stable_queue(ind), fx.out["queue"])
    rc = orc.reduced_costs_arcs(c, tail, head, fx.out["y"], fx.out["vbasis"])
    assert rc.tobytes() == fx.out["rc"].tobytes()
    assert (orc.price_summary(rc, 8)[0] == 0) == bool(fx.out["optimal"])


def test_c2_784_digests():
    s, d, M, x = c2_inputs()
    dg = digests()["ot_c2_784"]["digests"]
    small = Fixture("ot_c2_784_small").out
    F = orc.ot_flow_scores(x, s, d)
    assert cases.digest(F) == dg["scores"]
    assert cases.digest(orc.stable_queue(F)) == dg["queue"]
    tree = orc.max_weight_spanning_tree(F, *M.shape)
    assert np.array_equal(tree, small["tree"])
    y = orc.ot_tree_potentials(tree, M)
    np.testing.assert_allclose(y, small["y_tree"], rtol=RTOL, atol=RTOL * M.max())
    rc = orc.reduced_costs_ot(M, small["y_pert"])
    assert cases.digest(rc) == dg["rc_pert"]
    cnt, mn, ids, vals = orc.price_summary(rc, K=256)
    assert cnt == int(small["count_pert"]) and mn == float(small["min_rc_pert"])
    assert np.array_equal(ids, small["topk_ids_pert"])


def test_c2_784_push_phase():
    """784 x 784: the oracle's tree flows through the oracle's push loop and through the product's native
    push loop (sx_push_tree_h, host code): the reference's push count (934) and basis."""
    import ctypes
    from smart_crossover import _native
    s, d, M, x = c2_inputs()
    small = Fixture("ot_c2_784_small").out
    tree = small["tree"]
    flows = orc.ot_tree_flows(tree, s, d)
    vbasis, it = orc.push_tree_to_bfs(tree, flows, *M.shape)
    assert it == int(small["push_iter"]) == 934
    assert np.array_equal(np.flatnonzero(vbasis == 0), small["basic_tree"])
    pos = np.empty(4 * tree.size, dtype=np.int64)
    n_pos, n_it = ctypes.c_int64(0), ctypes.c_int64(0)
    tree64, flows64 = np.ascontiguousarray(tree, dtype=np.int64), np.ascontiguousarray(flows)
    assert _native.lib.sx_push_tree_h(tree64.ctypes.data, flows64.ctypes.data, tree.size, *M.shape, pos.ctypes.data,
                                      pos.size, ctypes.byref(n_pos), ctypes.byref(n_it)) == 0
    assert n_it.value == 934 and np.array_equal(np.sort(pos[:n_pos.value]), small["basic_tree"])


def test_mcf_mid_digests():
    tail, head, b, c, u, x = mcf_mid_inputs()
    dg = digests()["mcf_mid_20k"]["digests"]
    small = Fixture("mcf_mid_20k_small").out
    ind = orc.mcf_flow_scores(x, u, _A(tail, head, b.size))
    assert cases.digest(ind) == dg["scores"]
    assert cases.digest(orc.stable_queue(ind)) == dg["queue"]


def test_column_chunks_schedule():
    # algorithms.py:102: n/m > 1000 -> 10 m, else int(1.2 m); doubling (parameters.py:16)
    assert orc.column_chunks(1568, 614656, 614656, 3) == [(0, 1881), (1881, 3762), (3762, 7524)]
    assert orc.column_chunks(40000, 4 * 10**8, 4 * 10**8, 2) == [(0, 400000), (400000, 800000)]
