"""Host-side logic of the drop-in package (no GPU): problem classes, sub-problem assembly, the
push phase of the tree basis, the HiGHS backend, and the column-generation bookkeeping."""
import numpy as np
import pytest
import scipy.sparse as sp

import cases
from golden_util import OT_FULL, Fixture
from oracle import network_oracle as orc
from smart_crossover.formats import MinCostFlow, OptTransport, ot_incidence
from smart_crossover.network_methods.net_manager import MCFManagerStd, OTManager
from smart_crossover.network_methods.tree_BI import push_tree_to_bfs
from smart_crossover.output import Basis, Output
from smart_crossover.solver_caller.caller import SolverSettings
from smart_crossover.solver_caller.solving import generate_solver_caller, solve_mcf, solve_ot


def test_problem_classes_validate_like_the_reference():
    with pytest.raises(ValueError):
        OptTransport(np.array([0.5, 0.5]), np.array([0.5, 0.4]), np.zeros((2, 2)))       # formats.py:144-145
    A = sp.csr_matrix(np.array([[1.0, -1.0], [-1.0, 1.0]]))
    with pytest.raises(ValueError):
        MinCostFlow(A, np.array([1.0, 0.0]), np.ones(2), np.ones(2))                      # formats.py:120-121
    mcf = MinCostFlow(A.tocsc(), np.array([1.0, -1.0]), np.ones(2), np.ones(2))
    assert sp.isspmatrix_csr(mcf.A) and np.array_equal(mcf.l, np.zeros(2)) and mcf.name == "mcf_instance"
    b = Basis(np.array([0.0, -1.0]), np.array([-1.0]))
    assert b.vbasis.dtype.kind == "i" and b.cbasis.dtype.kind == "i"                      # output.py:15-17


def test_to_mcf_matches_the_reference_layout():
    S, D = 3, 4
    rng = np.random.default_rng(0)
    s = np.full(S, 1 / S); d = np.full(D, 1 / D); M = rng.random((S, D))
    mcf = OptTransport(s, d, M).to_MCF()
    # reference formats.py:155-160: A = [-kron(I_S, 1_D^T); kron(1_S^T, I_D)], b = [-s, d], c = M.flatten()
    ref = sp.vstack([-sp.kron(np.eye(S), np.ones((1, D))), sp.kron(np.ones((1, S)), np.eye(D))]).toarray()
    assert np.array_equal(mcf.A.toarray(), ref)
    assert np.array_equal(mcf.b, np.hstack([-s, d])) and np.array_equal(mcf.c, M.flatten())
    assert np.all(np.isinf(mcf.u)) and np.array_equal(ot_incidence(S, D).toarray(), ref)


@pytest.mark.parametrize("name", OT_FULL)
def test_push_tree_to_bfs_matches_golden(name):
    fx = Fixture(name)
    mgr = OTManager(OptTransport(fx.inp["s"], fx.inp["d"], fx.inp["M"]))
    # the host push loop, fed the reference's own tree flows (the device computes them in production:
    # tests/test_gpu_parity.py::test_tree_flows_golden and test_gpu_e2e.py run that path)
    vbasis, push_iter = push_tree_to_bfs(mgr, fx.out["tree"], _flows=fx.out["tree_flows"])
    assert push_iter == int(fx.out["push_iter"])
    assert np.array_equal(vbasis.astype(np.int64), fx.out["vbasis_tree"])


def test_native_push_loop_matches_the_oracle_on_random_instances():
    """sx_push_tree_h (host C++, sparse support) against the oracle's dense restatement of tree_BI.py:77-114 on
    random shapes: same basis, same push count; the capacity-retry protocol of the C entry point."""
    import ctypes
    from smart_crossover import _native
    rng = np.random.default_rng(17)
    total_pushes = 0
    for trial in range(40):
        S, D = int(rng.integers(2, 40)), int(rng.integers(2, 40))
        s, d, M = cases.ot_points(S, D, 500 + trial)
        x = cases.interior_flow(s, d, M, 500 + trial, 0.33)
        F = orc.ot_flow_scores(x, s, d)
        tree = orc.max_weight_spanning_tree(F, S, D)
        if tree.size != S + D - 1:
            continue
        flows = orc.ot_tree_flows(tree, s, d)
        mgr = OTManager(OptTransport(s, d, M))
        try:
            vb_ref, it_ref = orc.push_tree_to_bfs(tree, flows, S, D)
        except AssertionError:
            with pytest.raises(AssertionError):
                push_tree_to_bfs(mgr, tree, _flows=flows)
            continue
        vb, it = push_tree_to_bfs(mgr, tree, _flows=flows)
        assert it == it_ref and np.array_equal(vb.astype(np.int64), vb_ref)
        total_pushes += it
        # capacity protocol: too small a buffer reports the size needed, the second call fills it
        n_pos, n_it = ctypes.c_int64(0), ctypes.c_int64(0)
        tr, fl = np.ascontiguousarray(tree, np.int64), np.ascontiguousarray(flows, np.float64)
        small = np.empty(1, dtype=np.int64)
        rc = _native.lib.sx_push_tree_h(tr.ctypes.data, fl.ctypes.data, tr.size, S, D, small.ctypes.data, 1,
                                        ctypes.byref(n_pos), ctypes.byref(n_it))
        assert rc == _native.SX_ERR_WORKSPACE and n_pos.value == int((vb_ref == 0).sum()) and n_it.value == it_ref
        buf = np.empty(n_pos.value, dtype=np.int64)
        rc = _native.lib.sx_push_tree_h(tr.ctypes.data, fl.ctypes.data, tr.size, S, D, buf.ctypes.data, buf.size,
                                        ctypes.byref(n_pos), ctypes.byref(n_it))
        assert rc == 0 and np.array_equal(np.sort(buf), np.flatnonzero(vb_ref == 0))
    assert total_pushes > 50                                        # the loop was really exercised
    # argument validation
    n_pos, n_it = ctypes.c_int64(0), ctypes.c_int64(0)
    bad = np.array([99], dtype=np.int64)
    one = np.array([1.0])
    assert _native.lib.sx_push_tree_h(bad.ctypes.data, one.ctypes.data, 1, 3, 3, bad.ctypes.data, 1,
                                      ctypes.byref(n_pos), ctypes.byref(n_it)) == -1


@pytest.mark.parametrize("name", OT_FULL)
def test_sub_problem_assembly_equals_slicing_the_full_matrix(name):
    fx = Fixture(name)
    ot = OptTransport(fx.inp["s"], fx.inp["d"], fx.inp["M"])
    mgr = OTManager(ot)
    mgr.add_free_variables(fx.out["queue"][:50])
    mgr.add_free_variables(fx.out["vbasis_tree"] == 0)          # boolean mask, as the tnet driver passes
    sub = mgr.get_sub_problem()
    mask = mgr.mask_sub_ot.ravel()
    full = ot.to_MCF()
    assert np.array_equal(sub.A.toarray(), full.A.tocsc()[:, mask].toarray())     # net_manager.py:452
    assert np.array_equal(sub.c, full.c[mask]) and np.array_equal(sub.b, full.b)
    x_sub = np.arange(mask.sum(), dtype=float)
    assert np.array_equal(mgr.recover_x_from_sub_x(x_sub)[mask], x_sub)
    # big-M extension keeps the reference's host-side layout
    mgr2 = OTManager(OptTransport(fx.inp["s"], fx.inp["d"], fx.inp["M"]))
    mgr2._device_cost = lambda: None                              # no GPU in this test
    S, D = fx.inp["M"].shape
    mgr2.extend_by_bigM(7.0)
    assert mgr2.ot.M.shape == (S + 1, D + 1) and mgr2.ot.M[S, D] == 0 and mgr2.ot.M[0, D] == 7.0
    assert mgr2.artificial_vars.size == S + D + 1 and mgr2.mask_sub_ot.sum() == S + D + 1
    mgr2.add_free_variables(np.array([0, D + 1]))                 # original ids -> (0,0) and (1,1)
    assert mgr2.mask_sub_ot[0, 0] and mgr2.mask_sub_ot[1, 1]
    mgr2.set_initial_basis()
    assert mgr2.basis.cbasis.size == S + D + 2 and (mgr2.basis.vbasis[mgr2.artificial_vars] == 0).all()


@pytest.mark.parametrize("name", ["mcf_small_200", "mcf_ties_60"])
def test_mcf_restricted_master_equals_the_reference_formulas(name):
    """`MCFManagerStd.update_subproblem` / `add_free_variables` / `fix_variables` build the restricted master
    from the arcs' end points and a mask of freed arcs; the reference slices `A[:, non_fix]`, multiplies
    `A[:, fix_up] @ u[fix_up]` and calls `np.setdiff1d` three times per round (net_manager.py:202-209,
    224-245).  Same matrices (canonical CSC arrays), bit-identical right-hand side, same index sets."""
    import scipy.sparse as sp
    from smart_crossover.formats import MinCostFlow
    from smart_crossover.network_methods.net_manager import MCFManagerStd
    fx = Fixture(name)
    tail, head, b, c, u, x = (fx.inp[k] for k in ("tail", "head", "b", "c", "u", "x"))
    E, N = c.size, b.size
    A = sp.lil_matrix((N, E), dtype=int)
    A[head, np.arange(E)] = -1
    A[tail, np.arange(E)] = 1
    mgr = MCFManagerStd(MinCostFlow(A=A.tocsr(), b=b.copy(), c=c.copy(), u=u * 1.37))      # non-integer capacities
    x = np.clip(x, 0, mgr.mcf.u)
    mgr.rescale_cost(np.max(np.abs(c)))
    low, up = np.where(x < mgr.mcf.u / 2)[0], np.where(x >= mgr.mcf.u / 2)[0]
    mgr.fix_variables(ind_fix_to_low=low, ind_fix_to_up=up)
    ref = {"fix_low": low, "fix_up": up, "non_fix": np.setdiff1d(range(E), np.append(low, up))}
    ref["fix"] = np.setdiff1d(range(E), ref["non_fix"])
    mgr.extend_by_bigM(mgr.m * np.max(mgr.mcf.c))
    ref["non_fix"] = np.append(ref["non_fix"], np.arange(E, E + N))
    queue = np.random.default_rng(0).permutation(E)
    for lo, hi in ((0, 50), (50, 150), (150, 400)):
        mgr.add_free_variables(queue[lo:hi])
        ref["non_fix"] = np.append(ref["non_fix"], queue[lo:hi])
        for key in ("fix", "fix_low", "fix_up"):
            ref[key] = np.setdiff1d(ref[key], queue[lo:hi])
        for key in ref:
            assert np.array_equal(mgr.var_info[key], ref[key]), key
        mgr.update_subproblem()
        A_ref = sp.csc_matrix(mgr.mcf.A[:, ref["non_fix"]])
        b_ref = mgr.mcf.b - mgr.mcf.A[:, ref["fix_up"]] @ mgr.mcf.u[ref["fix_up"]]
        A_sub = sp.csc_matrix(mgr.mcf_sub.A)
        assert np.array_equal(A_sub.indptr, A_ref.indptr) and np.array_equal(A_sub.indices, A_ref.indices)
        assert np.array_equal(A_sub.data, A_ref.data)
        assert mgr.mcf_sub.b.tobytes() == b_ref.tobytes()
        assert np.array_equal(mgr.mcf_sub.c, mgr.mcf.c[ref["non_fix"]])
        assert np.array_equal(mgr.mcf_sub.u, mgr.mcf.u[ref["non_fix"]])


def test_bigm_extended_cost_view_equals_the_reference_extension():
    """`extend_by_bigM` keeps `ot.M` as a view (no (S+1) x (D+1) host copy): shape, single entries, flat takes
    and the materialised array equal the reference's `np.vstack([np.hstack(...)])` (net_manager.py:390-393)."""
    from smart_crossover.network_methods.net_manager import BigMExtendedCost, _take_flat
    rng = np.random.default_rng(4)
    M = rng.random((7, 5))
    bigM = 123.5
    ref = np.vstack([np.hstack([M, bigM * np.ones((7, 1))]), np.hstack([bigM * np.ones((1, 5)), np.array([[0]])])])
    view = BigMExtendedCost(M, bigM)
    assert view.shape == ref.shape and view.size == ref.size and view.ndim == 2
    assert np.array_equal(np.asarray(view), ref) and np.array_equal(view.ravel(), ref.ravel())
    ids = rng.permutation(ref.size)
    assert np.array_equal(view.take_flat(ids), ref.ravel()[ids]) and np.array_equal(_take_flat(view, ids), ref.ravel()[ids])
    assert view[7, 5] == 0.0 and view[0, 5] == bigM and view[7, 0] == bigM and view[3, 2] == M[3, 2] and view[-1, -1] == 0.0
    assert view.max() == ref.max() and np.array_equal(view[2:4], ref[2:4])
    assert np.array_equal(_take_flat(M, np.array([0, 6, 34])), M.ravel()[[0, 6, 34]])


def test_highs_backend_solves_and_reports_like_a_solver_caller():
    fx = Fixture("ot_c1_40x40")
    ot = OptTransport(fx.inp["s"], fx.inp["d"], fx.inp["M"])
    out = solve_ot(ot, solver="HGS", settings=SolverSettings(log_console=0))
    assert out.status == "OPTIMAL" and isinstance(out, Output)
    assert abs(out.obj_val - float(fx.out["tnet_obj"])) <= 1e-9 * abs(out.obj_val)
    mcf = ot.to_MCF()
    rc = mcf.c - mcf.A.T @ out.y                                   # dual sign = Gurobi Pi (net_manager.py:483)
    assert rc.min() >= -1e-6
    assert (out.basis.vbasis == 0).sum() + (out.basis.cbasis == 0).sum() == mcf.b.size
    warm = solve_mcf(mcf, solver="HGS", warm_start_basis=out.basis, settings=SolverSettings(log_console=0))
    assert warm.iter_count == 0 and abs(warm.obj_val - out.obj_val) < 1e-12
    with pytest.raises(ImportError):
        generate_solver_caller("GRB")
    with pytest.raises(ValueError):
        generate_solver_caller("XYZ")


def test_mcf_manager_bookkeeping():
    fx = Fixture("mcf_small_200")
    tail, head, b, c, u, x = (fx.inp[k] for k in ("tail", "head", "b", "c", "u", "x"))
    E, N = c.size, b.size
    A = sp.csr_matrix((np.concatenate([np.ones(E), -np.ones(E)]),
                       (np.concatenate([tail, head]), np.concatenate([np.arange(E)] * 2))), shape=(N, E))
    mcf = MinCostFlow(A, b.copy(), c.copy(), u.copy())
    mgr = MCFManagerStd(mcf)
    t2, h2 = mgr._endpoints(A)
    assert np.array_equal(t2, tail) and np.array_equal(h2, head)
    mgr.rescale_cost(c.max())
    assert mcf.c.max() == 1.0 and mgr.recover_obj_val(2.0) == 2.0 * c.max()        # caller's object is rebound
    mgr.fix_variables(ind_fix_to_up=np.where(x >= u / 2)[0], ind_fix_to_low=np.where(x < u / 2)[0])
    assert mgr.var_info["non_fix"].size == 0 and mgr.var_info["fix"].size == E
    mgr.extend_by_bigM(N * 1.0)
    assert mgr.mcf.A.shape == (N + 1, E + N) and mgr.artificial_vars.size == N
    assert np.allclose(np.asarray(mgr.mcf.A.sum(axis=0)).ravel(), 0)               # every column: one +1, one -1
    t3, h3 = mgr._endpoints(mgr.mcf.A)
    assert (t3 >= 0).all() and (h3 >= 0).all()
    mgr.update_subproblem()
    mgr.set_initial_basis()
    assert mgr.mcf_sub.c.size == N and (mgr.basis.vbasis[mgr.var_info["fix_up"]] == -2).all()
    mgr.add_free_variables(fx.out["queue"][:10])
    assert np.array_equal(mgr.var_info["non_fix"][-10:], fx.out["queue"][:10])
    assert not np.isin(fx.out["queue"][:10], mgr.var_info["fix"]).any()


def test_exp_nonpos_of_the_warm_start_is_within_two_ulp_of_libm(tmp_path):
    """csrc/sx_expm.cuh is host-compilable: the exponential the Sinkhorn passes use (x <= 0, table + degree-6
    Taylor) stays within 2 ulp of libm on (-700, 0], is exactly 1 at 0, 0 from -700 down, and passes NaN."""
    import ctypes
    import pathlib
    import shutil
    import subprocess
    if shutil.which("g++") is None:
        pytest.skip("no host compiler")
    hdr = pathlib.Path(__file__).resolve().parents[1] / "smart-crossover_b200" / "csrc" / "sx_expm.cuh"
    src = tmp_path / "expm_shim.cpp"
    src.write_text(f'#include "{hdr}"\n'
                   "static const double tab[sx::kExpTabSize] = {SX_EXP_TAB_VALUES};\n"
                   'extern "C" void expm_eval(const double *x, long n, double *out) {\n'
                   "    for (long i = 0; i < n; ++i) out[i] = sx::exp_nonpos(x[i], tab);\n}\n")
    so = tmp_path / "expm_shim.so"
    subprocess.run(["g++", "-O2", "-mfma", "-x", "c++", "-shared", "-fPIC", "-o", str(so), str(src)], check=True)
    fn = ctypes.CDLL(str(so)).expm_eval
    fn.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p]
    rng = np.random.default_rng(5)
    x = np.concatenate([-700 * rng.random(400000), -5 * rng.random(400000), -60 * rng.random(200000) ** 3,
                        -np.ldexp(1.0, -np.arange(1, 60)), -np.log(2) / 32 * np.arange(0, 20000)])
    x = np.ascontiguousarray(x[x > -700.0])
    out = np.empty_like(x)
    fn(x.ctypes.data, x.size, out.ctypes.data)
    ref = np.exp(x)
    assert np.max(np.abs(out - ref) / ref) < 2 * np.finfo(np.float64).eps
    edge = np.array([0.0, -0.0, -700.0, -745.2, -1e9, -np.inf, np.nan])
    got = np.empty_like(edge)
    fn(edge.ctypes.data, edge.size, got.ctypes.data)
    assert got[0] == 1.0 and got[1] == 1.0 and np.all(got[2:6] == 0.0) and np.isnan(got[6])
