"""CPU oracle for the network-crossover hot path (NumPy + a small C helper).

TEST INFRASTRUCTURE ONLY.  This module restates, on the CPU, what the
reference (`/root/reference/src/smart_crossover`, cited as file:line below)
computes on the path  flow -> sorted arcs -> spanning-tree basis -> potentials
-> pricing.  It exists so the CUDA path can be checked against it.  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import it; the product package
(`smart-crossover_b200/`) never does and has no CPU fallback.

Pinning: the reference ships no tests or golden vectors,
so this oracle is pinned against outputs of the reference itself, produced in
the authoring container by `tests/golden/make_golden.py` (reference imported
through `tests/golden/_ref_shim.py`) and committed under `tests/golden/`.
`tests/test_oracle_golden.py` replays every fixture.

Conventions:
  OT arc id      k = i * D + j                    (net_manager.py:366,379)
  OT node ids    sources 0..S-1, sinks S..S+D-1   (formats.py:156-160)
  OT incidence   A[i, k] = -1, A[S+j, k] = +1     (formats.py:156-158)
  MCF incidence  A[tail, k] = +1, A[head, k] = -1 (scripts/min2mcf.py:36-37)
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

TOLERANCE_FOR_REDUCED_COSTS = 1e-6   # parameters.py:8
TOLERANCE_FOR_ARTIFICIAL_VARS = 1e-8  # parameters.py:7

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build_c_helper(force: bool = False) -> str:
    """Compile oracle/sx_oracle.c with gcc into oracle/_build/libsxoracle.so."""
    out_dir = os.path.join(_HERE, "_build")
    so = os.path.join(out_dir, "libsxoracle.so")
    src = os.path.join(_HERE, "sx_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        os.makedirs(out_dir, exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off",
                               "-o", so, src, "-lm"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(build_c_helper())
        i64p = ctypes.POINTER(ctypes.c_int64)
        f64p = ctypes.POINTER(ctypes.c_double)
        lib.sxo_kruskal.restype = ctypes.c_int64
        lib.sxo_kruskal.argtypes = [i64p, ctypes.c_int64, i64p, i64p, ctypes.c_int64,
                                    ctypes.c_int64, ctypes.c_int64, i64p]
        lib.sxo_tree_potentials.restype = ctypes.c_int
        lib.sxo_tree_potentials.argtypes = [i64p, i64p, f64p, ctypes.c_int64, ctypes.c_int64,
                                            ctypes.c_int64, f64p]
        lib.sxo_mcf_node_sums.restype = None
        lib.sxo_mcf_node_sums.argtypes = [i64p, i64p, f64p, ctypes.POINTER(ctypes.c_ubyte), f64p,
                                          ctypes.c_int64, f64p, f64p]
        lib.sxo_push_tree.restype = ctypes.c_int64
        lib.sxo_push_tree.argtypes = [f64p, ctypes.c_int64, ctypes.c_int64]
        _LIB = lib
    return _LIB


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct)) if a is not None else None


# --------------------------------------------------------------------------
# K1: flow indicators and the sorted queue
# --------------------------------------------------------------------------
def ot_flow_scores(x: np.ndarray, s: np.ndarray, d: np.ndarray) -> np.ndarray:
    """F[k] = max(x_ij / s_i, x_ij / d_j), flattened row-major.  net_manager.py:377-379."""
    S, D = s.size, d.size
    X = np.asarray(x, dtype=np.float64).reshape(S, D)
    with np.errstate(divide="ignore", invalid="ignore"):
        F = np.maximum(X / s.reshape(S, 1), X / d.reshape(1, D))
    return F.ravel()


def stable_queue(scores: np.ndarray) -> np.ndarray:
    """`np.argsort(scores)[::-1]` run with a stable sort: descending score, ties by
    DESCENDING arc id.  net_manager.py:184,379 (argsort made stable per north_star)."""
    return np.argsort(scores, kind="stable")[::-1].astype(np.int64)


def kruskal_order(scores: np.ndarray) -> np.ndarray:
    """Order in which SciPy's Kruskal visits the arcs for tree_BI.py:47,53: a stable
    argsort of the negated weights = descending score, ties by ASCENDING arc id."""
    return np.argsort(-np.asarray(scores, dtype=np.float64), kind="stable").astype(np.int64)


def mcf_flow_scores(x, u, A_csr) -> np.ndarray:
    """Flow indicators of `MCFManagerStd.get_sorted_flows`, net_manager.py:165-182.

    x_hat: arcs with x > u/2 are reversed (x_hat = u - x, incidence column negated),
    out-of-bounds flows are zeroed (:166-169); per-node out/in sums f1, f2 by
    sequential CSR row sums (:171-175); f = max(f1, f2), f_inv = 1/f or 0 (:176-177);
    indicator[k] = max over the two end nodes of |f_inv[node] * x_hat[k]| (:178-182).
    """
    x = np.asarray(x, dtype=np.float64)
    u = np.asarray(u, dtype=np.float64)
    A = A_csr.tocsr()
    if not A.has_sorted_indices:
        A = A.sorted_indices()
    N, E = A.shape
    mask = x > u / 2
    x_hat = x * (~mask) + u * mask - x * mask
    x_hat[(x < 0) | (x > u)] = 0
    f1 = np.empty(N)
    f2 = np.empty(N)
    ptr = A.indptr.astype(np.int64)
    idx = A.indices.astype(np.int64)
    val = A.data.astype(np.float64)
    flip = mask.astype(np.uint8)
    _lib().sxo_mcf_node_sums(_p(ptr, ctypes.c_int64), _p(idx, ctypes.c_int64),
                             _p(val, ctypes.c_double), _p(flip, ctypes.c_ubyte),
                             _p(x_hat, ctypes.c_double), N, _p(f1, ctypes.c_double),
                             _p(f2, ctypes.c_double))
    f = np.maximum(f1, f2)
    f_inv = np.divide(1, f, out=np.zeros_like(f), where=f != 0)
    # per-arc max over its nonzeros of |f_inv[row] * x_hat[col] * a|
    rows = np.repeat(np.arange(N, dtype=np.int64), np.diff(ptr))
    contrib = np.abs(f_inv[rows] * x_hat[idx] * np.where(flip[idx] == 1, -val, val))
    ind = np.zeros(E)
    np.maximum.at(ind, idx, contrib)
    return ind


def mcf_endpoints(A_csr):
    """tail (+1 row) and head (-1 row) of every column of A.  scripts/min2mcf.py:36-37."""
    A = A_csr.tocsc()
    E = A.shape[1]
    tail = np.full(E, -1, dtype=np.int64)
    head = np.full(E, -1, dtype=np.int64)
    cols = np.repeat(np.arange(E, dtype=np.int64), np.diff(A.indptr))
    pos = A.data > 0
    neg = A.data < 0
    tail[cols[pos]] = A.indices[pos]
    head[cols[neg]] = A.indices[neg]
    return tail, head


# --------------------------------------------------------------------------
# K2: spanning-tree basis identification
# --------------------------------------------------------------------------
def spanning_forest(order, N, S=0, D=0, tail=None, head=None) -> np.ndarray:
    """Kruskal over `order`; returns the kept arc ids sorted ascending (see sx_oracle.c)."""
    order = np.ascontiguousarray(order, dtype=np.int64)
    out = np.empty(max(N - 1, 1), dtype=np.int64)
    t = None if tail is None else np.ascontiguousarray(tail, dtype=np.int64)
    h = None if head is None else np.ascontiguousarray(head, dtype=np.int64)
    cnt = _lib().sxo_kruskal(_p(order, ctypes.c_int64), order.size, _p(t, ctypes.c_int64),
                             _p(h, ctypes.c_int64), S, D, N, _p(out, ctypes.c_int64))
    return np.sort(out[:cnt])


def max_weight_spanning_tree(scores: np.ndarray, S: int, D: int, drop_zero_weight=True) -> np.ndarray:
    """tree_BI.py:32-59.  Tree arc ids ascending.  `drop_zero_weight` reproduces the
    reference quirk that tree arcs of weight exactly 0 vanish from the result
    (`np.flatnonzero(min_tree)` at tree_BI.py:56 on a dense matrix whose tree entries
    are the weights themselves; SURVEY.md hard part H2)."""
    tree = spanning_forest(kruskal_order(scores), S + D, S=S, D=D)
    if drop_zero_weight:
        tree = tree[np.asarray(scores)[tree] != 0]
    return tree


# --------------------------------------------------------------------------
# K3: node potentials from the tree
# --------------------------------------------------------------------------
def tree_potentials(plus, minus, cost, N, root) -> np.ndarray:
    """y with y[plus_t] - y[minus_t] = cost_t for every tree arc and y[root] = 0.

    SURVEY.md section 8 row a5 / oracle definition (3):  B = A[:-1, tree]
    (tree_BI.py:74),  B^T y[:-1] = c[tree],  y[m-1] = 0.  Raises ValueError if the
    arcs are not a spanning tree."""
    plus = np.ascontiguousarray(plus, dtype=np.int64)
    minus = np.ascontiguousarray(minus, dtype=np.int64)
    cost = np.ascontiguousarray(cost, dtype=np.float64)
    y = np.zeros(N)
    rc = _lib().sxo_tree_potentials(_p(plus, ctypes.c_int64), _p(minus, ctypes.c_int64),
                                    _p(cost, ctypes.c_double), plus.size, N, root,
                                    _p(y, ctypes.c_double))
    if rc != 0:
        raise ValueError("arcs do not form a spanning tree")
    return y


def ot_tree_potentials(tree, M, root=None) -> np.ndarray:
    """Potentials for an OT tree: plus = S + j, minus = i, root = last node (tree_BI.py:28)."""
    S, D = M.shape
    tree = np.asarray(tree, dtype=np.int64)
    i, j = tree // D, tree % D
    return tree_potentials(S + j, i, M[i, j], S + D, S + D - 1 if root is None else root)


# --------------------------------------------------------------------------
# K4: pricing
# --------------------------------------------------------------------------
def reduced_costs_ot(M: np.ndarray, y: np.ndarray) -> np.ndarray:
    """rc = c - A^T y for the OT incidence = fl(M_ij - fl(y_{S+j} - y_i)).
    net_manager.py:483 (association fixed by SciPy's csc_matvec: SURVEY.md H4)."""
    S, D = M.shape
    return (M - (y[S:S + D][None, :] - y[:S][:, None])).ravel()


def reduced_costs_arcs(c, tail, head, y, vbasis=None) -> np.ndarray:
    """rc = c - (y_tail - y_head), negated where vbasis == -2.  net_manager.py:302-303."""
    rc = c - (y[tail] - y[head])
    if vbasis is not None:
        flip = np.asarray(vbasis) == -2
        rc[flip] = -rc[flip]
    return rc


def price_summary(rc: np.ndarray, K: int, tol: float = TOLERANCE_FOR_REDUCED_COSTS):
    """Violator count, min rc, and the top-K most violating arcs (rc ascending, ties by
    ascending arc id) among rc < -tol.  SURVEY.md section 8 row a9; the optimality flag
    of net_manager.py:318,496 is `count == 0`."""
    viol = np.flatnonzero(rc < -tol)
    count = int(viol.size)
    minrc = float(rc.min()) if rc.size else np.inf
    o = viol[np.argsort(rc[viol], kind="stable")][:K]
    return count, minrc, o.astype(np.int64), rc[o]


def price_dense_ot_blocked(M, y, K, tol=TOLERANCE_FOR_REDUCED_COSTS, threads=1, block_rows=256):
    """Row-blocked (optionally multi-threaded) pricing pass used as the CPU baseline.
    Same arithmetic as `reduced_costs_ot` + `price_summary` without materialising rc."""
    S, D = M.shape
    u = y[:S]
    v = y[S:S + D]

    def work(r0):
        r1 = min(S, r0 + block_rows)
        rc = M[r0:r1] - (v[None, :] - u[r0:r1, None])
        flat = rc.ravel()
        viol = np.flatnonzero(flat < -tol)
        return r0, viol.size, (flat.min() if flat.size else np.inf), viol + r0 * D, flat[viol]

    starts = range(0, S, block_rows)
    if threads > 1:
        with ThreadPoolExecutor(threads) as ex:
            parts = list(ex.map(work, starts))
    else:
        parts = [work(r) for r in starts]
    count = sum(p[1] for p in parts)
    minrc = min((p[2] for p in parts), default=np.inf)
    ids = np.concatenate([p[3] for p in parts]) if parts else np.zeros(0, np.int64)
    vals = np.concatenate([p[4] for p in parts]) if parts else np.zeros(0)
    o = np.argsort(vals, kind="stable")[:K]
    return int(count), float(minrc), ids[o].astype(np.int64), vals[o]


# --------------------------------------------------------------------------
# a4: tree primal flows and the push to a basic feasible solution
# --------------------------------------------------------------------------
def ot_tree_flows(tree, s, d) -> np.ndarray:
    """Solve B x_T = b[:-1], B = A[:-1, tree], b = [-s, d]  (tree_BI.py:74-76) by leaf
    elimination on the tree (exact same linear system; rounding differs from SuperLU,
    SURVEY.md H7)."""
    S, D = s.size, d.size
    N = S + D
    tree = np.asarray(tree, dtype=np.int64)
    i, j = tree // D, S + tree % D
    b = np.concatenate([-np.asarray(s, float), np.asarray(d, float)])
    deg = np.zeros(N, dtype=np.int64)
    np.add.at(deg, i, 1)
    np.add.at(deg, j, 1)
    adj = [[] for _ in range(N)]
    for t in range(tree.size):
        adj[i[t]].append(t)
        adj[j[t]].append(t)
    resid = b.copy()
    flow = np.zeros(tree.size)
    done = np.zeros(tree.size, dtype=bool)
    root = N - 1
    stack = [v for v in range(N) if deg[v] == 1 and v != root]
    while stack:
        v = stack.pop()
        if deg[v] != 1 or v == root:
            continue
        t = next(t for t in adj[v] if not done[t])
        done[t] = True
        # row v of A x = b:  (+1 if v is the sink end, -1 if the source end) * x_t = resid[v]
        xt = resid[v] if v == j[t] else -resid[v]
        flow[t] = xt
        w = i[t] if v == j[t] else j[t]
        resid[w] -= xt if w == j[t] else -xt
        deg[v] -= 1
        deg[w] -= 1
        if deg[w] == 1 and w != root:
            stack.append(w)
    return flow


def push_tree_to_bfs(tree, tree_flows, S, D):
    """tree_BI.py:77-114 given the tree primal flows: returns (vbasis, push_iter)."""
    dense = np.zeros(S * D)
    dense[np.asarray(tree, dtype=np.int64)] = tree_flows
    it = _lib().sxo_push_tree(_p(dense, ctypes.c_double), S, D)
    if it < 0:
        raise AssertionError("push_tree_to_bfs: reference assert would fire (tree_BI.py:93-94)")
    vbasis = -np.ones(S * D, dtype=np.int64)
    vbasis[dense > 0] = 0
    return vbasis, int(it)


# --------------------------------------------------------------------------
# a8: column-generation chunk schedule
# --------------------------------------------------------------------------
def column_chunks(m: int, n: int, queue_len: int, rounds: int):
    """(left, right) pointers of the first `rounds` CG iterations.  algorithms.py:101-102,
    114,135-136 with COLUMN_GENERATION_RATIO = 2 (parameters.py:16)."""
    num = int(10 * m) if n / m > 1000 else int(1.2 * m)
    left = 0
    out = []
    for _ in range(rounds):
        if left >= queue_len:
            break
        right = min(num, queue_len)
        out.append((left, right))
        num = int(2 * num)
        left = right
    return out


def sinkhorn_knopp(a, b, M, reg, num_iter_max=1000, stop_thr=1e-9, check_every=10):
    """Sinkhorn-Knopp as POT runs it for `ot.sinkhorn(a, b, M, reg, numItermax)` -- the warm start of the
    reference's OT experiments (scripts/run_network_crossover.py:96; POT is a third-party package, not
    in the reference tree: its published algorithm restated): K = exp(-M / reg), u = 1 / S, then
    v = b / (K^T u), u = a / (K v); every `check_every` iterations the column-marginal error
    || v * (K^T u) - b ||_2 is compared with stop_thr.  Plain (not log-domain) arithmetic, so only for
    reg large enough that K does not underflow.  Returns (plan, iterations, last error)."""
    a, b, M = (np.asarray(t, dtype=np.float64) for t in (a, b, M))
    K = np.exp(-M / reg)
    u = np.full(a.size, 1.0 / a.size)
    v = np.full(b.size, 1.0 / b.size)
    err, it = np.inf, 0
    for it in range(1, num_iter_max + 1):
        v = b / (K.T @ u)
        u = a / (K @ v)
        if stop_thr > 0 and (it - 1) % check_every == 0:
            err = float(np.linalg.norm(v * (K.T @ u) - b))
            if err < stop_thr:
                break
    return u[:, None] * K * v[None, :], it, err
