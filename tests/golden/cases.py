"""Deterministic synthetic instances shared by the golden generator, the tests and bench.py.

Everything here uses only IEEE-exact operations (+ - * / on float64, integer
arithmetic, PCG64 streams) so that the same seed gives bit-identical inputs in
the authoring container and on the GPU box: no exp/log, whose NumPy SIMD
kernels differ by an ulp between CPU generations.  `digest()` hashes the
result; the golden fixtures store the digest of the inputs they were generated
from and the tests refuse to compare if it differs.

Shapes follow BASELINE.json `configs` / SURVEY.md section 8d.
"""
import hashlib

import numpy as np


def digest(*arrays) -> str:
    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode())
        h.update(str(a.shape).encode())
        h.update(a.tobytes())
    return h.hexdigest()


def _marginal(rng, n):
    """Strictly positive, generic marginal with an exactly-representable integer total."""
    w = rng.integers(1000, 11000, size=n).astype(np.float64)
    return w / w.sum()


def ot_points(S, D, seed):
    """Squared-Euclidean cost between uniform points in the unit square (C1/C4/C5 geometry)."""
    rng = np.random.default_rng(seed)
    P = rng.random((S, 2))
    Q = rng.random((D, 2))
    dx = P[:, 0:1] - Q[None, :, 0]
    dy = P[:, 1:2] - Q[None, :, 1]
    M = dx * dx + dy * dy
    return _marginal(rng, S), _marginal(rng, D), M


def ot_grid(side, seed):
    """MNIST-shaped instance (C2): side x side pixel grid on both sides, integer-valued
    squared-Euclidean cost => heavy ties.  side = 28 gives 784 x 784."""
    rng = np.random.default_rng(seed)
    yy, xx = np.divmod(np.arange(side * side), side)
    dy = (yy[:, None] - yy[None, :]).astype(np.float64)
    dx = (xx[:, None] - xx[None, :]).astype(np.float64)
    M = dy * dy + dx * dx
    return _marginal(rng, side * side), _marginal(rng, side * side), M


def interior_flow(s, d, M, seed, scale, sharp=40.0):
    """Strictly positive 'interior' flow concentrated on cheap arcs (rational kernel
    instead of the Gibbs kernel exp(-M/tau), see module docstring).  `scale` is the
    typical cost (passed in, not reduced from M, to stay order-independent).  Exact
    feasibility is not required: the reference accepts any "interior-point / inaccurate
    solution" (algorithms.py:28)."""
    rng = np.random.default_rng(seed + 7919)
    t = 1.0 + sharp * (M / scale)
    kern = 1.0 / (t * t * t * t)
    X = (s[:, None] * d[None, :]) * kern * (1.0 + 1e-3 * rng.random(M.shape))
    return X.ravel()


def product_flow(s, d):
    """Tie-heavy flow X = s d^T: every entry of X/s is constant down a column."""
    return (s[:, None] * d[None, :]).ravel()


def planted_duals(M, seed, noise):
    """Potentials for pricing runs.  With y_i = a_i (sources) and y_{S+j} = b_j,
    b_j = min_i (M_ij + a_i), the reduced cost rc_ij = M_ij - (y_{S+j} - y_i) is >= 0 and
    tight once per column (dual feasible).  `noise` > 0 perturbs y so that a controlled
    fraction of arcs violates rc >= -tol."""
    S, D = M.shape
    rng = np.random.default_rng(seed + 104729)
    a = rng.random(S)
    b = (M + a[:, None]).min(axis=0)
    y = np.concatenate([a, b])
    if noise > 0:
        y = y + noise * (rng.random(S + D) - 0.5)
    return y


def netgen_like(N, E, seed, n_supply=None):
    """NETGEN-style min-cost flow (C3 shape) in min2mcf sign convention.

    Spanning ring + random arcs, integer costs in [1, 1e4], integer capacities in
    [1, 1e3] (finite), balanced integer supplies.  Returns (tail, head, b, c, u)."""
    rng = np.random.default_rng(seed)
    ring_t = np.arange(N, dtype=np.int64)
    ring_h = (ring_t + 1) % N
    extra = E - N
    t = rng.integers(0, N, size=extra)
    h = (t + 1 + rng.integers(0, N - 1, size=extra)) % N
    tail = np.concatenate([ring_t, t])
    head = np.concatenate([ring_h, h])
    perm = rng.permutation(E)
    tail, head = tail[perm], head[perm]
    c = rng.integers(1, 10001, size=E).astype(np.float64)
    u = rng.integers(1, 1001, size=E).astype(np.float64)
    k = n_supply or max(2, N // 1000)
    nodes = rng.choice(N, size=2 * k, replace=False)
    amt = rng.integers(1, 50, size=k).astype(np.float64)
    b = np.zeros(N)
    b[nodes[:k]] = amt
    b[nodes[k:]] = -amt
    return tail, head, b, c, u


def mcf_interior_flow(u, seed):
    """x = u * r^4, r ~ U[0,1): skewed to 0, ~6 % above u/2 (exercises arc reversal)."""
    rng = np.random.default_rng(seed + 31337)
    r = rng.random(u.size)
    r2 = r * r
    return u * (r2 * r2)
