"""Problem classes of the network-crossover path: `StandardLP`, `MinCostFlow`, `OptTransport`.

Same fields, defaults and validation as the reference (`formats.py:83-161`) so instances
pickled by the reference's converters (`scripts/min2mcf.py`, `scripts/mnist2ot.py`) and code
that constructs them positionally keep working.  `GeneralLP` (`formats.py:10-80`) belongs to
the perturbation-crossover path and is out of scope.

Difference that matters at scale: `OptTransport.to_MCF` assembles the node-arc incidence
matrix directly in sparse form instead of going through a dense S x S identity
(`formats.py:155-158`), and the managers never need it for the device path.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Union

import numpy as np
import scipy.sparse as sp


@dataclass
class StandardLP:
    """min c^T x  s.t.  A x = b,  l <= x <= u  with l in {0, -inf} (default 0)."""

    A: Union[sp.csr_matrix, np.ndarray]
    b: np.ndarray
    c: np.ndarray
    u: np.ndarray
    name: str = "lp_instance"
    l: Optional[np.ndarray] = None

    def __post_init__(self) -> None:
        if self.l is None:
            self.l = np.zeros_like(self.u)


@dataclass
class MinCostFlow(StandardLP):
    """Min-cost flow as an LP over the node-arc incidence matrix A (N x E, one +1 and one -1
    per column; `scripts/min2mcf.py:36-37` puts +1 at the tail).  Supplies must balance."""

    name: str = "mcf_instance"

    def __post_init__(self) -> None:
        super().__post_init__()
        self.A = self.A.tocsr()
        if not np.isclose(np.sum(self.b), 0, atol=1e-8):
            raise ValueError("The sum of the b array must be equal to 0.")


def ot_incidence(S: int, D: int) -> sp.csr_matrix:
    """Incidence matrix of the complete bipartite graph K_{S,D} with arc k = i * D + j:
    -1 at source row i, +1 at sink row S + j (reference `formats.py:155-158`)."""
    n = S * D
    k = np.arange(n, dtype=np.int64)
    rows = np.concatenate([k // D, S + k % D])
    cols = np.concatenate([k, k])
    vals = np.concatenate([-np.ones(n), np.ones(n)])
    return sp.csr_matrix((vals, (rows, cols)), shape=(S + D, n))


@dataclass
class OptTransport:
    """Optimal transport between marginals s (S,) and d (D,) with cost matrix M (S x D)."""

    s: np.ndarray
    d: np.ndarray
    M: Union[sp.csr_matrix, np.ndarray]
    name: str = "ot_instance"

    def __post_init__(self) -> None:
        if not np.isclose(np.sum(self.s), np.sum(self.d), atol=1e-8):
            raise ValueError("The sum of the s and d arrays must be the same.")

    def to_MCF(self) -> MinCostFlow:
        """The equivalent min-cost flow: b = [-s, d], c = M row-major, no capacities."""
        S, D = self.s.size, self.d.size
        return MinCostFlow(A=ot_incidence(S, D), b=np.hstack([-self.s, self.d]),
                           c=np.asarray(self.M).flatten(), u=np.full(S * D, np.inf))
