"""Result containers shared by the solver interface and the crossover drivers.

Field-for-field compatible with the reference's `output.py:8-53`; basis codes follow
the Gurobi convention used throughout the reference (0 basic, -1 at lower bound,
-2 at upper bound, -3 superbasic).
"""
import datetime
from dataclasses import dataclass
from typing import Optional

import numpy as np

BASIC, AT_LOWER, AT_UPPER, SUPERBASIC = 0, -1, -2, -3


@dataclass
class Basis:
    """Variable / constraint basis status arrays (cast to int on construction,
    reference output.py:15-17)."""

    vbasis: np.ndarray
    cbasis: np.ndarray

    def __post_init__(self):
        self.vbasis = np.asarray(self.vbasis).astype(int)
        self.cbasis = np.asarray(self.cbasis).astype(int)

    def basic_columns(self) -> np.ndarray:
        return np.flatnonzero(self.vbasis == BASIC)


@dataclass(frozen=True)
class Output:
    """What a solve returns: primal/dual vertex, optional barrier point, objective, run time
    (a `datetime.timedelta`), iteration counts, reduced costs, basis, status string."""

    x: Optional[np.ndarray] = None
    y: Optional[np.ndarray] = None
    x_bar: Optional[np.ndarray] = None
    obj_val: Optional[float] = None
    runtime: Optional[datetime.timedelta] = None
    iter_count: Optional[float] = None
    bar_iter_count: Optional[int] = None
    rcost: Optional[np.ndarray] = None
    basis: Optional[Basis] = None
    status: Optional[str] = None

    def __str__(self) -> str:
        return (f"Output(obj_val={self.obj_val}, runtime={self.runtime}, "
                f"iter_count={self.iter_count}, bar_iter_count={self.bar_iter_count})")
