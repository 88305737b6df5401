"""HiGHS backend for the restricted-master LP re-solve (host side, off the timed path).

The reference re-solves the restricted master through its `SolverCaller`
interface (`solver_caller/caller.py:44-235`) with Gurobi / CPLEX / MOSEK
adapters (`solver_caller/gurobi.py`, `cplex.py`, `mosek.py`).  None of those
closed-source solvers is installed offline, so this module adds the one solver
that is: HiGHS, bundled with SciPy (`scipy.optimize._highspy._core`).  It
implements the subset of the interface the network path exercises
(SURVEY.md appendix A): `read_stdlp` / `read_mcf`, `add_warm_start_basis`,
`run_default`, `run_network_simplex`, `run_barrier_no_crossover`, `return_*`.

Basis codes follow the reference (Gurobi convention, `output.py:8-17`):
0 basic, -1 at lower, -2 at upper, -3 superbasic.  Duals have the sign of
Gurobi's `Pi`, i.e. reduced cost = c - A^T y (`net_manager.py:302,483`).
"""
import datetime
import time

import numpy as np
import scipy.sparse as sp
import scipy.optimize._highspy._core as _hc

from smart_crossover.output import Basis, Output
from smart_crossover.solver_caller.caller import SolverCaller

_TO_HIGHS = {0: _hc.HighsBasisStatus.kBasic, -1: _hc.HighsBasisStatus.kLower,
             -2: _hc.HighsBasisStatus.kUpper, -3: _hc.HighsBasisStatus.kZero}
_FROM_HIGHS = {_hc.HighsBasisStatus.kBasic: 0, _hc.HighsBasisStatus.kLower: -1,
               _hc.HighsBasisStatus.kUpper: -2, _hc.HighsBasisStatus.kZero: -3,
               _hc.HighsBasisStatus.kNonbasic: -1}


class HgsCaller(SolverCaller):
    """HiGHS adapter with the method names of the reference's `SolverCaller`."""

    solver_name = "HGS"

    def __init__(self, solver_settings=None):
        self.settings = solver_settings
        self.model = _hc._Highs()
        self.model.setOptionValue("output_flag", False)
        self._runtime = 0.0
        self._method = "default"
        self._n = 0
        self._m = 0

    # -- model input ------------------------------------------------------
    def read_stdlp(self, stdlp) -> None:
        """min c^T x, A x = b, l <= x <= u  (formats.py:83-101)."""
        A = sp.csc_matrix(stdlp.A)
        m, n = A.shape
        lp = _hc.HighsLp()
        lp.num_col_, lp.num_row_ = n, m
        lp.col_cost_ = np.asarray(stdlp.c, dtype=np.float64)
        lower = np.zeros(n) if stdlp.l is None else np.asarray(stdlp.l, dtype=np.float64)
        upper = np.asarray(stdlp.u, dtype=np.float64)
        lp.col_lower_ = np.where(np.isneginf(lower), -_hc.kHighsInf, lower)
        lp.col_upper_ = np.where(np.isposinf(upper), _hc.kHighsInf, upper)
        b = np.asarray(stdlp.b, dtype=np.float64)
        lp.row_lower_ = b
        lp.row_upper_ = b
        lp.a_matrix_.format_ = _hc.MatrixFormat.kColwise
        lp.a_matrix_.start_ = A.indptr.astype(np.int32)
        lp.a_matrix_.index_ = A.indices.astype(np.int32)
        lp.a_matrix_.value_ = A.data.astype(np.float64)
        self.model.passModel(lp)
        self._n, self._m = n, m

    def read_mcf(self, mcf) -> None:
        self.read_stdlp(mcf)

    def read_ot(self, ot) -> None:
        self.read_mcf(ot.to_MCF())

    def add_warm_start_basis(self, basis) -> None:
        hb = _hc.HighsBasis()
        hb.col_status = [_TO_HIGHS.get(int(v), _hc.HighsBasisStatus.kLower) for v in basis.vbasis]
        hb.row_status = [_hc.HighsBasisStatus.kBasic if int(v) == 0 else _hc.HighsBasisStatus.kLower
                         for v in basis.cbasis]
        hb.valid = True
        hb.alien = True
        self.model.setOptionValue("presolve", "off")
        self.model.setBasis(hb)

    # -- runs ---------------------------------------------------------------
    def _apply_settings(self) -> None:
        s = self.settings
        if s is None:
            return
        self.model.setOptionValue("time_limit", float(s.timeLimit))
        self.model.setOptionValue("dual_feasibility_tolerance", float(s.optimalityTol))
        self.model.setOptionValue("ipm_optimality_tolerance", float(s.barrierTol))
        if getattr(s, "log_console", 0) and getattr(s, "log_file", ""):
            self.model.setOptionValue("log_file", s.log_file)

    def _run(self) -> None:
        self._apply_settings()
        t0 = time.perf_counter()
        self.model.run()
        self._runtime = time.perf_counter() - t0

    def run_default(self) -> None:
        self._method = "default"
        self.model.setOptionValue("solver", "simplex")
        self._run()

    run_simplex = run_default
    run_network_simplex = run_default
    run_dual_simplex = run_default

    def run_primal_simplex(self) -> None:
        self._method = "default"
        self.model.setOptionValue("solver", "simplex")
        self.model.setOptionValue("simplex_strategy", 4)
        self._run()

    def run_barrier(self) -> None:
        self._method = "barrier"
        self.model.setOptionValue("solver", "ipm")
        self.model.setOptionValue("run_crossover", "on")
        self._run()

    def run_barrier_no_crossover(self) -> None:
        self._method = "barrier_nc"
        self.model.setOptionValue("solver", "ipm")
        self.model.setOptionValue("run_crossover", "off")
        self._run()

    # -- results ------------------------------------------------------------
    def return_status(self) -> str:
        st = self.model.getModelStatus()
        if st == _hc.HighsModelStatus.kOptimal:
            return "OPTIMAL"
        if st == _hc.HighsModelStatus.kInfeasible:
            return "INFEASIBLE"
        if st == _hc.HighsModelStatus.kUnbounded:
            return "UNBOUNDED"
        return "UNKNOWN"

    def return_x(self) -> np.ndarray:
        return np.array(self.model.getSolution().col_value, dtype=np.float64)

    def return_y(self) -> np.ndarray:
        return np.array(self.model.getSolution().row_dual, dtype=np.float64)

    def return_barx(self):
        return self.return_x() if self._method.startswith("barrier") else None

    def return_reduced_cost(self) -> np.ndarray:
        return np.array(self.model.getSolution().col_dual, dtype=np.float64)

    def return_obj_val(self) -> float:
        return float(self.model.getObjectiveValue())

    def return_runtime(self) -> datetime.timedelta:
        return datetime.timedelta(seconds=self._runtime)

    def return_iter_count(self) -> int:
        return int(self.model.getInfo().simplex_iteration_count)

    def return_bar_iter_count(self) -> int:
        return int(self.model.getInfo().ipm_iteration_count)

    def return_basis(self):
        if self._method == "barrier_nc":
            return None
        hb = self.model.getBasis()
        vbasis = np.array([_FROM_HIGHS[s] for s in hb.col_status])
        cbasis = np.array([0 if s == _hc.HighsBasisStatus.kBasic else -1 for s in hb.row_status])
        return Basis(vbasis, cbasis)

    def return_output(self):
        """Same shape as `SolverCaller.return_output` (caller.py:164-179)."""
        status = self.return_status()
        if status != "OPTIMAL":
            return Output(runtime=self.return_runtime(), status=status)
        return Output(x=self.return_x(), y=self.return_y(), x_bar=self.return_barx(),
                      obj_val=self.return_obj_val(), runtime=self.return_runtime(),
                      iter_count=self.return_iter_count(),
                      bar_iter_count=self.return_bar_iter_count(),
                      basis=self.return_basis(), status=status)
