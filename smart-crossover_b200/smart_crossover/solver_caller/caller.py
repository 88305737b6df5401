"""Solver interface of the restricted-master re-solve (host side; not on the timed path).

`SolverSettings` has the reference's fields and defaults (`solver_caller/caller.py:17-41`).
`SolverCaller` lists the operations the crossover drivers rely on (`caller.py:44-235`); the only
backend shipped here is HiGHS (`highs.py`) -- the reference's Gurobi / CPLEX / MOSEK adapters wrap
closed-source solvers that are not installable offline and are out of scope.
"""
from dataclasses import dataclass


@dataclass
class SolverSettings:
    presolve: str = "on"
    crossover: str = "on"
    barrierTol: float = 1e-8
    optimalityTol: float = 1e-6
    timeLimit: int = 3600
    log_file: str = ""
    log_console: int = 1
    iterLimit: int = 1000
    simplexPricing: str = ""


class SolverCaller:
    """Protocol of a solver backend.  See `highs.HgsCaller` for the implementation."""

    solver_name: str = ""

    def read_stdlp(self, stdlp): raise NotImplementedError
    def read_mcf(self, mcf): raise NotImplementedError
    def read_ot(self, ot): self.read_mcf(ot.to_MCF())
    def add_warm_start_basis(self, basis): raise NotImplementedError
    def run_default(self): raise NotImplementedError
    def run_simplex(self): raise NotImplementedError
    def run_network_simplex(self): raise NotImplementedError
    def run_primal_simplex(self): raise NotImplementedError
    def run_dual_simplex(self): raise NotImplementedError
    def run_barrier(self): raise NotImplementedError
    def run_barrier_no_crossover(self): raise NotImplementedError
    def return_output(self): raise NotImplementedError
