"""`solve_mcf` / `solve_ot` / `solve_lp`: build a solver backend, load the problem, run a method.

Call surface of the reference's `solver_caller/solving.py:13-133`.  Solver keys: "HGS" (HiGHS,
bundled with SciPy) is available; "GRB", "CPL", "MSK" name the reference's vendor adapters, which
are out of scope here and raise with an explanation.
"""
from typing import Optional

from smart_crossover.formats import MinCostFlow, OptTransport, StandardLP
from smart_crossover.output import Basis, Output
from smart_crossover.solver_caller.caller import SolverSettings

_SIMPLEX_LIKE = ("default", "simplex", "network_simplex", "primal_simplex", "dual_simplex")


def generate_solver_caller(solver: str = "GRB", solver_settings: SolverSettings = SolverSettings()):
    if solver == "HGS":
        from smart_crossover.solver_caller.highs import HgsCaller
        return HgsCaller(solver_settings)
    if solver in ("GRB", "CPL", "MSK"):
        raise ImportError(f"solver '{solver}': the Gurobi/CPLEX/MOSEK adapters of the reference wrap closed-source "
                          "solvers and are not part of this build; pass solver='HGS' (HiGHS).")
    raise ValueError("Invalid solver specified. Choose from 'GRB', 'CPL', 'MSK' and 'HGS'.")


def solve_problem(solver_caller, method: str, settings: SolverSettings,
                  warm_start_basis: Optional[Basis] = None, warm_start_solution=None) -> Output:
    if method in _SIMPLEX_LIKE:
        if warm_start_solution is not None and hasattr(solver_caller, "add_warm_start_solution"):
            solver_caller.add_warm_start_solution(warm_start_solution)
        if warm_start_basis is not None:
            solver_caller.add_warm_start_basis(warm_start_basis)
        {"default": solver_caller.run_default, "simplex": solver_caller.run_simplex,
         "network_simplex": solver_caller.run_network_simplex,
         "primal_simplex": solver_caller.run_primal_simplex,
         "dual_simplex": solver_caller.run_dual_simplex}[method]()
    elif method == "barrier":
        if settings.crossover == "on":
            solver_caller.run_barrier()
        else:
            solver_caller.run_barrier_no_crossover()
    else:
        raise ValueError("Invalid method specified. Choose from 'default' or 'barrier'.")
    return solver_caller.return_output()


def solve_lp(lp: StandardLP, solver: str = "GRB", method: str = "default",
             settings: SolverSettings = SolverSettings(), warm_start_basis: Optional[Basis] = None,
             warm_start_solution=None) -> Output:
    if not isinstance(lp, StandardLP):
        raise ValueError("Invalid LP format.")
    caller = generate_solver_caller(solver, settings)
    caller.read_stdlp(lp)
    return solve_problem(caller, method, settings, warm_start_basis, warm_start_solution)


def solve_mcf(mcf: MinCostFlow, solver: str = "GRB", method: str = "default",
              settings: SolverSettings = SolverSettings(), warm_start_basis: Optional[Basis] = None) -> Output:
    caller = generate_solver_caller(solver, settings)
    caller.read_mcf(mcf)
    return solve_problem(caller, method, settings, warm_start_basis)


def solve_ot(ot: OptTransport, solver: str = "GRB", method: str = "default",
             settings: SolverSettings = SolverSettings(), warm_start_basis: Optional[Basis] = None) -> Output:
    caller = generate_solver_caller(solver, settings)
    caller.read_ot(ot)
    return solve_problem(caller, method, settings, warm_start_basis)
