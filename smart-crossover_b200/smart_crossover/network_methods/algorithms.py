"""Network crossover drivers: `network_crossover` (TNET / CNET_OT / CNET_MCF) and `column_generation`.

Call surface, control flow, printed progress lines and returned `Output` follow the reference
(`network_methods/algorithms.py:14-144`); the managers run the scoring/sort, tree build and pricing
on the GPU.  Solver time is reported by the solver and added to the algorithm's own wall time, as
in the reference (:119-123, :76).
"""
from typing import Optional

import numpy as np

from smart_crossover.formats import MinCostFlow, OptTransport
from smart_crossover.network_methods.net_manager import MCFManagerStd, NetworkManager, OTManager
from smart_crossover.network_methods.tree_BI import tree_basis_identify
from smart_crossover.output import Output
from smart_crossover.parameters import COLUMN_GENERATION_RATIO
from smart_crossover.solver_caller.caller import SolverSettings
from smart_crossover.solver_caller.solving import generate_solver_caller
from smart_crossover.timer import Timer

_METHODS = ("tnet", "cnet_ot", "cnet_mcf")


def network_crossover(x: np.ndarray,
                      ot: Optional[OptTransport] = None,
                      mcf: Optional[MinCostFlow] = None,
                      method: str = "tnet",
                      solver: str = "GRB",
                      solver_settings: SolverSettings = SolverSettings(log_console=0)) -> Output:
    """Turn an interior / inexact flow `x` into an optimal basic solution.

    method 'tnet'     (OT)  : spanning-tree basis from the flow, then column generation;
           'cnet_ot'  (OT)  : big-M start, column generation over the flow-sorted arcs;
           'cnet_mcf' (MCF) : costs rescaled (this rebinds `mcf.c`, as in the reference), every arc
                              fixed at the bound nearer to x, big-M start, column generation.
    """
    print(f"*** Running {method} algorithm. ***")
    if method not in _METHODS:
        raise ValueError("Invalid method specified. Choose from 'tnet', 'cnet_ot', or 'cnet_mcf'.")
    # fail before any GPU work: an unavailable solver backend (the default "GRB" is kept from the reference;
    # only "HGS" ships here) or a barrier run that returned no interior point
    generate_solver_caller(solver, solver_settings)
    if x is None:
        raise ValueError("x is None: the interior-point run returned no solution (check its status)")
    timer = Timer()
    timer.start_timer()
    push_iter = 0
    manager = MCFManagerStd(mcf) if method == "cnet_mcf" else OTManager(ot)

    queue, flow_indicators = manager.get_sorted_flows(x)

    if method == "tnet":
        tree_basis, push_iter = tree_basis_identify(manager, flow_indicators)
        manager.set_basis(tree_basis)
        manager.add_free_variables(tree_basis.vbasis == 0)
    else:
        if method == "cnet_ot":
            manager.extend_by_bigM(manager.m * np.max(ot.M))
        else:
            manager.rescale_cost(np.max(np.abs(mcf.c)))
            manager.fix_variables(ind_fix_to_up=np.where(x >= mcf.u / 2)[0],
                                  ind_fix_to_low=np.where(x < mcf.u / 2)[0])
            manager.extend_by_bigM(manager.m * np.max(mcf.c))
        manager.update_subproblem()
        manager.set_initial_basis()

    timer.end_timer()
    cg = column_generation(manager, queue, solver, solver_settings)
    total = timer.total_duration + cg.runtime
    print(f"*** Optimal solution found with {cg.iter_count + push_iter} simplex iterations in {total} seconds. ***")
    return Output(x=cg.x, obj_val=cg.obj_val, runtime=total, iter_count=cg.iter_count + push_iter, basis=cg.basis)


def column_generation(net_manager: NetworkManager, queue: np.ndarray, solver: str,
                      solver_settings: SolverSettings) -> Output:
    """Grow the restricted master along `queue` (chunk doubling) until the full problem prices out.

    Chunk schedule of the reference (:102, :135-136): the first restricted master has 10 m columns
    when n / m > 1000, else int(1.2 m); the target then doubles every round.
    """
    timer = Timer()
    timer.start_timer()
    left = 0
    target = int(10 * net_manager.m) if net_manager.n / net_manager.m > 1000 else int(1.2 * net_manager.m)
    x, obj_val, iter_count, rounds = None, None, 0, 1
    while True:
        if left >= len(queue):
            print(' ##### Column generation fails! #####')
            break
        right = min(target, len(queue))
        net_manager.add_free_variables(queue[left:right])
        net_manager.update_subproblem()

        timer.end_timer()                                   # the solver reports its own time
        sub = net_manager.solve_subproblem(solver, solver_settings)
        obj_val = net_manager.recover_obj_val(sub.obj_val)
        timer.accumulate_time(sub.runtime)
        timer.start_timer()

        net_manager.set_basis(net_manager.recover_basis_from_sub_basis(sub.basis))
        x = net_manager.recover_x_from_sub_x(sub.x)
        optimal = net_manager.check_optimality_condition(x, sub.y)

        target = int(COLUMN_GENERATION_RATIO * target)
        left = right
        iter_count += sub.iter_count
        print(f"***  CG iteration {rounds} completed. ***")
        rounds += 1
        if optimal:
            break
    timer.end_timer()
    return Output(x=x, obj_val=obj_val, runtime=timer.total_duration, iter_count=iter_count, basis=net_manager.basis)
