"""Experiment driver of the network crossover (reference `scripts/run_network_crossover.py`).

Loads pickled instances (`.ot` from mnist2ot.py, `.mcf` from min2mcf.py), produces the interior /
first-order point the crossover starts from, and runs `network_crossover`:

  ot   'total'      Sinkhorn warm start on the GPU (`warm_start.sinkhorn`, reg = 10, 1000 iterations, as
                    the reference's POT call at :96), then TNET and CNET_OT
       'crossover'  barrier point without crossover from the LP solver, then TNET and CNET_OT
  mcf  'crossover'  barrier point without crossover, then CNET_MCF

The reference runs Gurobi / CPLEX; offline the only LP backend is HiGHS (`solver='HGS'`).

    python run_network_crossover.py ot|mcf FOLDER [total|crossover] [solver]
"""
import os
import pickle
import sys
from datetime import datetime
from typing import List

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from smart_crossover.formats import MinCostFlow, OptTransport  # noqa: E402
from smart_crossover.network_methods.algorithms import network_crossover  # noqa: E402
from smart_crossover.solver_caller.caller import SolverSettings  # noqa: E402
from smart_crossover.solver_caller.solving import solve_mcf, solve_ot  # noqa: E402
from smart_crossover.warm_start import sinkhorn  # noqa: E402


def _load(folder: str, suffix: str, prefix: str = "") -> list:
    out = []
    for name in sorted(os.listdir(folder)):
        if name.endswith(suffix) and name.startswith(prefix):
            with open(os.path.join(folder, name), "rb") as f:
                out.append(pickle.load(f))
    return out


def load_opt_transport_instances(folder: str) -> List[OptTransport]:
    return _load(folder, ".ot")


def load_min_cost_flow_instances(folder: str, prefix: str = "") -> List[MinCostFlow]:
    return _load(folder, ".mcf", prefix)


def main(problem: str, folder: str, test_object: str = "crossover", solver: str = "HGS") -> dict:
    results = {}
    if problem in ("mcf", "goto"):
        for mcf in load_min_cost_flow_instances(folder):
            bar = solve_mcf(mcf, method="barrier", solver=solver, settings=SolverSettings(crossover="off"))
            if bar.status != "OPTIMAL" or bar.x_bar is None:
                print(f"{mcf.name}: barrier run ended with status {bar.status}; skipped")
                continue
            out = network_crossover(x=bar.x_bar, mcf=mcf, method="cnet_mcf", solver=solver,
                                    solver_settings=SolverSettings(presolve="on"))
            results[mcf.name] = {"cnet": (out.runtime, out.iter_count), "obj": out.obj_val}
    elif problem == "ot":
        for ot in load_opt_transport_instances(folder):
            if test_object == "total":
                start = datetime.now()
                x = sinkhorn(ot.s, ot.d, ot.M, reg=10, numItermax=1000).flatten()
                warm = datetime.now() - start
            else:
                bar = solve_ot(ot, method="barrier", solver=solver, settings=SolverSettings(crossover="off"))
                if bar.status != "OPTIMAL" or bar.x_bar is None:
                    print(f"{ot.name}: barrier run ended with status {bar.status}; skipped")
                    continue
                x, warm = bar.x_bar, None
            tnet = network_crossover(x=x, ot=ot, method="tnet", solver=solver, solver_settings=SolverSettings(presolve="on"))
            cnet = network_crossover(x=x, ot=ot, method="cnet_ot", solver=solver, solver_settings=SolverSettings(presolve="on"))
            results[ot.name] = {"sinkhorn": warm, "tnet": (tnet.runtime, tnet.iter_count),
                                "cnet": (cnet.runtime, cnet.iter_count), "obj": tnet.obj_val}
    else:
        raise ValueError("problem must be 'ot', 'mcf' or 'goto'")
    for name, r in results.items():
        print(name, r)
    return results


if __name__ == "__main__":
    if len(sys.argv) < 3:
        sys.exit(__doc__)
    main(*sys.argv[1:5])
