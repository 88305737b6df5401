#!/usr/bin/env python
"""End-to-end goldens at the larger sizes, from the UNMODIFIED reference (authoring container only):

    python tests/golden/make_golden_e2e.py

`network_crossover` (reference network_methods/algorithms.py:14-78) with the HiGHS caller patched in exactly as
in make_golden.py, on
  * the 784 x 784 instance of BASELINE.json configs[1] (inputs regenerated from the seed, digest checked
    against digests.json): methods tnet and cnet_ot -> e2e_c2_784.npz
  * the 20 000-node / 200 000-arc NETGEN-style MCF of digests.json: method cnet_mcf -> e2e_mcf_20k.npz
Stored: objective, simplex iteration count, the basic arcs of the final basis, wall time of the reference run.
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (installs the shim, patches the solver factory)
import cases  # noqa: E402
from make_golden import MinCostFlow, OptTransport, QUIET, build_mcf, ref_alg, ref_solving  # noqa: E402


def main():
    digests = json.load(open(os.path.join(HERE, "digests.json")))
    s, d, M = cases.ot_grid(28, 20260002)
    x = cases.interior_flow(s, d, M, 20260002, 260.0)
    assert cases.digest(s, d, M, x) == digests["ot_c2_784"]["inputs"]
    out = {}
    for method in ("tnet", "cnet_ot"):
        t0 = time.perf_counter()
        res = ref_alg.network_crossover(x=x.copy(), ot=OptTransport(s.copy(), d.copy(), M.copy()), method=method,
                                        solver="HGS", solver_settings=QUIET)
        out[f"{method}_seconds"] = np.float64(time.perf_counter() - t0)
        out[f"{method}_obj"] = np.float64(res.obj_val)
        out[f"{method}_iters"] = np.int64(res.iter_count)
        out[f"{method}_basic"] = np.flatnonzero(res.basis.vbasis == 0).astype(np.int64)
    direct = ref_solving.solve_mcf(OptTransport(s, d, M).to_MCF(), solver="HGS", settings=QUIET)
    out["direct_obj"] = np.float64(direct.obj_val)
    mg.save_full("e2e_c2_784", {}, out)

    tail, head, b, c, u = cases.netgen_like(20000, 200000, 23)
    x = cases.mcf_interior_flow(u, 23)
    assert cases.digest(tail, head, b, c, u, x) == digests["mcf_mid_20k"]["inputs"]
    t0 = time.perf_counter()
    res = ref_alg.network_crossover(x=x.copy(), mcf=build_mcf(tail, head, b, c, u), method="cnet_mcf",
                                    solver="HGS", solver_settings=QUIET)
    out = {"cnet_mcf_seconds": np.float64(time.perf_counter() - t0), "cnet_mcf_obj": np.float64(res.obj_val),
           "cnet_mcf_iters": np.int64(res.iter_count),
           "cnet_mcf_basic": np.flatnonzero(res.basis.vbasis == 0).astype(np.int64)}
    direct = ref_solving.solve_mcf(build_mcf(tail, head, b, c, u), solver="HGS", settings=QUIET)
    out["direct_obj"] = np.float64(direct.obj_val)
    mg.save_full("e2e_mcf_20k", {}, out)


if __name__ == "__main__":
    main()
