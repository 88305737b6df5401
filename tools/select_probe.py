"""Probe of the top-k selection after a pricing pass: candidate counts, refinement levels, kernel time."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "smart-crossover_b200")]
import torch  # noqa: E402
import bench  # noqa: E402
from smart_crossover import device as dev  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 7500
D = int(sys.argv[2]) if len(sys.argv) > 2 else 60000
K = 1024
device = torch.device("cuda")
P, Q, a = bench.make_points(S, D, device)
M = bench.make_slab(P, Q, 0, S)


def state(pr):
    w = pr.sel[:128].view(torch.int32).cpu().numpy()
    n_cand = int(pr.sel[:8].view(torch.int64).item())
    return {"n_cand": n_cand, "bstar": int(w[2]), "n_sure": int(w[4]), "n_list": [int(x) for x in w[8:16]]}


def run(name, y):
    pr = dev.Pricer(device, K)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    tp = ts = 0.0
    for rep in range(6):
        pr.reset()
        ev[0].record()
        pr.price_dense(M, D, 0, S, D, y[:S], y[S:])
        ev[1].record()
        pr.select()
        ev[2].record()
        torch.cuda.synchronize()
        if rep >= 2:
            tp += ev[0].elapsed_time(ev[1]) / 4
            ts += ev[1].elapsed_time(ev[2]) / 4
    res = pr.fetch()
    print(f"{name}: price {tp * 1e3:.1f} us, select {ts * 1e3:.1f} us, violators {res.n_violating}, status {pr.status}, "
          f"{state(pr)}", flush=True)


y_planted = bench.planted_duals(P, Q, a, M, 0, 1e-4, 1)
run("planted + noise, 1e-4 violators (bench workload)", y_planted)
y_opt = bench.planted_duals(P, Q, a, M, 0, 0.0, 1)
run("optimal (no violators)", y_opt)
y_ties = y_opt.clone()
y_ties[S:] += 2e-3
run("every column's tight arc at the same rc (D near-ties on top)", y_ties)
y_far = y_planted.clone()
y_far[S:] += 0.05
run("far from optimal", y_far)
