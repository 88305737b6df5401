"""Probe for the radix argsort: stable argsort of n fp64 keys, uniform and score-like, per tuning."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "smart-crossover_b200"))
import torch  # noqa: E402
from smart_crossover import device as dev  # noqa: E402
from smart_crossover._native import lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 27
modes = [int(a) for a in sys.argv[2:]] or [512]
g = torch.Generator(device="cuda").manual_seed(1)
uni = torch.rand(n, generator=g, device="cuda", dtype=torch.float64)
score = (uni * uni * uni * uni) / (1.0 + 40.0 * torch.rand(n, generator=g, device="cuda", dtype=torch.float64)) ** 4
for name, key in (("uniform", uni), ("score-like", score)):
    for threads in modes:
        lib.sx_sort_set_tuning(threads)
        best = 1e9
        for rep in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            order, skey = dev.argsort_f64(key)
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        ok = bool((skey[1:] >= skey[:-1]).all())
        print(f"{name:10s} n={n} threads={threads}: {best * 1e3:.3f} ms ({n / best / 1e9:.2f} Gkeys/s) sorted={ok}", flush=True)
