"""Probe for ncu: one full argsort + Kruskal order of n score-like fp64 keys (default 4e8)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "smart-crossover_b200"))
import torch  # noqa: E402
from smart_crossover import device as dev  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000_000
g = torch.Generator(device="cuda").manual_seed(1)
key = torch.rand(n, generator=g, device="cuda", dtype=torch.float64)
key = key * key * key * key
for rep in range(2):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    order, skey = dev.argsort_f64(key)
    ev[1].record()
    korder = dev.kruskal_order(skey, order)
    ev[2].record()
    torch.cuda.synchronize()
    print(f"n={n} argsort {ev[0].elapsed_time(ev[1]):.3f} ms, kruskal_order {ev[1].elapsed_time(ev[2]):.3f} ms", flush=True)
