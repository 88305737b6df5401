"""Radix argsort probe: n fp64 keys, one line per tuning code (sx_sort_set_tuning: 256 / 384 / 512 / 1024 threads per
downsweep tile)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "smart-crossover_b200"))
import torch  # noqa: E402
from smart_crossover import device as dev  # noqa: E402
from smart_crossover._native import lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 27
g = torch.Generator(device="cuda").manual_seed(1)
uni = torch.rand(n, generator=g, device="cuda", dtype=torch.float64)
key = (uni * uni * uni * uni) / (1.0 + 40.0 * torch.rand(n, generator=g, device="cuda", dtype=torch.float64)) ** 4
ref = None
for codes in [a.split(",") for a in sys.argv[2:]] or [["512"]]:
    for c in codes:
        assert lib.sx_sort_set_tuning(int(c)) == 0
    best = 1e9
    for rep in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        order, skey = dev.argsort_f64(key)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    if ref is None:
        ref = order.clone()
    same = bool((order == ref).all())
    print(f"n={n} tuning={'+'.join(codes)}: {best:.3f} ms ({n / best / 1e6:.2f} Gkeys/s) same_order_as_first={same}", flush=True)
lib.sx_sort_set_tuning(2)
