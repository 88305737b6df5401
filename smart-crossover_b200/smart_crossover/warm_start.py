"""First-order warm start for the OT crossover: an entropic Sinkhorn point computed on the GPU.

The reference's experiment driver calls POT for it, `ot.sinkhorn(ot.s, ot.d, ot.M, reg=10,
numItermax=1000)` (`scripts/run_network_crossover.py:96`), and hands the flattened plan to
`network_crossover`.  POT is not part of the reference tree; `sinkhorn` below keeps its call shape
(marginals, cost matrix, `reg`, `numItermax`, `stopThr`; returns the S x D plan) and runs the same
Sinkhorn-Knopp iteration in the log domain on the device (`sx_sinkhorn_ot`, csrc/sx_sinkhorn.cu).
"""
from __future__ import annotations

import ctypes

import numpy as np


def sinkhorn(a, b, M, reg: float, numItermax: int = 1000, stopThr: float = 1e-9, log: bool = False):
    """Entropic-regularised OT plan gamma = diag(u) exp(-M / reg) diag(v) between marginals a (S,)
    and b (D,).  With log=True also returns {'niter', 'err', 'f', 'g'} (potentials reg*log u, reg*log v)."""
    import torch
    from smart_crossover import device as dev
    from smart_crossover._native import check, lib

    dev._require_cuda()
    a_t, b_t = dev._f64(np.asarray(a, dtype=np.float64)), dev._f64(np.asarray(b, dtype=np.float64))
    M_t = dev._f64(M) if not (isinstance(M, torch.Tensor) and M.is_cuda) else M.contiguous()
    S, D = M_t.shape
    if a_t.numel() != S or b_t.numel() != D:
        raise ValueError("marginals do not match the cost matrix")
    f = torch.empty(S, dtype=torch.float64, device=M_t.device)
    g = torch.empty(D, dtype=torch.float64, device=M_t.device)
    X = torch.empty(S, D, dtype=torch.float64, device=M_t.device)
    ws = dev._ws(lib.sx_sinkhorn_workspace_bytes(S, D), M_t.device)
    iters, err = ctypes.c_int64(0), ctypes.c_double(0.0)
    check(lib.sx_sinkhorn_ot(dev._ptr(M_t), M_t.stride(0), S, D, dev._ptr(a_t), dev._ptr(b_t), float(reg),
                             int(numItermax), float(stopThr), 10, dev._ptr(f), dev._ptr(g), dev._ptr(X),
                             ctypes.byref(iters), ctypes.byref(err), dev._ptr(ws), ws.numel(), dev._stream()),
          "sx_sinkhorn_ot")
    plan = X.cpu().numpy()
    if log:
        return plan, {"niter": iters.value, "err": err.value, "f": f.cpu().numpy(), "g": g.cpu().numpy()}
    return plan
