#!/usr/bin/env python
"""Golden fixtures for the data converters, produced by the UNMODIFIED reference scripts
(`/root/reference/scripts/min2mcf.py`, `mnist2ot.py`).  Authoring container only.

    python tests/golden/make_golden_scripts.py
"""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_shim  # noqa: E402

_ref_shim.install()
sys.modules.setdefault("idx2numpy", types.ModuleType("idx2numpy"))      # imported at module top, unused here


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def synthetic_images(n=4, seed=5):
    rng = np.random.default_rng(seed)
    imgs = rng.integers(0, 256, size=(n, 28, 28)).astype(np.uint8)
    imgs[rng.random((n, 28, 28)) < 0.8] = 0                            # mostly background, like MNIST
    return imgs


def main():
    ref_min = load("/root/reference/scripts/min2mcf.py", "ref_min2mcf")
    mcf = ref_min.parse_min_file(os.path.join(HERE, "tiny_goto.min"), "tiny_goto")
    np.savez_compressed(os.path.join(HERE, "min2mcf_tiny.npz"), A=mcf.A.toarray(), b=mcf.b, c=mcf.c, u=mcf.u,
                        l=mcf.l, A_dtype=str(mcf.A.dtype))

    ref_mn = load("/root/reference/scripts/mnist2ot.py", "ref_mnist2ot")
    imgs = synthetic_images()
    out = {"images": imgs}
    for k in (1, 2):
        norm = [ref_mn.normalize_and_amplify(im, k) for im in imgs]
        out[f"norm0_k{k}"] = norm[0]
        cost = ref_mn.create_cost_matrix(k)
        out[f"cost_k{k}_shape"] = np.array(cost.shape)
        out[f"cost_k{k}_sum"] = np.array([cost.sum(), (cost * np.arange(cost.shape[1])[None, :]).sum()])
        out[f"cost_k{k}_block"] = cost[:40, -40:]
        inst = ref_mn.make_opt_transport_instances(norm, cost, k)
        out[f"n_inst_k{k}"] = np.array(len(inst))
        for q, ot in enumerate(inst):
            out[f"k{k}_i{q}_s"], out[f"k{k}_i{q}_d"] = ot.s, ot.d
            out[f"k{k}_i{q}_Msum"] = np.array([ot.M.sum(), ot.M.shape[0], ot.M.shape[1]])
            if k == 1:
                out[f"k{k}_i{q}_M"] = ot.M
            out[f"k{k}_i{q}_name"] = np.array(ot.name)
    np.savez_compressed(os.path.join(HERE, "mnist2ot_tiny.npz"), **out)
    print("wrote min2mcf_tiny.npz, mnist2ot_tiny.npz")


if __name__ == "__main__":
    main()
