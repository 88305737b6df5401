"""Data converters against fixtures produced by the reference's own
scripts (tests/golden/make_golden_scripts.py): DIMACS .min -> MinCostFlow, MNIST -> OptTransport,
pickle round trip through the loaders.  CPU only."""
import os
import pickle
import struct
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, os.path.join(ROOT, "smart-crossover_b200", "scripts"))

import min2mcf  # noqa: E402
import mnist2ot  # noqa: E402


def test_parse_min_file_matches_the_reference_converter():
    ref = np.load(os.path.join(GOLD, "min2mcf_tiny.npz"))
    mcf = min2mcf.parse_min_file(os.path.join(GOLD, "tiny_goto.min"), "tiny_goto")
    assert mcf.name == "tiny_goto" and str(mcf.A.dtype) == str(ref["A_dtype"])
    assert np.array_equal(mcf.A.toarray(), ref["A"])
    for k in ("b", "c", "u", "l"):
        assert np.array_equal(getattr(mcf, k), ref[k]) and getattr(mcf, k).dtype == ref[k].dtype
    # +1 at the tail, -1 at the head (scripts/min2mcf.py:35-36); lower bounds ignored
    assert mcf.A[0, 0] == 1 and mcf.A[1, 0] == -1 and mcf.u[9] == 2 and np.all(mcf.l == 0)


def test_min_file_folder_conversion_and_loader(tmp_path):
    import shutil
    import run_network_crossover as drv
    shutil.copy(os.path.join(GOLD, "tiny_goto.min"), tmp_path / "tiny_goto.min")
    min2mcf.main(str(tmp_path))
    (mcf,) = drv.load_min_cost_flow_instances(str(tmp_path))
    ref = np.load(os.path.join(GOLD, "min2mcf_tiny.npz"))
    assert mcf.name == "tiny_goto" and np.array_equal(mcf.A.toarray(), ref["A"]) and np.array_equal(mcf.b, ref["b"])


def _write_idx(path, images):
    with open(path, "wb") as f:
        f.write(struct.pack(">HBB", 0, 0x08, images.ndim))
        f.write(struct.pack(">" + "I" * images.ndim, *images.shape))
        f.write(images.astype(np.uint8).tobytes())


def test_mnist_converter_matches_the_reference(tmp_path):
    ref = np.load(os.path.join(GOLD, "mnist2ot_tiny.npz"))
    imgs = ref["images"]
    _write_idx(tmp_path / "train-images-idx3-ubyte", imgs)
    assert np.array_equal(mnist2ot.load_mnist_data(str(tmp_path)), imgs)
    for k in (1, 2):
        norm = [mnist2ot.normalize_and_amplify(im, k) for im in imgs]
        assert norm[0].tobytes() == ref[f"norm0_k{k}"].tobytes()
        cost = mnist2ot.create_cost_matrix(k)
        assert tuple(cost.shape) == tuple(ref[f"cost_k{k}_shape"])
        assert np.array_equal(cost[:40, -40:], ref[f"cost_k{k}_block"])
        assert cost.sum() == ref[f"cost_k{k}_sum"][0]
        assert (cost * np.arange(cost.shape[1])[None, :]).sum() == ref[f"cost_k{k}_sum"][1]
        inst = mnist2ot.make_opt_transport_instances(norm, cost, k)
        assert len(inst) == int(ref[f"n_inst_k{k}"])
        for q, ot in enumerate(inst):
            assert ot.name == str(ref[f"k{k}_i{q}_name"])
            assert ot.s.tobytes() == ref[f"k{k}_i{q}_s"].tobytes() and ot.d.tobytes() == ref[f"k{k}_i{q}_d"].tobytes()
            assert [ot.M.sum(), ot.M.shape[0], ot.M.shape[1]] == list(ref[f"k{k}_i{q}_Msum"])
            if k == 1:
                assert np.array_equal(ot.M, ref[f"k{k}_i{q}_M"])


def test_ot_pickles_round_trip_through_the_loader(tmp_path):
    import run_network_crossover as drv
    ref = np.load(os.path.join(GOLD, "mnist2ot_tiny.npz"))
    norm = [mnist2ot.normalize_and_amplify(im, 1) for im in ref["images"]]
    inst = mnist2ot.make_opt_transport_instances(norm, mnist2ot.create_cost_matrix(1), 1)
    mnist2ot.save_opt_transport_instances(inst, str(tmp_path))
    back = drv.load_opt_transport_instances(str(tmp_path))
    assert [o.name for o in back] == [o.name for o in inst]
    assert all(np.array_equal(a.M, b.M) and np.array_equal(a.s, b.s) for a, b in zip(inst, back))
    with open(tmp_path / "mnist_1_0.ot", "rb") as f:
        assert type(pickle.load(f)).__module__ == "smart_crossover.formats"     # the path reference pickles name
