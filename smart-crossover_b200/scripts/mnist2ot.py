"""MNIST digit pairs -> pickled `OptTransport` instances (reference `scripts/mnist2ot.py:12-88`).

Each image, repeated k x k times per pixel and normalised to total mass 1, is a marginal; the cost
between pixels is their Manhattan distance on the (28 k) x (28 k) grid (integer valued, `:30-40`);
pixels with zero mass are dropped from both sides (`:47-55`).  The IDX file is decoded here directly
(the reference uses the `idx2numpy` package).

    python mnist2ot.py train-images-idx3-ubyte OUTPUT_DIR [num_images] [k ...]
"""
import os
import pickle
import struct
import sys
from typing import List

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from smart_crossover.formats import OptTransport  # noqa: E402

_IDX_DTYPES = {0x08: np.uint8, 0x09: np.int8, 0x0B: ">i2", 0x0C: ">i4", 0x0D: ">f4", 0x0E: ">f8"}


def read_idx(path: str) -> np.ndarray:
    """Decode an IDX file (magic: 0, 0, dtype code, ndim; then big-endian uint32 dims; then the data)."""
    with open(path, "rb") as f:
        raw = f.read()
    zero, code, ndim = struct.unpack(">HBB", raw[:4])
    if zero != 0 or code not in _IDX_DTYPES:
        raise ValueError(f"{path}: not an IDX file")
    dims = struct.unpack(">" + "I" * ndim, raw[4:4 + 4 * ndim])
    return np.frombuffer(raw, dtype=_IDX_DTYPES[code], offset=4 + 4 * ndim, count=int(np.prod(dims))).reshape(dims)


def load_mnist_data(data_folder: str) -> np.ndarray:
    return read_idx(os.path.join(data_folder, "train-images-idx3-ubyte"))


def select_random_images(x_train: np.ndarray, num_images: int = 20) -> np.ndarray:
    return x_train[np.random.choice(x_train.shape[0], num_images, replace=False)]


def normalize_and_amplify(image: np.ndarray, k: int) -> np.ndarray:
    big = np.kron(image.astype(np.float64), np.ones((k, k)))
    return big / np.sum(big)


def create_cost_matrix(k: int) -> np.ndarray:
    n = 28 * k
    yy, xx = np.divmod(np.arange(n * n), n)                        # row-major pixel -> (y, x)
    return np.abs(yy[:, None] - yy[None, :]) + np.abs(xx[:, None] - xx[None, :])


def make_opt_transport_instances(images: List[np.ndarray], cost_matrix: np.ndarray, k: int) -> List[OptTransport]:
    instances = []
    for i in range(0, len(images) - 1, 2):
        src, dst = images[i].ravel(), images[i + 1].ravel()
        rows, cols = np.flatnonzero(src), np.flatnonzero(dst)
        ot = OptTransport(s=src[rows], d=dst[cols], M=cost_matrix[np.ix_(rows, cols)], name=f"mnist_{k}_{i // 2 % 10}")
        print(ot.name)
        instances.append(ot)
    return instances


def save_opt_transport_instances(instances: List[OptTransport], output_folder: str) -> None:
    os.makedirs(output_folder, exist_ok=True)
    for i, inst in enumerate(instances):
        with open(os.path.join(output_folder, f"mnist_{i // 10 + 1}_{i % 10}.ot"), "wb") as f:
            pickle.dump(inst, f)


def main(idx_file: str, output_folder: str, num_images: int = 20, ks=(1, 2)) -> None:
    images = select_random_images(read_idx(idx_file), num_images)
    instances = []
    for k in ks:
        cost = create_cost_matrix(k)
        instances += make_opt_transport_instances([normalize_and_amplify(im, k) for im in images], cost, k)
    save_opt_transport_instances(instances, output_folder)


if __name__ == "__main__":
    if len(sys.argv) < 3:
        sys.exit(__doc__)
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    main(sys.argv[1], sys.argv[2], n, tuple(int(v) for v in sys.argv[4:]) or (1, 2))
