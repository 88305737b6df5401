"""DIMACS min-cost-flow files (`.min`, e.g. the GOTO / NETGEN generators) -> pickled `MinCostFlow`.

Same conversion as the reference's `scripts/min2mcf.py:12-41`: `p min N E` gives the sizes, `n id b`
the non-zero supplies (1-based ids), `a tail head low cap cost` the arcs in file order; the incidence
matrix has +1 at the tail and -1 at the head of each arc (`:35-36`), the lower bound is ignored
(`:34`, bounds are [0, cap]).  The matrix is assembled directly in sparse form from the parsed arc
columns instead of element by element.

    python min2mcf.py INPUT_DIR [OUTPUT_DIR]        # every *.min -> *.mcf (pickle)
"""
import glob
import os
import pickle
import sys

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from smart_crossover.formats import MinCostFlow  # noqa: E402


def parse_min_file(file_path: str, name: str) -> MinCostFlow:
    num_nodes = num_arcs = None
    node_ids, node_b = [], []
    arcs = []
    with open(file_path, "r") as f:
        for line in f:
            if not line:
                continue
            kind = line[0]
            if kind == "a":
                arcs.append(line.split()[1:6])
            elif kind == "n":
                tok = line.split()
                node_ids.append(int(tok[1]))
                node_b.append(int(tok[2]))
            elif kind == "p" and num_nodes is None:
                tok = line.split()
                num_nodes, num_arcs = int(tok[2]), int(tok[3])
    if num_nodes is None:
        raise ValueError(f"{file_path}: no problem line ('p min <nodes> <arcs>')")
    b = np.zeros(num_nodes)
    if node_ids:
        b[np.asarray(node_ids) - 1] = node_b                       # later lines win, like the reference's loop
    data = np.asarray(arcs, dtype=np.int64).reshape(-1, 5)
    n_read = data.shape[0]
    tail, head, upper, cost = data[:, 0] - 1, data[:, 1] - 1, data[:, 3], data[:, 4]
    cols = np.arange(n_read, dtype=np.int64)
    # +1 at the tail, -1 at the head; a self-loop ends up as +1 (the reference writes -1 first, then +1)
    loop = tail == head
    rows = np.concatenate([tail, head[~loop]])
    vals = np.concatenate([np.ones(n_read, dtype=np.int64), -np.ones(int((~loop).sum()), dtype=np.int64)])
    A = sp.csr_matrix((vals, (rows, np.concatenate([cols, cols[~loop]]))), shape=(num_nodes, num_arcs), dtype=np.int64)
    c = np.zeros(num_arcs)
    u = np.zeros(num_arcs)
    c[:n_read] = cost
    u[:n_read] = upper
    return MinCostFlow(A=A, b=b, c=c, u=u, name=name)


def main(input_folder: str, output_folder: str = None) -> None:
    output_folder = output_folder or input_folder
    os.makedirs(output_folder, exist_ok=True)
    for min_file in sorted(glob.glob(os.path.join(input_folder, "*.min"))):
        base = os.path.basename(min_file)
        mcf = parse_min_file(min_file, base[:-4])
        with open(os.path.join(output_folder, os.path.splitext(base)[0] + ".mcf"), "wb") as f:
            pickle.dump(mcf, f)
        print(f"{base}: {mcf.b.size} nodes, {mcf.c.size} arcs")


if __name__ == "__main__":
    if len(sys.argv) < 2:
        sys.exit(__doc__)
    main(*sys.argv[1:3])
