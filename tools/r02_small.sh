#!/bin/bash
# Round 2: small-instance paths (single-launch sorts, single-CTA Kruskal, cooperative Euler tour): full GPU
# suite + tree-build timings at 784^2, 20000^2 and the 1M/10M MCF.
O=gpurun_out
TAG=${1:-v8}
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02_pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r02_pytest_$TAG.log
python bench.py --tree-only 784 > $O/r02_tree_784_$TAG.json 2>&1; cut -c1-600 $O/r02_tree_784_$TAG.json
python bench.py --tree-only 20000 > $O/r02_tree_20000_$TAG.json 2>&1; cut -c1-600 $O/r02_tree_20000_$TAG.json
python bench.py --tree-only -1 > $O/r02_tree_mcf_$TAG.json 2>&1; cut -c1-600 $O/r02_tree_mcf_$TAG.json
