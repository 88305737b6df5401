#!/bin/bash
# Validate the split / score / upsweep changes and time the score-kernel variants and the sort.
O=gpurun_out
python -m pytest tests -m gpu -x -q -k "prefix or score or sort or tree or golden or kruskal" > $O/pytest_v8b.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_v8b.log
for c in 3 4; do python bench.py --tree-only 20000 --score-ctas $c 2>&1 | cut -c1-700; done
python bench.py --tree-only 784 2>&1 | cut -c1-500
python tools/sort_probe.py 134217728 512 2>&1 | tail -4
