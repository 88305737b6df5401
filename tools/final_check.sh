#!/bin/bash
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/pytest_v9_final.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_v9_final.log
python bench.py --tree-only -1 > $O/tree_mcf_v9.json 2>&1; cut -c1-420 $O/tree_mcf_v9.json
python bench.py --tree-only 20000 > $O/tree_20000_v9.json 2>&1; cut -c1-300 $O/tree_20000_v9.json
python bench.py --tree-only 784 > $O/tree_784_v9.json 2>&1; cut -c1-300 $O/tree_784_v9.json
