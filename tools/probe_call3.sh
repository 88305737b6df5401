#!/bin/bash
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/pytest_v8c.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_v8c.log
python bench.py --tree-only 20000 2>&1 | cut -c1-700
python bench.py --tree-only -1 2>&1 | cut -c1-700
python tools/sort_probe.py 134217728 512 2>&1 | tail -4
