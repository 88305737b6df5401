#!/bin/bash
# Full GPU test suite, default bench line, and ncu full captures of the score / sort kernels (20 000^2).
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/pytest_v8.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_v8.log
T="python bench.py --tree-only 20000"
$T > $O/tree_plain_v8.log 2>&1 || { tail -5 $O/tree_plain_v8.log; exit 1; }
cat $O/tree_plain_v8.log | cut -c1-600
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"score_ot_vec_kernel|rs_downsweep_kernel|rs_upsweep_kernel|ko_scatter|ko_block_heads|pf_split" -c 40 -f -o $O/r01_sort_score_v8 $T > $O/ncu_g.log 2>&1
tail -2 $O/ncu_g.log
