#!/bin/bash
O=gpurun_out
python -m pytest tests -m gpu -x -q -k "prefix or tree or kruskal or golden" > $O/pytest_v8d.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_v8d.log
T="python bench.py --tree-only 20000"
$T > $O/tree_plain.log 2>&1 || { tail -5 $O/tree_plain.log; exit 1; }
cut -c1-400 $O/tree_plain.log
python bench.py --tree-only 784 2>&1 | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"score|pf_|rs_|ko_|kruskal|kr_|tree_" -c 700 \
    --csv --log-file $O/r01_tree_launches_v8.csv $T > $O/ncu_e.log 2>&1
timeout 240 ncu --set full --clock-control none --import-source on -k regex:"score_ot_vec_kernel|pf_split_kernel|pf_filter_kernel|kruskal_kernel|tree_jump_kernel" -c 6 -f -o $O/r01_tree_kernels_v8 $T > $O/ncu_f.log 2>&1
tail -1 $O/ncu_f.log
