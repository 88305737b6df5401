#!/bin/bash
# Round profile (v8): full GPU test suite, bench line, launch lists, ncu full captures of the tree-build kernels.
# Run under gpurun from the repo root; everything lands in gpurun_out/.  The pricing / selection kernels are
# unchanged since the v5 captures (r01_price_tma_v5, r01_topk_select_v5); the sort / Kruskal-order captures
# of the current code are r01_sort_v8 / r01_ko_v8 (tools/profile_sort.sh).
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/pytest_v8_final.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_v8_final.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_v8.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke_v8.log
python bench.py > $O/r01_bench_v8_n1.json 2> $O/r01_bench_v8_n1.err || { tail -5 $O/r01_bench_v8_n1.err; exit 1; }
cut -c1-400 $O/r01_bench_v8_n1.json
B="python bench.py --steps 3 --warmup 3 --no-tree --no-cpu"
$B > $O/plain_a.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"price|topk|pass_begin|merge|exchange" -c 24 \
    --csv --log-file $O/r01_launches_v8.csv $B > $O/ncu_a.log 2>&1
T="python bench.py --tree-only 20000"
$T > $O/tree_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"score|pf_|rs_|ko_|kruskal|kr_|tree_" -c 700 \
    --csv --log-file $O/r01_tree_launches_v8.csv $T > $O/ncu_e.log 2>&1
timeout 240 ncu --set full --clock-control none --import-source on -k regex:"score_ot_vec_kernel|pf_split_kernel|pf_filter_kernel|kruskal_kernel|tree_jump_kernel" -c 6 -f -o $O/r01_tree_kernels_v8 $T > $O/ncu_f.log 2>&1
tail -1 $O/ncu_f.log
echo done
