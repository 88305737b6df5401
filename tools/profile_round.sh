#!/bin/bash
# Round profile: bench line, launch lists, ncu full captures of the pricing / selection / tree kernels.
# Run under gpurun from the repo root; everything lands in gpurun_out/.
set -x
O=gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-tree --no-cpu"
python bench.py --steps 30 --warmup 5 > $O/r01_bench_v5_n1.json 2> $O/r01_bench_v5_n1.err || exit 1
$B > $O/plain_a.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"price|topk|pass_begin|merge|exchange" -c 24 \
    --csv --log-file $O/r01_launches_v5.csv $B > $O/ncu_a.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:price_dense_tma -s 3 -c 1 --csv --log-file $O/r01_price_traffic_c5_v5.csv $B > $O/ncu_b.log 2>&1
B2="python bench.py --size 20000 --steps 2 --warmup 3 --no-tree --no-cpu"
$B2 > $O/plain_b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:price_dense_tma -s 3 -c 1 -f -o $O/r01_price_tma_v5 $B2 > $O/ncu_c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:topk_select -s 3 -c 1 -f -o $O/r01_topk_select_v5 $B2 > $O/ncu_d.log 2>&1
T="python bench.py --tree-only 20000"
$T > $O/tree_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"score|pf_|rs_|ko_|kruskal|kr_|tree_" -c 400 \
    --csv --log-file $O/r01_tree_launches_v5.csv $T > $O/ncu_e.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"kruskal_kernel|tree_jump_kernel|pf_hist_kernel" -c 4 -f -o $O/r01_tree_kernels_v5 $T > $O/ncu_f.log 2>&1
echo done
