#!/bin/bash
O=gpurun_out
python bench.py > $O/r01_bench_v8_n1.json 2> $O/r01_bench_v8_n1.err || { tail -5 $O/r01_bench_v8_n1.err; exit 1; }
cut -c1-300 $O/r01_bench_v8_n1.json
T="python bench.py --tree-only 20000"
$T > $O/tree_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"score|pf_|rs_|ko_|kruskal|kr_|tree_" -c 700 \
    --csv --log-file $O/r01_tree_launches_v8.csv $T > $O/ncu_e.log 2>&1
echo done
