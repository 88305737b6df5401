#!/bin/bash
# Round 2 profile (run under gpurun, 1 GPU): DRAM traffic of the fused pricing kernel at the slab shapes of
# N = 1 and N = 8 for both legs, launch list of a bench run, ncu --set full of the fused kernel at 20 000^2.
O=gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-tree --no-cpu --no-c4 --no-manager"
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"
for shape in "60000 60000" "60000 7500" "20000 20000" "20000 2500"; do
  set -- $shape
  $B --size $1 --rows $2 > $O/plain_$1_$2.log 2>&1 || { tail -3 $O/plain_$1_$2.log; continue; }
  ncu --metrics $M --clock-control none -k regex:price_fused --launch-skip 4 -c 2 --csv \
      --log-file $O/r02_price_traffic_$1_$2.csv $B --size $1 --rows $2 > $O/ncu_tr_$1_$2.log 2>&1
  echo "traffic $1 x $2 rc=$?"
done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"price|topk|pass_begin|merge|exchange|fused" -c 24 \
    --csv --log-file $O/r02_launches.csv $B > $O/ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:price_fused --launch-skip 4 -c 1 -f \
    -o $O/r02_price_fused_20000 $B --size 20000 > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -1 $O/ncu_full.log
