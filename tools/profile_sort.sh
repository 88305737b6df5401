#!/bin/bash
O=gpurun_out
python tools/ko_probe.py > $O/ko_plain.log 2>&1 || { tail -5 $O/ko_plain.log; exit 1; }
cat $O/ko_plain.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"ko_" -c 3 -f -o $O/r01_ko_v8 python tools/ko_probe.py > $O/ncu_h.log 2>&1; tail -1 $O/ncu_h.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"rs_downsweep|rs_upsweep" -c 4 -f -o $O/r01_sort_v8 python tools/ko_probe.py > $O/ncu_i.log 2>&1; tail -1 $O/ncu_i.log
