#!/bin/bash
# tail stealing on/off: parity tests, then bench at the N = 1 and N = 8 slab shapes of both legs
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pricer.py -m gpu -x -q -k "price or fused or pricer or c4_full or c5_full or emulated" > $O/r02_steal_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/r02_steal_tests.log
for st in 0 1; do
  for shape in "60000 60000" "60000 7500" "20000 20000" "20000 2500"; do
    set -- $shape
    SX_FUSED_STEAL=$st python bench.py --size $1 --rows $2 --no-c4 --steps 100 --warmup 10 --no-tree --no-cpu --no-manager > $O/r02_steal${st}_$1_$2.json 2>$O/r02_steal.err || tail -2 $O/r02_steal.err
  done
done
