#!/bin/bash
# Round 2: persistent pricer behind the managers (sx_ot_pricer) -- parity on 1 and 2 GPUs in one process,
# end-to-end tests, bench lines with the OTManager leg.
O=gpurun_out
TAG=${1:-v6}
timeout 900 python -m pytest tests/test_gpu_pricer.py tests/test_gpu_e2e.py -m gpu -x -q > $O/r02_pricer_$TAG.log 2>&1; echo "pricer+e2e rc=$?"; tail -5 $O/r02_pricer_$TAG.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-tree --no-cpu > $O/r02_bench_${TAG}_n1.json 2> $O/r02_bench_${TAG}_n1.err; echo "bench n1 rc=$?"; cut -c1-200 $O/r02_bench_${TAG}_n1.json; tail -3 $O/r02_bench_${TAG}_n1.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu > $O/r02_bench_${TAG}_n2.json 2> $O/r02_bench_${TAG}_n2.err; echo "bench n2 rc=$?"; cut -c1-200 $O/r02_bench_${TAG}_n2.json; tail -3 $O/r02_bench_${TAG}_n2.err
