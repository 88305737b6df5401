#!/bin/bash
# bench lines at N = 8, 4, 2, 1 on one box (clean timed loop, in-kernel merge), no tests
O=gpurun_out
TAG=${1:-v21}
for n in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 200 --warmup 10 --no-cpu > $O/r02_bench_${TAG}_n$n.json 2> $O/r02_bench_${TAG}_n$n.err; echo "bench n$n rc=$?"
done
timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu --no-tree > $O/r02_bench_${TAG}_n1.json 2> $O/r02_bench_${TAG}_n1.err; echo "bench n1 rc=$?"
