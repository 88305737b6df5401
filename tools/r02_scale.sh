#!/bin/bash
# Round 2 scaling session on one 8-GPU box: real-rank parity tests, then bench lines at N = 8, 4, 2, 1.
O=gpurun_out
TAG=${1:-v17}
timeout 900 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_pricer.py -m gpu -x -q > $O/r02_multirank8_$TAG.log 2>&1; echo "multirank+pricer rc=$?"; tail -3 $O/r02_multirank8_$TAG.log
for n in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 100 --warmup 5 --no-cpu > $O/r02_bench_${TAG}_n$n.json 2> $O/r02_bench_${TAG}_n$n.err; echo "bench n$n rc=$?"; tail -2 $O/r02_bench_${TAG}_n$n.err | cut -c1-300
done
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu --no-tree > $O/r02_bench_${TAG}_n1.json 2> $O/r02_bench_${TAG}_n1.err; echo "bench n1 rc=$?"
