#!/bin/bash
# Final 8-GPU session of round 2: real-rank parity tests, bench at N = 8 and N = 1 (rate-balanced shards, one-process
# OTManager leg), then N = 4 and N = 2 if time allows.
O=gpurun_out
TAG=${1:-v31}
timeout 600 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_pricer.py -m gpu -x -q > $O/r02_multirank8_$TAG.log 2>&1; echo "multirank+pricer rc=$?"; tail -2 $O/r02_multirank8_$TAG.log
for n in 8 4 2; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 200 --warmup 10 --no-cpu > $O/r02_bench_${TAG}_n$n.json 2> $O/r02_bench_${TAG}_n$n.err; echo "bench n$n rc=$?"
done
timeout 400 python bench.py --steps 200 --warmup 10 --no-cpu --no-tree > $O/r02_bench_${TAG}_n1.json 2> $O/r02_bench_${TAG}_n1.err; echo "bench n1 rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 200 --warmup 10 --no-cpu --no-balance --no-manager > $O/r02_bench_${TAG}_n8_equal.json 2> $O/r02_bench_${TAG}_n8_equal.err; echo "bench n8 equal shards rc=$?"
