#!/bin/bash
# Round 2: fused pricing pass -- parity (1 GPU), real-rank parity (2 GPUs), bench lines N=1 / N=2.
O=gpurun_out
TAG=${1:-v4}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "price or fused or emulated or slab or c4_full or c5_full" > $O/r02_parity_price_$TAG.log 2>&1; echo "parity-price rc=$?"; tail -5 $O/r02_parity_price_$TAG.log
timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q > $O/r02_multirank_$TAG.log 2>&1; echo "multirank rc=$?"; tail -5 $O/r02_multirank_$TAG.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-tree --no-cpu > $O/r02_bench_${TAG}_n1.json 2> $O/r02_bench_${TAG}_n1.err; echo "bench n1 rc=$?"; cut -c1-300 $O/r02_bench_${TAG}_n1.json; tail -3 $O/r02_bench_${TAG}_n1.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu > $O/r02_bench_${TAG}_n2.json 2> $O/r02_bench_${TAG}_n2.err; echo "bench n2 rc=$?"; cut -c1-300 $O/r02_bench_${TAG}_n2.json; tail -3 $O/r02_bench_${TAG}_n2.err
