#!/bin/bash
# Round 2: real-rank parity test + N=1 / N=2 bench lines (both legs: 60 000^2 and 20 000^2).
# Run under `gpurun --gpus 2`; everything lands in gpurun_out/.
O=gpurun_out
TAG=${1:-v1}
EXTRA=${2:-}
python -m pytest tests/test_gpu_multirank.py -m gpu -x -q > $O/r02_multirank_$TAG.log 2>&1; echo "multirank rc=$?"; tail -3 $O/r02_multirank_$TAG.log
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "nan or price_dense_random or row_slabs or emulated" > $O/r02_parity_quick_$TAG.log 2>&1; echo "parity-quick rc=$?"; tail -2 $O/r02_parity_quick_$TAG.log
python bench.py --steps 50 --warmup 5 --no-tree --no-cpu $EXTRA > $O/r02_bench_${TAG}_n1.json 2> $O/r02_bench_${TAG}_n1.err; echo "bench n1 rc=$?"; cut -c1-600 $O/r02_bench_${TAG}_n1.json; tail -3 $O/r02_bench_${TAG}_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu $EXTRA > $O/r02_bench_${TAG}_n2.json 2> $O/r02_bench_${TAG}_n2.err; echo "bench n2 rc=$?"; cut -c1-600 $O/r02_bench_${TAG}_n2.json; tail -3 $O/r02_bench_${TAG}_n2.err
