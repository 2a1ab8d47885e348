#!/bin/bash
# round 2, GPU call 5a (8 GPUs, tight timeouts): C3 at 8 GPUs (value + e2e through the C ABI), multi-GPU tests, C4 k=63
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29551 bench.py --gpus 8 --steps 3 --warmup 2 > $O/r2c5_bench_c3_n8.json 2> $O/r2c5_bench_c3_n8.err
timeout 200 python -m pytest tests -m gpu -q -x -k "multi_gpu or cxx or cli_uses" > $O/r2c5_pytest_multi.log 2>&1; echo "pytest rc=$?" >> $O/r2c5_pytest_multi.log
timeout 150 $TR --master-port 29552 bench.py --gpus 8 --workload c4k63 --steps 2 --warmup 1 --no-e2e --no-verify > $O/r2c5_bench_c4k63_n8.json 2> $O/r2c5_bench_c4k63_n8.err
echo done
