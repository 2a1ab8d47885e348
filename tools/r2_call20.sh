#!/bin/bash
# round 2, GPU call 20 (2 GPUs): timeline of the C++ multi-GPU driver (e2e, host buffers in / out) at C3
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
TPC_VERBOSE=1 timeout 300 $TR --master-port 29581 bench.py --gpus 2 --steps 2 --warmup 1 --no-cpu-baseline --no-probe --no-verify > $O/r2c20_bench_c3_n2.json 2> $O/r2c20_bench_c3_n2.err
echo done
