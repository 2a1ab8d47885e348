#!/bin/bash
# round 2, GPU call 5b (8 GPUs, tight timeout): C5 -- 310 Gbp streamed from host memory, position-windowed
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
free -g > $O/r2c5_c5_mem.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
TPC_VERBOSE=1 timeout 560 $TR --master-port 29554 bench.py --gpus 8 --workload c5 --steps 1 --warmup 0 > $O/r2c5_bench_c5_n8.json 2> $O/r2c5_bench_c5_n8.err
echo "rc=$?" >> $O/r2c5_c5_mem.txt
echo done
