#!/bin/bash
# round 2, GPU call 4b (2 GPUs): the bench line at N = 2 with the e2e leg through tpc_multi_junctions_host
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --steps 3 --warmup 2 > $O/r2c4b_bench_c3_n2.json 2> $O/r2c4b_bench_c3_n2.err
echo done
