#!/bin/bash
# round 2, GPU call 18 (8 GPUs): C3 at N = 8 with the final code (ownership fused into k_bin_list, cached memory budget, sparse n-mask upload)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29571 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline --no-probe > $O/r2c18_bench_c3_n8.json 2> $O/r2c18_bench_c3_n8.err
echo done
