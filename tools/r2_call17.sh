#!/bin/bash
# round 2, GPU call 17 (2 GPUs): reduce-scatter beside the index build (Python and C++ drivers), cached memory budget -- multi-GPU tests, C3 at N = 2
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 python -m pytest tests -m gpu -q -x -k "multi_gpu or cxx or cli_uses or multi_rank" > $O/r2c17_pytest_multi.log 2>&1; echo "pytest rc=$?" >> $O/r2c17_pytest_multi.log
timeout 300 $TR --master-port 29561 bench.py --gpus 2 --steps 4 --warmup 2 --no-cpu-baseline --no-probe > $O/r2c17_bench_c3_n2.json 2> $O/r2c17_bench_c3_n2.err
echo done
