#!/bin/bash
# round 2, GPU call 5 (8 GPUs): C3 scaling point, C++ multi-GPU e2e, C4, C5 (windowed, streamed from host), multi-GPU tests
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=index,name,memory.used --format=csv > $O/r2c5_gpus.txt; free -g >> $O/r2c5_gpus.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29551 bench.py --gpus 8 --steps 3 --warmup 2 > $O/r2c5_bench_c3_n8.json 2> $O/r2c5_bench_c3_n8.err
timeout 900 python -m pytest tests -m gpu -q -x -k "multi or cxx or torchrun or cli_uses" > $O/r2c5_pytest_multi.log 2>&1; echo "pytest rc=$?" >> $O/r2c5_pytest_multi.log
timeout 600 $TR --master-port 29552 bench.py --gpus 8 --workload c4k63 --steps 2 --warmup 1 --no-e2e > $O/r2c5_bench_c4k63_n8.json 2> $O/r2c5_bench_c4k63_n8.err
timeout 600 $TR --master-port 29553 bench.py --gpus 8 --workload c4k127 --steps 2 --warmup 1 --no-e2e > $O/r2c5_bench_c4k127_n8.json 2> $O/r2c5_bench_c4k127_n8.err
TPC_VERBOSE=1 timeout 1500 $TR --master-port 29554 bench.py --gpus 8 --workload c5 --steps 1 --warmup 1 > $O/r2c5_bench_c5_n8.json 2> $O/r2c5_bench_c5_n8.err
echo done
