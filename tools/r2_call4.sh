#!/bin/bash
# round 2, GPU call 4 (2 GPUs): windowed sessions (1 and 2 GPUs), C++ multi-GPU timeline
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/r2c4_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2c4_pytest.log
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --workload c5mini --steps 2 --warmup 1 > $O/r2c4_c5mini_n1.json 2> $O/r2c4_c5mini_n1.err
TPC_VERBOSE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --workload c5mini --steps 2 --warmup 1 > $O/r2c4_c5mini_n2.json 2> $O/r2c4_c5mini_n2.err
TPC_VERBOSE=1 timeout 900 python tools/mgpu_cxx_check.py c3 2 > $O/r2c4_mgpu_cxx.json 2> $O/r2c4_mgpu_cxx.err
echo done
