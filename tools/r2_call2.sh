#!/bin/bash
# round 2, GPU call 2 (2 GPUs): full GPU suite incl. the NCCL / C++ multi-GPU tests, new defaults at C3 on 1 and 2 GPUs
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/r2c2_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2c2_pytest.log
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > $O/r2c2_bench_c3_n1.json 2> $O/r2c2_bench_c3_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 2 > $O/r2c2_bench_c3_n2.json 2> $O/r2c2_bench_c3_n2.err
timeout 900 python tools/mgpu_cxx_check.py c3 2 > $O/r2c2_mgpu_cxx.json 2> $O/r2c2_mgpu_cxx.err
echo done
