#!/bin/bash
# round 2, GPU call 10 (1 GPU): list-driven direct kernels -- full GPU suite, C3 direct-path line, c5mini
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q > $O/r2c10_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2c10_pytest.log
timeout 300 python bench.py --workload c5mini --steps 3 --warmup 1 > $O/r2c10_c5mini_n1.json 2> $O/r2c10_c5mini_n1.err
TPC_DIRECT_LIST=0 timeout 300 python bench.py --workload c5mini --steps 3 --warmup 1 --no-verify > $O/r2c10_c5mini_n1_inline.json 2> $O/r2c10_c5mini_n1_inline.err
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > $O/r2c10_bench_c3.json 2> $O/r2c10_bench_c3.err
echo done
