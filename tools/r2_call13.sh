#!/bin/bash
# round 2, GPU call 13 (1 GPU): sparse n-mask upload (e2e A/B on one box), host timeline of find_candidates at C3
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests -m gpu -q -x -k "sparse_n_mask or host_buffer or empty_and_degenerate or filter_shape or abi" > $O/r2c13_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2c13_pytest.log
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-verify > $O/r2c13_bench_sparse.json 2> $O/r2c13_bench_sparse.err
TPC_SPARSE_MASK=0 timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-verify > $O/r2c13_bench_dense.json 2> $O/r2c13_bench_dense.err
TPC_VERBOSE=1 timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-verify --no-e2e > $O/r2c13_verbose.json 2> $O/r2c13_verbose.err
echo done
