#!/bin/bash
# round 2, GPU call 23 (1 GPU): k_apply_query with the warp-aggregated mark queue (TPC_QUERY_AGG=1) -- binned-path parity tests,
# C3 with the digest check against the direct path, A/B against the inline marks on the same box
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
TPC_QUERY_AGG=1 timeout 400 python -m pytest tests -m gpu -q -x -k "binned or sub_rounds or mark_list or headline or boundaries or sharded or multi_rank or golden" > $O/r2c23_pytest_agg.log 2>&1; echo "pytest rc=$?" >> $O/r2c23_pytest_agg.log
TPC_QUERY_AGG=1 timeout 200 python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-e2e --no-probe > $O/r2c23_bench_agg1.json 2> $O/r2c23_bench_agg1.err
TPC_QUERY_AGG=0 timeout 200 python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-e2e --no-probe --no-verify > $O/r2c23_bench_agg0.json 2> $O/r2c23_bench_agg0.err
echo done
