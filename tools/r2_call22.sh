#!/bin/bash
# round 2, GPU call 22 (1 GPU): final state -- full GPU suite, smoke, default bench invocation
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/r2c22_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2c22_pytest.log
timeout 120 python __graft_entry__.py smoke > $O/r2c22_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r2c22_smoke.log
timeout 600 python bench.py > $O/r2c22_bench_default.json 2> $O/r2c22_bench_default.err
echo done
