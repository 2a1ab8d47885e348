#!/bin/bash
# round 2, GPU call 24 (1 GPU): final state (queued marks by default) -- full GPU suite, default bench invocation
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/r2c24_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2c24_pytest.log
timeout 400 python bench.py > $O/r2c24_bench_default.json 2> $O/r2c24_bench_default.err
echo done
