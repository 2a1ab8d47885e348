#!/bin/bash
# round 2, GPU call 19 (1 GPU): final code -- full GPU suite, smoke, C3 bench line (all legs), launch list at C3
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/r2c19_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2c19_pytest.log
timeout 120 python __graft_entry__.py smoke > $O/r2c19_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r2c19_smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > $O/r2c19_bench_c3.json 2> $O/r2c19_bench_c3.err
B="python bench.py --no-e2e --no-verify --no-probe --no-cpu-baseline --steps 1 --warmup 0"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r2c19_launches_c3.csv $B > $O/r2c19_launches.log 2>&1
echo done
