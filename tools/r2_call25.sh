#!/bin/bash
# round 2, GPU call 25 (1 GPU, the last 2 minutes of the budget): launch list of the first sub-round of a C3 step, final code (queued marks)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
B="python bench.py --no-e2e --no-verify --no-probe --no-cpu-baseline --steps 1 --warmup 0"
TPC_BENCH_CLOCK_MS=0 timeout 80 ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file $O/r2c25_launches_c3.csv $B > $O/r2c25_launches.log 2>&1
echo done
