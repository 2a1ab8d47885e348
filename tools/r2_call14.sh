#!/bin/bash
# round 2, GPU call 14 (1 GPU): does the nvidia-smi sampler disturb the step (50 ms / 200 ms / off, same box)?  e2e timeline.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
for ms in 50 200 0 50 0; do
  TPC_BENCH_CLOCK_MS=$ms timeout 200 python bench.py --steps 6 --warmup 2 --no-cpu-baseline --no-verify --no-e2e --no-probe > $O/r2c14_clock${ms}_$RANDOM.json 2>> $O/r2c14.err
done
TPC_VERBOSE=1 timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-verify --no-probe > $O/r2c14_e2e_verbose.json 2> $O/r2c14_e2e_verbose.err
echo done
