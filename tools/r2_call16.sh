#!/bin/bash
# round 2, GPU call 16 (1 GPU): process-wide host trace of 8 steps at C3 (where does the step-to-step jitter of round 0 come from?)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
TPC_VERBOSE=1 timeout 300 python bench.py --steps 8 --warmup 2 --no-cpu-baseline --no-verify --no-e2e --no-probe > $O/r2c16_trace.json 2> $O/r2c16_trace.err
TPC_VERBOSE=1 TPC_BENCH_CLOCK_MS=0 timeout 300 python bench.py --steps 8 --warmup 2 --no-cpu-baseline --no-verify --no-e2e --no-probe > $O/r2c16_trace_noclock.json 2> $O/r2c16_trace_noclock.err
echo done
