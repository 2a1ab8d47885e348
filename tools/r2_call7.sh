#!/bin/bash
# round 2, GPU call 7 (1 GPU): L2 persisting window over the slice (knob), a more skewed input (re-bin path timed)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
B="python bench.py --no-e2e --no-verify --no-probe --no-cpu-baseline --steps 3 --warmup 2"
TPC_L2_PERSIST=1 TPC_VERBOSE=1 timeout 300 $B > $O/r2c7_c3_persist.json 2> $O/r2c7_c3_persist.err
timeout 300 $B > $O/r2c7_c3_default.json 2> $O/r2c7_c3_default.err
TPC_L2_PERSIST=1 timeout 200 python bench.py --sim-world 8 --steps 2 --warmup 1 > $O/r2c7_sim8_persist.json 2>&1
timeout 400 python tools/skew_bench.py 40000000 0.15 > $O/r2c7_skew15.json 2> $O/r2c7_skew15.err
echo done
