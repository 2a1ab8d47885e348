#!/bin/bash
# round 2, GPU call 15 (1 GPU): record scratch kept between the calls of a session -- full GPU suite, step-time jitter over 5 runs
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/r2c15_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2c15_pytest.log
for i in 1 2 3 4; do
  timeout 200 python bench.py --steps 6 --warmup 2 --no-cpu-baseline --no-verify --no-e2e --no-probe > $O/r2c15_keep_$i.json 2>> $O/r2c15.err
done
TPC_KEEP_SCRATCH=0 timeout 200 python bench.py --steps 6 --warmup 2 --no-cpu-baseline --no-verify --no-e2e --no-probe > $O/r2c15_nokeep_1.json 2>> $O/r2c15.err
TPC_KEEP_SCRATCH=0 timeout 200 python bench.py --steps 6 --warmup 2 --no-cpu-baseline --no-verify --no-e2e --no-probe > $O/r2c15_nokeep_2.json 2>> $O/r2c15.err
echo done
