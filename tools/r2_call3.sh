#!/bin/bash
# round 2, GPU call 3 (1 GPU): regressions of call 2 undone? timeline, zeroing knob, mark list at 4 / 8 shards
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/r2c3_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2c3_pytest.log
B="python bench.py --no-e2e --no-verify --no-probe --no-cpu-baseline --steps 2 --warmup 1"
TPC_VERBOSE=1 timeout 300 $B > $O/r2c3_c3_default.json 2> $O/r2c3_c3_default.err
TPC_QUERY_ZERO=0 timeout 300 $B > $O/r2c3_c3_nozero.json 2>&1
for w in 4 8; do
  TPC_VERBOSE=1 timeout 300 python bench.py --sim-world $w --steps 2 --warmup 1 > $O/r2c3_sim$w.json 2> $O/r2c3_sim$w.err
  TPC_MARK_LIST=0 timeout 300 python bench.py --sim-world $w --steps 2 --warmup 1 > $O/r2c3_sim${w}_nolist.json 2>&1
done
timeout 1200 python tools/cli_vs_reference.py c2 > $O/r2c3_cli_c2.json 2> $O/r2c3_cli_c2.err
echo done
