#!/bin/bash
# round 2, GPU call 1: tests after the refactor, probes, baseline C3 timeline, tuning knobs, C2 vs the reference
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $O/r2c1_gpu.txt
nproc >> $O/r2c1_gpu.txt; free -g >> $O/r2c1_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2c1_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2c1_pytest.log
timeout 600 python tools/probe.py hbm 36 slice > $O/r2c1_probe.json 2> $O/r2c1_probe.err
timeout 600 ncu --set full --clock-control none -k regex:k_probe_slice -c 8 -o $O/r2c1_probe_ncu -f python -c "
import sys; sys.path.insert(0,'.')
from tools import benchutil
for mode in (1,2):
    print(benchutil.slice_probe(26, 3, 32<<20, 7, mode, 4, 4))
" > $O/r2c1_probe_ncu.log 2>&1
B="python bench.py --no-e2e --no-verify --no-probe --no-cpu-baseline --steps 2 --warmup 1"
TPC_VERBOSE=1 timeout 900 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > $O/r2c1_bench_c3.json 2> $O/r2c1_bench_c3.err
TPC_VERBOSE=1 timeout 300 python bench.py --sim-world 8 --steps 2 --warmup 1 > $O/r2c1_sim8.json 2> $O/r2c1_sim8.err
TPC_PIPELINE=0 timeout 300 $B > $O/r2c1_v_nopipe.json 2>&1
TPC_SLICE_LOG2=25 timeout 300 $B > $O/r2c1_v_slice25.json 2>&1
TPC_PIPE_BIN_CTAS=2 TPC_PIPE_FILL_CTAS=2 timeout 300 $B > $O/r2c1_v_pipe22.json 2>&1
TPC_SUBROUNDS=4 timeout 300 $B > $O/r2c1_v_sub4.json 2>&1
TPC_APPLY_CTAS=2 TPC_PIPELINE=0 timeout 300 $B > $O/r2c1_v_apply2.json 2>&1
timeout 600 ncu --set full --clock-control none -k regex:k_apply_ --launch-skip 380 --launch-count 8 -o $O/r2c1_apply_c3_ncu -f $B --steps 1 --warmup 0 > $O/r2c1_apply_c3_ncu.log 2>&1
timeout 300 python bench.py --workload c2 --steps 3 --warmup 2 --no-cpu-baseline > $O/r2c1_bench_c2.json 2> $O/r2c1_bench_c2.err
timeout 1200 python tools/cli_vs_reference.py c2 > $O/r2c1_cli_c2.json 2> $O/r2c1_cli_c2.err
echo done
