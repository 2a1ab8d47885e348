#!/bin/bash
# round 2, GPU call 6 (1 GPU): final code -- full GPU suite, skewed input, C3 / C2 bench lines, launch list + ncu captures at C3
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q > $O/r2c6_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2c6_pytest.log
timeout 400 python tools/skew_bench.py 40000000 0.05 > $O/r2c6_skew.json 2> $O/r2c6_skew.err
timeout 600 python bench.py --steps 5 --warmup 3 > $O/r2c6_bench_c3.json 2> $O/r2c6_bench_c3.err
timeout 300 python bench.py --filter-mode direct --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $O/r2c6_bench_c3_direct.json 2> $O/r2c6_bench_c3_direct.err
timeout 300 python bench.py --workload c2 --steps 5 --warmup 3 > $O/r2c6_bench_c2.json 2> $O/r2c6_bench_c2.err
B="python bench.py --no-e2e --no-verify --no-probe --no-cpu-baseline --steps 1 --warmup 0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r2c6_launches_c3.csv $B > $O/r2c6_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_apply_fill|k_apply_query" --launch-skip 300 --launch-count 4 -o $O/r2c6_ncu_apply -f $B > $O/r2c6_ncu_apply.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_bin_list|k_own|k_insert|k_emit" -c 6 -o $O/r2c6_ncu_bin -f $B > $O/r2c6_ncu_bin.log 2>&1
echo done
