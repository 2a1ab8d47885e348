#!/bin/bash
# round 2, GPU call 9 (1 GPU): compute-sanitizer (memcheck, racecheck, initcheck-free) over small parity cases of every code path
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
SEL="test_binned_rounds_and_shards or test_mark_list_variants or test_windowed_run_edge_cases or test_empty_and_degenerate or (test_sub_rounds_share_one_ownership_scan and family_k25 and 3-1) or (test_binned_filter_passes_match_golden and family_k63 and 13-0) or test_get_id_surface or (test_gpu_graphdump and example)"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 99 --log-file $O/r2c9_memcheck.log python -m pytest tests -m gpu -q -x -k "$SEL" > $O/r2c9_memcheck_pytest.log 2>&1; echo "memcheck rc=$?" >> $O/r2c9_memcheck_pytest.log
SEL2="test_binned_rounds_and_shards or (test_mark_list_variants and family_k25) or test_windowed_run_edge_cases"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 99 --log-file $O/r2c9_racecheck.log python -m pytest tests -m gpu -q -x -k "$SEL2" > $O/r2c9_racecheck_pytest.log 2>&1; echo "racecheck rc=$?" >> $O/r2c9_racecheck_pytest.log
tail -5 $O/r2c9_memcheck.log $O/r2c9_racecheck.log
echo done
