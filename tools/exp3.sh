#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/exp3_tests.log 2>&1
for cfg in "c3 0" "c3 1" "c2 0"; do
  set -- $cfg
  TPC_SUBROUNDS=$2 timeout 300 python bench.py --workload $1 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/exp3_$1_$2.json
done
python - <<'PY'
import json
for n in ("c3_0","c3_1","c2_0"):
    try:
        d=json.loads(open(f"gpurun_out/exp3_{n}.json").read())
        print(n, d["value"], d["ms_per_step"], d["stages_ms"], d["result"]["candidate_marks"], d["result"]["candidate_kmers"], d["gpu_launches"])
    except Exception as e:
        print(n, "fail", e, open(f"gpurun_out/exp3_{n}.json").read()[:1500])
PY
cat gpurun_out/exp3_tests.log
