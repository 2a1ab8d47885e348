#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/exp1_tests.log 2>&1
for r in 1 2 3 4; do
  timeout 300 python bench.py --workload c3 --rounds $r --steps 2 --warmup 1 --no-e2e --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/exp1_r$r.json
done
python - <<'PY'
import json
for r in (1,2,3,4):
    try:
        d=json.loads(open(f"gpurun_out/exp1_r{r}.json").read())
        print(r, d["value"], d["ms_per_step"], d["stages_ms"], d["result"]["candidate_marks"], d["result"]["candidate_kmers"], d["gpu_launches"])
    except Exception as e:
        print(r, "fail", e, open(f"gpurun_out/exp1_r{r}.json").read()[:500])
PY
cat gpurun_out/exp1_tests.log
