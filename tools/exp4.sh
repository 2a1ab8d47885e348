#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/exp4_tests.log 2>&1
timeout 300 python bench.py --workload c3 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/exp4_c3.json
timeout 300 python bench.py --workload c2 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/exp4_c2.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/exp4_launches_c3.csv \
  python bench.py --workload c3 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/exp4_ncu_bench.log 2>&1
python tools/agg_launches.py gpurun_out/exp4_launches_c3.csv
python - <<'PY'
import json
for n in ("c3","c2"):
    try:
        d=json.loads(open(f"gpurun_out/exp4_{n}.json").read())
        print(n, d["value"], d["ms_per_step"], d["stages_ms"], d["result"]["candidate_marks"], d["result"]["candidate_kmers"], d["gpu_launches"])
    except Exception as e:
        print(n, "fail", e, open(f"gpurun_out/exp4_{n}.json").read()[:1500])
PY
cat gpurun_out/exp4_tests.log
