#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for sl in 26 25; do
  TPC_SLICE_LOG2=$sl timeout 300 python bench.py --workload c3 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/exp6_c3_$sl.json
done
for sl in 26 25 24; do
  TPC_SLICE_LOG2=$sl timeout 300 python bench.py --workload c2 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/exp6_c2_$sl.json
done
python - <<'PY'
import json
for n in ("c3_26","c3_25","c2_26","c2_25","c2_24"):
    try:
        d=json.loads(open(f"gpurun_out/exp6_{n}.json").read())
        print(n, d["value"], d["ms_per_step"], d["stages_ms"], round(sum(d["stages_ms"].values()),1), d["gpu_launches"])
    except Exception as e:
        print(n, "fail", e, open(f"gpurun_out/exp6_{n}.json").read()[:1500])
PY
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 )
