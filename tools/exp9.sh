#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) | tee gpurun_out/exp9_pytest.log
for r in 0 8; do
  TPC_BIN_R=$r timeout 300 python bench.py --workload c3 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/exp9_c3_r$r.json
done
python - <<'PY'
import json
for n in (0,8):
    try:
        d=json.loads(open(f"gpurun_out/exp9_c3_r{n}.json").read())
        print(n, d["value"], d["ms_per_step"], d["stages_ms"], d["gpu_launches"], d["result"])
    except Exception as e:
        print(n, "fail", e, open(f"gpurun_out/exp9_c3_r{n}.json").read()[:1500])
PY
