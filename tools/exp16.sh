#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) | tee gpurun_out/exp16_pytest.log
run() { n=$1; shift; env "$@" timeout 300 python bench.py --workload c3 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/exp16_$n.json; }
run pipe TPC_X=0
run pipe31 TPC_PIPE_BIN_CTAS=3 TPC_PIPE_FILL_CTAS=1
run nopipe TPC_PIPELINE=0
python - <<'PY'
import json
for n in ("pipe","pipe31","nopipe"):
    try:
        d=json.loads(open(f"gpurun_out/exp16_{n}.json").read())
        print(n, d["value"], d["ms_per_step"], d["stages_ms"], d["gpu_launches"], d["result"]["junctions"], d["result"]["records"], d["result"]["candidate_marks"])
    except Exception as e:
        print(n, "fail", e, open(f"gpurun_out/exp16_{n}.json").read()[:1500])
PY
