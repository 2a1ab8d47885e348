#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { # name env...
  n=$1; shift
  env "$@" timeout 300 python bench.py --workload c3 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/exp12_$n.json
}
run base TPC_X=0
run s6 TPC_SUBROUNDS=6
run s6half TPC_SUBROUNDS=6 TPC_BIN_CTAS=2 TPC_APPLY_CTAS=2
run s4half TPC_SUBROUNDS=4 TPC_BIN_CTAS=2 TPC_APPLY_CTAS=2
python - <<'PY'
import json
for n in ("base","s6","s6half","s4half"):
    try:
        d=json.loads(open(f"gpurun_out/exp12_{n}.json").read())
        print(n, d["value"], d["ms_per_step"], d["stages_ms"], d["gpu_launches"], d["result"]["candidate_marks"])
    except Exception as e:
        print(n, "fail", e, open(f"gpurun_out/exp12_{n}.json").read()[:1500])
PY
