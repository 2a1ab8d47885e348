#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) | tee gpurun_out/exp18_pytest.log
timeout 300 python bench.py --workload c3 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/exp18_c3.json
timeout 300 python bench.py --workload c3 --sim-world 8 --steps 2 --warmup 1 2>&1 | tail -1 > gpurun_out/exp18_c3_sim8.json
timeout 300 python bench.py --workload c4k63 --sim-world 8 --steps 2 --warmup 1 2>&1 | tail -1 > gpurun_out/exp18_c4k63_sim8.json
python - <<'PY'
import json
for n in ("c3",):
    try:
        d=json.loads(open(f"gpurun_out/exp18_{n}.json").read())
        print(n, d["value"], d["ms_per_step"], d["stages_ms"], d["gpu_launches"], d["result"])
    except Exception as e:
        print(n, "fail", e, open(f"gpurun_out/exp18_{n}.json").read()[:1500])
print(open("gpurun_out/exp18_c3_sim8.json").read())
print(open("gpurun_out/exp18_c4k63_sim8.json").read())
PY
