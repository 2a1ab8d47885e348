#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python bench.py --workload c2 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/exp5_c2.json
timeout 300 python bench.py --workload c3 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/exp5_c3.json
ncu --set full --import-source on --clock-control none -k regex:"k_insert|k_bin_list|k_apply_fill|k_apply_query|k_bin|k_emit" -c 40 -f -o gpurun_out/exp5_c2_full \
  python bench.py --workload c2 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/exp5_ncu_c2.log 2>&1
TPC_SUBROUNDS=3 ncu --set full --import-source on --clock-control none -k regex:"k_insert|k_bin_list|k_own" -c 3 -f -o gpurun_out/exp5_c2s3_full \
  python bench.py --workload c2 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/exp5_ncu_c2s3.log 2>&1
python - <<'PY'
import json
for n in ("c3","c2"):
    try:
        d=json.loads(open(f"gpurun_out/exp5_{n}.json").read())
        print(n, d["value"], d["ms_per_step"], d["stages_ms"], d["result"]["candidate_marks"], d["result"]["candidate_kmers"], d["gpu_launches"])
    except Exception as e:
        print(n, "fail", e, open(f"gpurun_out/exp5_{n}.json").read()[:1500])
PY
ls -la gpurun_out/
