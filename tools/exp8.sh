#!/bin/bash
# slice-size experiment (cross-die L2 effect) + ncu of the sharded binning kernels at 1/3 ownership
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for sl in 25 27 26; do
  TPC_SLICE_LOG2=$sl timeout 300 python bench.py --workload c3 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/exp8_c3_s$sl.json
done
python - <<'PY'
import json
for n in (25,27,26):
    try:
        d=json.loads(open(f"gpurun_out/exp8_c3_s{n}.json").read())
        print(n, d["value"], d["ms_per_step"], d["stages_ms"], d["gpu_launches"])
    except Exception as e:
        print(n, "fail", e, open(f"gpurun_out/exp8_c3_s{n}.json").read()[:1500])
PY
timeout 600 ncu --set full --clock-control none --import-source on -f -o gpurun_out/exp8_prof_c2_sim3 \
    -k regex:'k_(bin_list|own)' -c 4 \
    python bench.py --workload c2 --sim-world 3 --steps 1 --warmup 0 > gpurun_out/exp8_ncu.log 2>&1
tail -3 gpurun_out/exp8_ncu.log
