#!/bin/bash
# Final round-1 evidence on one B200: parity tests, the default bench (both arms), launch list (C3), ncu --set full (C2).
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) | tee gpurun_out/fin_pytest.log
timeout 600 python bench.py 2>gpurun_out/fin_bench.err | tail -1 > gpurun_out/fin_bench_c3.json
timeout 400 python bench.py --impl reference --steps 1 --warmup 1 2>>gpurun_out/fin_bench.err | tail -1 > gpurun_out/fin_bench_ref.json
timeout 300 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu-baseline 2>>gpurun_out/fin_bench.err | tail -1 > gpurun_out/fin_bench_c2.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/fin_launches_c3.csv \
    python bench.py --workload c3 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/fin_ncu_c3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -f -o gpurun_out/fin_prof_c2 \
    -k regex:'k_(bin|own|apply|insert|emit|classify|build)' -c 60 \
    python bench.py --workload c2 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/fin_ncu_c2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -f -o gpurun_out/fin_prof_c2_sim8 \
    -k regex:'k_(bin_list|own|insert)' -c 4 \
    python bench.py --workload c2 --sim-world 8 --steps 1 --warmup 0 > gpurun_out/fin_ncu_c2_sim8.log 2>&1
python - <<'PY'
import json
for n in ("c3","ref","c2"):
    try:
        d=json.loads(open(f"gpurun_out/fin_bench_{n}.json").read())
        print(n, d.get("value"), d.get("ms_per_step"), d.get("stages_ms"), d.get("gpu_launches"), d.get("e2e"), d.get("cpu_baseline"), d.get("roofline",{}).get("frac"), d.get("roofline",{}).get("random_access"))
    except Exception as e:
        print(n, "fail", e, open(f"gpurun_out/fin_bench_{n}.json").read()[:1500])
PY
tail -5 gpurun_out/fin_bench.err
