#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/probe.py 27 28 29 30 > gpurun_out/exp2_probe.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/exp2_launches_c3_r3.csv \
  python bench.py --workload c3 --rounds 3 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/exp2_bench.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open("gpurun_out/exp2_launches_c3_r3.csv") if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[vi].replace(",", "")); u = r[ui]
    ms = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v if u in ("ms","msecond") else v*1e3
    n = r[ki].split("(")[0]
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += ms
for n, (c, ms) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{ms:10.3f} ms {c:6d} x  {n}")
PY
cat gpurun_out/exp2_probe.log
