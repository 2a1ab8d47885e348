#!/bin/bash
# round 2, GPU call 11 (1 GPU): ownership fused into k_bin_list (single-round GPUs), list-driven direct passes at 1/32 ownership
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q > $O/r2c11_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2c11_pytest.log
for w in 4 8; do
  timeout 200 python bench.py --sim-world $w --steps 3 --warmup 1 > $O/r2c11_sim${w}_fused.json 2>&1
  TPC_FUSED_OWN=0 timeout 200 python bench.py --sim-world $w --steps 3 --warmup 1 > $O/r2c11_sim${w}_planes.json 2>&1
done
timeout 300 python - > $O/r2c11_windowed32.json 2> $O/r2c11_windowed32.err <<'PY'
import json, os, sys, time
sys.path.insert(0, ".")
from tools import benchutil
from twopaco_b200 import api
dg = benchutil.synth_family_device(0x4831, 12, 4, 8_000_000, 0.001)
host = dg.to_host()
out = {}
os.environ["TPC_WINDOW_TILES"] = "4096"
for flag, name in (("1", "list"), ("0", "inline")):
    os.environ["TPC_DIRECT_LIST"] = flag
    best = None
    for _ in range(2):
        t0 = time.perf_counter()
        img, st = api.junctions_host(host, k=31, filter_bits=32, q=5, rounds=32)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    out[name] = {"wall_ms": round(best * 1e3, 1), "ms_fill_32_passes": round(st.ms_fill, 1), "ms_query_insert_32_passes": round(st.ms_query, 1),
                 "digest": api.image_digest_host(img)}
out["equal"] = out["list"]["digest"] == out["inline"]["digest"]
print(json.dumps(out))
PY
echo done
