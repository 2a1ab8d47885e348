#!/bin/bash
# round 2, GPU call 8 (1 GPU): ranking probe (shared atomics vs match / ballot groups)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 200 python - > $O/r2c8_rank_probe.json 2> $O/r2c8_rank_probe.err <<'PY'
import json, sys
sys.path.insert(0, ".")
from tools import benchutil
out = {}
for buckets in (128, 256, 16):
    for mode, name in ((0, "shared_atomic_per_record"), (1, "match_any+warp_private_hist"), (2, "ballots+warp_private_hist"), (3, "match_any+atomic_per_group")):
        out[f"{name}_b{buckets}"] = round(benchutil.rank_probe(mode, buckets, 2000) / 1e9, 2)
print(json.dumps(out, indent=1))
PY
echo done
