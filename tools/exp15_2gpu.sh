#!/bin/bash
# 2 GPUs: multi-GPU parity (device and host-buffer paths) + bench with e2e
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_check.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -20 | tee gpurun_out/exp15_mgpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 2 --warmup 1 2>gpurun_out/exp15_bench.err | tail -1 > gpurun_out/exp15_bench_n2.json
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/exp15_bench_n2.json").read())
    print(d["value"], d["ms_per_step"], d["stages_ms"], d["result"], d.get("e2e"))
except Exception as e:
    print("fail", e, open("gpurun_out/exp15_bench_n2.json").read()[:1500]); print(open("gpurun_out/exp15_bench.err").read()[-3000:])
PY
