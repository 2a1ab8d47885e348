#!/bin/bash
# 8 GPUs: multi-GPU parity + bench at N=8 and N=4 (with e2e)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_check.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -14 | tee gpurun_out/exp17_mgpu.log
for n in 8 4; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 3 --warmup 2 2>gpurun_out/exp17_bench_n$n.err | tail -1 > gpurun_out/exp17_bench_n$n.json
done
python - <<'PY'
import json
for n in (8,4):
    try:
        d=json.loads(open(f"gpurun_out/exp17_bench_n{n}.json").read())
        print(n, d["value"], d["ms_per_step"], d["stages_ms"], d["result"], d.get("e2e"))
    except Exception as e:
        print("fail", e, open(f"gpurun_out/exp17_bench_n{n}.json").read()[:1500]); print(open(f"gpurun_out/exp17_bench_n{n}.err").read()[-3000:])
PY
