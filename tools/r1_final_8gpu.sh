#!/bin/bash
# Final round-1 evidence on 8 B200: multi-GPU parity, scaling N=8/4 (with e2e), C4 (k=63, k=127) at N=8
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_check.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -14 > gpurun_out/fin8_mgpu.log
tail -3 gpurun_out/fin8_mgpu.log
for n in 8 4; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 3 --warmup 3 2>gpurun_out/fin8_bench_n$n.err | tail -1 > gpurun_out/fin8_bench_c3_n$n.json
done
for w in c4k63 c4k127; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --workload $w --steps 2 --warmup 2 --no-e2e 2>gpurun_out/fin8_bench_$w.err | tail -1 > gpurun_out/fin8_bench_${w}_n8.json
done
python - <<'PY'
import json
for n in ("c3_n8","c3_n4","c4k63_n8","c4k127_n8"):
    try:
        d=json.loads(open(f"gpurun_out/fin8_bench_{n}.json").read())
        print(n, d["value"], d["ms_per_step"], d["stages_ms"], d.get("untimed_ms"), d["result"], d.get("e2e"))
    except Exception as e:
        print(n, "fail", e, open(f"gpurun_out/fin8_bench_{n}.json").read()[:800])
PY
