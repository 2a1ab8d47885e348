#!/usr/bin/env python
"""Multi-GPU parity check, run under torchrun on a box with >= 2 GPUs (or, with TPC_MGPU_BACKEND=gloo, with several
ranks sharing one GPU and exchanging over gloo):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_check.py

Every rank runs its hash-range shard (twopaco_b200.dist.sharded_run); rank 0 reassembles the
position slices and requires byte equality with the C oracle's image (whose ids use the same
first-appearance numbering).  Also run by tests/test_gpu_parity.py::test_multi_gpu_torchrun when
the box has >= 2 GPUs."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import oracle as O  # noqa: E402
from twopaco_b200 import api, synth  # noqa: E402
from twopaco_b200 import dist as tdist  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    backend = os.environ.get("TPC_MGPU_BACKEND", "nccl")
    if backend == "nccl":
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:   # ranks share the GPUs there are (one-GPU box: all on cuda:0) and exchange over gloo
        torch.cuda.set_device(local % torch.cuda.device_count())
        dist.init_process_group("gloo")
    cases = [dict(k=25, f=24, recs=synth.founder_family(91, 6, 3, 60_000, 0.01, n_runs=2) + [b"ACG", b""]),
             dict(k=63, f=22, recs=synth.founder_family(92, 4, 2, 50_000, 0.02, n_runs=1)),
             dict(k=9, f=20, recs=synth.reference_selftest_set(93))]
    for mode in ("direct", "binned"):
        os.environ["TPC_FILTER_MODE"] = mode
        os.environ["TPC_SLICE_LOG2"] = "12"
        for c in cases:
            g = api.pack_records(c["recs"])
            s = api.Session(k=c["k"], filter_bits=c["f"], shard_index=rank, shard_count=world)
            s.set_genome_host(g)
            info, out = tdist.sharded_run(s, g, rank, world)
            torch.cuda.synchronize()
            piece = out.to_host(info["slice_bytes"]).tobytes() if info["slice_bytes"] else b""
            pieces = [None] * world
            dist.all_gather_object(pieces, (info["slice_offset"], piece))
            if rank == 0:
                image = bytearray()
                for off, data in pieces:
                    assert off == len(image), (off, len(image))
                    image += data
                ref, nj, nm = O.find_junctions(c["recs"], c["k"])
                assert info["junctions"] == nj and info["records"] == nm, (info, nj, nm)
                assert bytes(image) == ref, f"multi-GPU image differs from the oracle (k={c['k']}, {mode})"
                print(f"mgpu ok: world={world} mode={mode} k={c['k']} junctions={nj} records={nm}", flush=True)
            s.close()
            # the same case from HOST buffers: 1/N upload + NCCL all-gather of the genome, image slices in host memory
            sh = tdist.host_shard(g.codes, g.n_mask, g.n_positions, g.rec_start, g.rec_len, rank, world)
            info, out_host, _ = tdist.sharded_run_host(sh, rank, world, c["k"], c["f"])
            pieces = [None] * world
            dist.all_gather_object(pieces, (info["slice_offset"], out_host[:info["slice_bytes"]].numpy().tobytes()))
            if rank == 0:
                image = b"".join(d for _, d in sorted(pieces))
                assert image == ref, f"multi-GPU host-buffer image differs from the oracle (k={c['k']}, {mode})"
                print(f"mgpu host-buffer ok: world={world} mode={mode} k={c['k']}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
