#!/usr/bin/env python
"""bench.py -- input Gbp/s to the exact junction set (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|dev] [--impl reference]

One "step" = one full pass of the junction-finding hot path (filter fill, candidate query,
exact pass, id index, ordered emit) over one synthetic genome set.

* value      : whole-job throughput, packed genome already resident in HBM, image left in HBM.
* e2e        : same metric through the C-ABI call with HOST buffers (tpc_junctions_host):
               H2D of the packed genome and D2H of the de_bruijn.bin image inside the timed region.
* roofline   : the dominant kernel (candidate query), algorithmic bytes / CUDA-event time.
* cpu_baseline: the UNMODIFIED reference (oracle/_ref/twopaco) on a bounded sample of the same
               workload with all host cores; the sample's output is also compared (canonical
               stream) with ours -- that, not the oracle, is what `parity_on_sample` reports.
Multi-GPU (torchrun): hash-range shards, NCCL all-gather of junction lists + OR-reduce of the
candidate masks, position-sharded emit.  Strong scaling: the genome set is fixed as N grows.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # SURVEY.md 8(d) / BASELINE.json configs[2]: 7 x 24 x 129,166,667 bp, p = 0.001, k = 25, -f 36
    "c3": dict(name="C3: 7 synthetic human-sized genomes (7x24x129166667 bp, 0.1% divergence), k=25 -f 36 -q 5",
               seed=0x4855, genomes=7, records=24, length=129_166_667, p=0.001, k=25, f=36, q=5, sample_bp=20_000_000),
    # configs[1]: 62 x 5 Mbp, p = 0.01, k = 25, -f 32
    "c2": dict(name="C2: 62 synthetic E. coli-like genomes (62x5 Mbp, 1% divergence), k=25 -f 32 -q 5",
               seed=0xEC01, genomes=62, records=1, length=5_000_000, p=0.01, k=25, f=32, q=5, sample_bp=400_000),
    # configs[3]: the same 7-genome set at k = 63 / k = 127 (2 / 4 words per k-mer), -f 37
    "c4k63": dict(name="C4: 7 synthetic human-sized genomes (7x24x129166667 bp, 0.1% divergence), k=63 -f 37 -q 5",
                  seed=0x4855, genomes=7, records=24, length=129_166_667, p=0.001, k=63, f=37, q=5, sample_bp=20_000_000),
    "c4k127": dict(name="C4: 7 synthetic human-sized genomes (7x24x129166667 bp, 0.1% divergence), k=127 -f 37 -q 5",
                   seed=0x4855, genomes=7, records=24, length=129_166_667, p=0.001, k=127, f=37, q=5, sample_bp=20_000_000),
    "dev": dict(name="dev: 7x2x4 Mbp, 0.1% divergence, k=25 -f 30 -q 5",
                seed=0xD0, genomes=7, records=2, length=4_000_000, p=0.001, k=25, f=30, q=5, sample_bp=500_000),
}


# (junctions, records, stubs) of the synthetic workloads: identical in every run of round 1 -- direct and binned filter
# passes, 1 / 2 / 4 / 8 GPUs, with and without sub-rounds and pipelining.  A run that deviates has lost or invented
# junctions (small parity tests can stay green while a race only shows at this scale), so the bench line says so.
EXPECTED_COUNTS = {"c3": (36118333, 248411004, 8), "c2": (3310265, 150000784, 22),
                   "c4k63": (34766769, 232728501, 19), "c4k127": (32624952, 208872372, 32)}


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self) -> dict:
        sm = [int(r[0]) for r in self.rows if len(r) >= 6 and r[0].isdigit()]
        mx = [int(r[1]) for r in self.rows if len(r) >= 6 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower() == "active"})
        return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the unmodified reference on a bounded sample
# ---------------------------------------------------------------------------------------------
def sample_records(dg, wl) -> list[bytes]:
    """Bounded sample of the workload: the first sample_bp bases of the first record of every genome."""
    return [dg.record_ascii(g * wl["records"], wl["sample_bp"]) for g in range(wl["genomes"])]


def run_reference_on(records: list[bytes], wl: dict, threads: int):
    from oracle import oracle as O
    with tempfile.TemporaryDirectory(prefix="tpc_bench_") as d:
        paths = []
        for i, r in enumerate(records):  # one FASTA per genome, as the reference is used
            p = os.path.join(d, f"g{i}.fa")
            O.write_fasta(p, [r], names=[f"g{i}_c0"])
            paths.append(p)
        t0 = time.perf_counter()
        img, log = O.run_reference(paths, wl["k"], min(wl["f"], 32), q=wl["q"], r=1, t=threads)
        dt = time.perf_counter() - t0
    return img, dt


def cpu_baseline(recs: list[bytes], wl: dict) -> dict:
    from oracle import oracle as O
    from twopaco_b200 import api
    if not O.have_reference():
        return {"value": None, "unit": "Gbp/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref/twopaco missing"}
    cores = os.cpu_count() or 1
    bp = sum(len(r) for r in recs)
    ref_img, dt = run_reference_on(recs, wl, cores)
    ours, _ = api.junctions_host(api.pack_records(recs), k=wl["k"], filter_bits=min(wl["f"], 32), q=wl["q"])
    return {"value": bp / dt / 1e9, "unit": "Gbp/s", "cores": cores, "kind": "reference",
            "sample": f"first {wl['sample_bp']} bp of record 0 of each of the {wl['genomes']} genomes ({bp} bp), "
                      f"-k {wl['k']} -f {min(wl['f'], 32)} -q {wl['q']} -t {cores}, wall {dt:.2f} s incl. FASTA parsing",
            "parity_on_sample": bool(O.canon_equal(bytes(ours), ref_img))}


# ---------------------------------------------------------------------------------------------
# one step of our arm
# ---------------------------------------------------------------------------------------------
class Runner:
    def __init__(self, wl, dg, rank, world):
        import torch
        from twopaco_b200 import api
        self.torch, self.api, self.wl, self.dg, self.rank, self.world = torch, api, wl, dg, rank, world
        self.session = None
        self.out = None
        self.last = {}

    def step(self):
        api, wl, dg = self.api, self.wl, self.dg
        from twopaco_b200 import dist as tdist
        if self.session is not None:
            self.session.close()
        s = api.Session(k=wl["k"], filter_bits=wl["f"], q=wl["q"], rounds=wl.get("rounds", 1),
                        shard_index=self.rank, shard_count=self.world)
        self.session = s
        dg.attach(s)
        self.last, self.out = tdist.sharded_run(s, dg, self.rank, self.world, self.out)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=os.environ.get("TPC_BENCH_WORKLOAD", "c3"), choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--sim-world", type=int, default=0,
                    help="kernel tuning aid: time only pass 1+2 of shard 0 of N on ONE GPU (not a bench value)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--rounds", type=int, default=0, help="-r of the run (default: the workload's, 1)")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.rounds:
        wl["rounds"] = args.rounds
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        if rank == 0:  # the other ranks exit without work
            reference_arm(args, wl, world)
        return

    import torch
    from twopaco_b200 import api
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    probe = {}
    if rank == 0 and not args.sim_world:   # random-access roofline probe while HBM is still empty
        for mode, name in ((0, "load32B"), (2, "load_condAtomicOr")):
            probe[name] = round(api.random_access_probe(wl["f"], mode, 1 << 30) / 1e9, 2)
    total_bp = wl["genomes"] * wl["records"] * wl["length"]
    dg = api.synth_family_device(wl["seed"], wl["genomes"], wl["records"], wl["length"], wl["p"], keep_ascii=(rank == 0))
    total_bp = dg.total_bp
    sample = sample_records(dg, wl) if rank == 0 else None   # bounded sample for the reference arm / parity check
    if dg.ascii is not None:                                  # the 1 byte/bp generator output is not an input of the path
        dg.ascii.close()
        dg.ascii = None

    if args.sim_world:
        for i in range(args.warmup + args.steps):
            s = api.Session(k=wl["k"], filter_bits=wl["f"], q=wl["q"], shard_index=0, shard_count=args.sim_world)
            dg.attach(s)
            s.find_candidates()
            st = s.stats()
            s.close()
        print(json.dumps({"sim_world": args.sim_world, "workload": args.workload,
                          "stages_ms": {k: round(getattr(st, k), 3) for k in ("ms_bin", "ms_bin_overlapped", "ms_fill", "ms_query", "ms_insert", "ms_classify")},
                          "bin_waves": st.bin_waves, "marks": st.candidate_marks, "candidate_kmers": st.candidate_kmers}))
        return

    runner = Runner(wl, dg, rank, world)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 0)):
        runner.step()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage_ms = {k: 0.0 for k in ("ms_bin", "ms_bin_overlapped", "ms_fill", "ms_query", "ms_insert", "ms_classify", "ms_index", "ms_emit")}
    launches = 0
    with ClockSampler(local_rank) as clocks:
        ev0.record()
        for _ in range(args.steps):
            runner.step()
            st = runner.session.stats()
            for k in stage_ms:
                stage_ms[k] += getattr(st, k)
            launches += st.kernel_launches
        ev1.record()
        barrier()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    st = runner.session.stats()

    # ---- roofline of the dominant filter-pass kernel ---------------------------
    # algorithmic HBM bytes, summed over the kernel's launches of one step:
    #   direct : k_fill  = 32 B (one filter sector) x owned k-mers + 0.375 B x positions (packed stream)
    #            k_query = the same + 1 bit/position of candidate mask
    #   binned : k_bin         = 0.375 B x positions + 12 B x records written        (per binning pass)
    #            k_apply_fill  = 8 B x records read + filter read once and written back once per wave
    #            k_apply_query = 8 B x records + 8 B x marks (position word + mask word) + filter read per wave
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    positions, recs, marks = st.positions, total_bp / world, st.candidate_marks
    filter_bytes = (1 << wl["f"]) / 8
    per = {k: v / args.steps for k, v in stage_ms.items()}
    if st.bin_waves:
        waves = st.bin_waves
        passes = 1 if waves == 1 else 2
        rounds_local = wl.get("rounds", 1) * max(st.sub_rounds, 1)          # hash sub-ranges this GPU runs in sequence
        if rounds_local * world > 1:
            # sharded binning: ONE ownership scan (k_own: stream in, P bit planes out) + per round a k_bin_list pass
            # (stream + planes in, 12 B per owned record out)
            planes = max(1, rounds_local.bit_length())
            bin_bytes = (0.375 + 0.125 * planes) * positions * (1 + passes * rounds_local) + passes * 12.0 * recs
        else:
            bin_bytes = passes * (0.375 * positions + 12.0 * recs)
        # pipelined rounds: the binning of round r+1 runs beside the fill of round r (ms_bin_overlapped, CUDA events on
        # its own stream); the binning kernels' time is the sum, the step only pays ms_bin for them
        kernels = {"k_bin": (per["ms_bin"] + per["ms_bin_overlapped"], bin_bytes),
                   "k_apply_fill": (per["ms_fill"], 8.0 * recs + 2.0 * filter_bytes * waves),
                   "k_apply_query": (per["ms_query"], 8.0 * recs + 8.0 * marks + filter_bytes * waves)}
    else:
        kernels = {"k_fill": (per["ms_fill"], 32.0 * recs + 0.375 * positions),
                   "k_query": (per["ms_query"], 32.0 * recs + 0.5 * positions)}
    dom = max(kernels, key=lambda k: kernels[k][0])
    dom_ms, dom_bytes = kernels[dom]
    gbps = lambda ms, nbytes: round(nbytes / (ms * 1e-3) / 1e9, 1) if ms > 0 else None
    achieved = gbps(dom_ms, dom_bytes) or 0.0
    # DRAM traffic of that kernel: measured once per round with `ncu --set full` on the C2 workload
    # (profiles/r*_traffic_c2.json: dram__bytes_read.sum + dram__bytes_write.sum per launch); reported here as
    # measured-traffic/algorithmic-bytes ratio of that capture x this run's algorithmic bytes.
    traffic, traffic_src = None, None
    for f in sorted((ROOT / "profiles").glob("r*_traffic_c2.json"), reverse=True):
        ratio = json.loads(f.read_text())["kernels"].get(dom, {}).get("traffic_over_algorithmic")
        if ratio:
            traffic, traffic_src = ratio * dom_bytes, f"{f.name}: ncu DRAM bytes / algorithmic bytes = {ratio} (C2 capture), scaled to this workload"
        break
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "peak_source": "measured" if peaks else "fallback", "traffic": traffic,
                "traffic_source": traffic_src,
                "algorithmic_bytes_per_step": dom_bytes, "ms_per_step": round(dom_ms, 3),
                "filter_path": "binned (L2-resident slices)" if st.bin_waves else "direct (random HBM sectors)",
                "filter_touches_per_s_G": {k: round(recs / (v[0] * 1e-3) / 1e9, 2) for k, v in kernels.items()
                                            if v[0] > 0 and k != "k_bin"},
                "all_filter_kernels": {k: {"ms": round(v[0], 3), "GBps": gbps(*v)} for k, v in kernels.items()}}
    if per.get("ms_bin_overlapped", 0) > 0:
        roofline["overlap"] = ("pipelined rounds: k_bin_list of round r+1 (3 CTAs/SM) runs beside k_apply_fill of round r (1 CTA/SM) on two "
                               "streams; k_bin's time is ms_bin + ms_bin_overlapped, the fill's time is what it takes while sharing the SMs")

    # the north-star's second roofline: uniform random 32-byte sector touches into a table of the filter's size
    # (k_probe, measured in this run before the workload was generated).  The binned path does not touch HBM at
    # random any more, so its effective rate (2 touches per owned k-mer over binning + fill + query) may exceed it.
    t_filter = sum(per[k] for k in ("ms_bin", "ms_fill", "ms_query")) * 1e-3
    eff = 2.0 * recs / t_filter / 1e9 if t_filter > 0 else None
    roofline["random_access"] = {
        "probe_Gtouch_s": probe, "unit": "G sector touches/s", "table_bits": wl["f"],
        "achieved_Gtouch_s": round(eff, 2) if eff else None,
        "frac": round(eff / probe["load_condAtomicOr"], 3) if eff and probe.get("load_condAtomicOr") else None,
        "note": "achieved = (1 fill + 1 query touch per owned k-mer) / (ms_bin + ms_fill + ms_query); "
                "frac is against the probe's load+conditional-atomicOr rate"}

    result = {
        "metric": "input Gbp/s to exact junction set", "value": round(total_bp / (ms_per_step * 1e-3) / 1e9, 4), "unit": "Gbp/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 3),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": wl["name"], "total_bp": total_bp, "k": wl["k"], "filter_bits": wl["f"], "q": wl["q"],
                   "parallelism": f"hash-range shards x{world}" if world > 1 else "single GPU",
                   "l2_hygiene": "inputs (packed genome + 2^f-bit filter) are far larger than the 126 MB L2"},
        "stages_ms": {k: round(v / args.steps, 3) for k, v in stage_ms.items()},
        "untimed_ms": round(ms_per_step - sum(v for k, v in stage_ms.items() if k != "ms_bin_overlapped") / args.steps, 3),
        "result": {**runner.last, "candidate_marks": st.candidate_marks, "candidate_kmers": st.candidate_kmers,
                   "counts_match_round1": (None if args.workload not in EXPECTED_COUNTS else
                                           (runner.last.get("junctions"), runner.last.get("records"), runner.last.get("stubs"))
                                           == EXPECTED_COUNTS[args.workload])},
        "gpu_launches": launches, "roofline": roofline,
    }
    if rank == 0:
        result["clocks"] = clocks.summary()

    # ---- end-to-end through the C ABI with host buffers (N = 1) ---------------------------------
    if not args.no_e2e and world == 1:
        runner.session.close()
        host = dg.to_host()
        codes = torch.from_numpy(host.codes.view(np.int64)).pin_memory()
        nmask = torch.from_numpy(host.n_mask.view(np.int64)).pin_memory()
        pinned = api.PackedGenome(codes.numpy().view(np.uint64), nmask.numpy().view(np.uint64), host.n_positions,
                                  host.rec_start, host.rec_len)
        out = torch.empty(runner.last["image_bytes"] + 4096, dtype=torch.uint8).pin_memory()
        out_np = out.numpy()
        times = []
        for i in range(1 + max(1, min(args.steps, 3))):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            img, st2 = api.junctions_host(pinned, k=wl["k"], filter_bits=wl["f"], q=wl["q"], out=out_np)
            torch.cuda.synchronize()
            if i:
                times.append(time.perf_counter() - t0)
        e2e_s = float(np.mean(times))
        result["e2e"] = {"value": round(total_bp / e2e_s / 1e9, 4), "unit": "Gbp/s",
                         "h2d_bytes_per_step": int(pinned.codes.nbytes + pinned.n_mask.nbytes),
                         "d2h_bytes_per_step": int(len(img)), "ms_per_step": round(e2e_s * 1e3, 3),
                         "api": "tpc_junctions_host (pinned host genome -> pinned host de_bruijn.bin image)"}
    elif not args.no_e2e:
        # N GPUs, host buffers in / host buffers out (twopaco_b200.dist.sharded_run_host): every rank uploads
        # 1/N of the packed genome from pinned host memory, chunk by chunk, NCCL all-gathers the chunks over
        # NVLink while the first pass already runs on those that have arrived, runs its shard, and copies its
        # slice of the image back to pinned host memory.  Wall clock between barriers, max over ranks.
        from twopaco_b200 import dist as tdist
        runner.session.close()
        runner.out = None
        L = api.lib()
        cw, mw = L.tpc_code_words(dg.n_positions), L.tpc_mask_words(dg.n_positions)
        shard = tdist.host_shard(cw, mw, dg.n_positions, dg.rec_start, dg.rec_len, rank, world, n_chunks=16,
                                 fetch=lambda a, lo, hi: (dg.codes if a == 0 else dg.n_mask).to_host((hi - lo) * 8, lo * 8).view(np.uint64))
        dg.codes.close(); dg.n_mask.close()                     # the e2e region starts from HOST buffers only
        out_host, dev_out, times = None, None, []
        for i in range(1 + max(1, min(args.steps, 3))):
            barrier()
            t0 = time.perf_counter()
            info, out_host, dev_out = tdist.sharded_run_host(shard, rank, world, wl["k"], wl["f"], wl["q"], wl.get("rounds", 1),
                                                             out_host, dev_out)
            barrier()
            if i:
                times.append(time.perf_counter() - t0)
        t = torch.tensor([float(np.mean(times)), float(shard.nbytes), float(info["slice_bytes"])], dtype=torch.float64, device="cuda")
        tmax = t.clone()
        torch.distributed.all_reduce(tmax, op=torch.distributed.ReduceOp.MAX)
        torch.distributed.all_reduce(t)
        e2e_s = float(tmax[0].item())
        result["e2e"] = {"value": round(total_bp / e2e_s / 1e9, 4), "unit": "Gbp/s",
                         "h2d_bytes_per_step": int(t[1].item()), "d2h_bytes_per_step": int(t[2].item()),
                         "ms_per_step": round(e2e_s * 1e3, 3),
                         "api": "twopaco_b200.dist.sharded_run_host (each rank: pinned 1/N of the packed genome -> NCCL all-gather -> "
                                "shard run -> its slice of the de_bruijn.bin image in pinned host memory)",
                         "junctions": info["junctions"], "records": info["records"]}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        result["cpu_baseline"] = cpu_baseline(sample, wl)
    if rank == 0:
        print(json.dumps(result))
    if world > 1:
        torch.distributed.destroy_process_group()


def reference_arm(args, wl, world) -> None:
    """--impl reference: the unmodified reference CPU implementation on a bounded sample of the
    same workload (rank 0 only), all host cores."""
    from oracle import oracle as O
    from twopaco_b200 import api
    if not O.have_reference():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/twopaco was not built (no /root/reference at build time)"}))
        return
    cores = os.cpu_count() or 1
    dg = api.synth_family_device(wl["seed"], wl["genomes"], wl["records"], min(wl["length"], wl["sample_bp"] * 2), wl["p"])
    recs = sample_records(dg, wl)
    bp = sum(len(r) for r in recs)
    times = []
    for i in range(args.warmup + args.steps):
        _, dt = run_reference_on(recs, wl, cores)
        if i >= args.warmup:
            times.append(dt)
    ms = float(np.mean(times)) * 1e3
    v = bp / (ms * 1e-3) / 1e9
    sample = (f"first {wl['sample_bp']} bp of record 0 of each of the {wl['genomes']} genomes ({bp} bp), "
              f"-t {cores}, wall incl. FASTA parsing")
    print(json.dumps({
        "impl": "reference", "metric": "input Gbp/s to exact junction set", "value": round(v, 6), "unit": "Gbp/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": wl["name"], "k": wl["k"], "filter_bits": min(wl["f"], 32), "q": wl["q"]},
        "cpu_baseline": {"value": round(v, 6), "unit": "Gbp/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": round(v, 6), "unit": "Gbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


if __name__ == "__main__":
    main()
