#!/usr/bin/env python
"""bench.py -- input Gbp/s to the exact junction set (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|dev] [--impl reference]

One "step" = one full pass of the junction-finding hot path (filter fill, candidate query,
exact pass, id index, ordered emit) over one synthetic genome set.

* value      : whole-job throughput, packed genome already resident in HBM, image left in HBM.
* e2e        : same metric through the C-ABI call with HOST buffers (tpc_junctions_host):
               H2D of the packed genome and D2H of the de_bruijn.bin image inside the timed region.
* roofline   : the dominant filter kernel: algorithmic HBM bytes / CUDA-event time against the measured HBM
               peak, plus `l2_random` -- the apply kernels' sector touches/s against a probe of the same access
               pattern (random 32-byte sectors inside one L2-resident slice, record stream beside it), and
               `hbm_random` -- the direct kernels' bound (random sectors in a 2^f-bit table in HBM).
* parity     : `result.image_digest` = position-keyed digest of the de_bruijn.bin image (sum over ranks of the
               slices' digests, tpc_image_digest_device): identical at N = 1/2/4/8, and compared inside every run
               with the digest of one untimed step through the DIRECT filter kernels (a different code path) and
               with tests/golden/workload_digests.json (provenance recorded there).
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
    # configs[4]: 100 haplotypes x 24 x 129,166,667 bp (310 Gbp) streamed from host memory, k = 31, -f 40 per GPU, -r 4:
    # position-windowed (the packed genome, 116 GB, never sits in HBM); output policy: the image (~0.7 TB: SURVEY appendix F)
    # stays on the GPUs, its size and position-keyed digest are returned
    "c5": dict(name="C5: 100 synthetic human haplotypes (100x24x129166667 bp, 0.1% divergence), k=31 -f 40 -q 5 -r 4, streamed from host",
               seed=0x4831, genomes=100, records=24, length=129_166_667, p=0.001, k=31, f=40, q=5, rounds=4, windowed=1 << 17,
               group=8, sample_bp=20_000_000),
    # the same driver at a size whose resident run fits, for the windowed == resident check
    "c5mini": dict(name="C5-mini: 12 synthetic haplotypes (12x4x8000000 bp, 0.1% divergence), k=31 -f 32 -q 5 -r 4, streamed from host",
                   seed=0x4831, genomes=12, records=4, length=8_000_000, p=0.001, k=31, f=32, q=5, rounds=4, windowed=4096,
                   group=5, sample_bp=2_000_000),
    "dev": dict(name="dev: 7x2x4 Mbp, 0.1% divergence, k=25 -f 30 -q 5",
                seed=0xD0, genomes=7, records=2, length=4_000_000, p=0.001, k=25, f=30, q=5, sample_bp=500_000),
}


def golden_digest(workload: str):
    """Digest / counts of a workload's image recorded in tests/golden/workload_digests.json (with provenance)."""
    try:
        return json.loads((ROOT / "tests" / "golden" / "workload_digests.json").read_text()).get(workload)
    except OSError:
        return None


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi samples of one GPU's clocks and throttle reasons every 50 ms.  Started before the warm-up steps (nvidia-smi needs
    about a second to come up on an 8-GPU box) and asked for the samples that fell inside [mark_begin(), mark_end()]; rank 0 only --
    eight nvidia-smi loops contend for the driver."""
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int, enabled: bool = True):
        self.interval_ms = int(os.environ.get("TPC_BENCH_CLOCK_MS", "50"))   # (0 = no sampling: tuning aid)
        enabled = enabled and self.interval_ms > 0
        self.index, self.rows, self.proc, self.enabled = index, [], None, enabled
        self.t0 = self.t1 = None

    def __enter__(self):
        if not self.enabled:
            return self
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", str(self.interval_ms)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def __exit__(self, *exc):
        if self.proc:
            if self.t1 is not None:
                time.sleep(0.12)   # the sample taken at the end of the region is still being printed
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self) -> dict:
        lo = self.t0 if self.t0 is not None else float("-inf")
        hi = (self.t1 if self.t1 is not None else float("inf")) + 0.06
        rows = [r for t, r in self.rows if lo <= t <= hi and len(r) >= 6]
        sm = [int(r[0]) for r in rows if r[0].isdigit()]
        mx = [int(r[1]) for r in rows if r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows for i in range(4) if r[2 + i].lower() == "active"})
        return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "samples_outside_region": len(self.rows) - len(rows)}


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the unmodified reference on a bounded sample
# ---------------------------------------------------------------------------------------------
def sample_records(wl) -> list[bytes]:
    """Bounded sample of the workload: the first sample_bp bases of the first record of every genome, from the
    numpy restatement of the on-device generator (same bases; checked by tests/test_gpu_parity.py)."""
    from tools.benchutil import hostsynth
    return [hostsynth.record_prefix(wl["seed"], wl["records"], wl["p"], g, 0, wl["sample_bp"], wl["length"])
            for g in range(wl["genomes"])]


def reference_filter_bits(wl) -> int:
    """-f of the reference run: the workload's, unless the host cannot hold a 2^f-bit filter comfortably."""
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 0
    return wl["f"] if avail > 3 * (1 << wl["f"]) // 8 else min(wl["f"], 32)


def run_reference_on(records: list[bytes], wl: dict, threads: int, f: int):
    from oracle import oracle as O
    with tempfile.TemporaryDirectory(prefix="tpc_bench_") as d:
        paths = []
        for i, r in enumerate(records):  # one FASTA per genome, as the reference is used
            p = os.path.join(d, f"g{i}.fa")
            O.write_fasta(p, [r], names=[f"g{i}_c0"])
            paths.append(p)
        t0 = time.perf_counter()
        img, log = O.run_reference(paths, wl["k"], f, q=wl["q"], r=1, t=threads)
        dt = time.perf_counter() - t0
    return img, dt


def cpu_baseline(recs: list[bytes], wl: dict) -> dict:
    from oracle import oracle as O
    from twopaco_b200 import api
    if not O.have_reference():
        return {"value": None, "unit": "Gbp/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref/twopaco missing"}
    cores = os.cpu_count() or 1
    bp = sum(len(r) for r in recs)
    f = reference_filter_bits(wl)
    ref_img, dt = run_reference_on(recs, wl, cores, f)
    ours, _ = api.junctions_host(api.pack_records(recs), k=wl["k"], filter_bits=f, q=wl["q"])
    return {"value": bp / dt / 1e9, "unit": "Gbp/s", "cores": cores, "kind": "reference",
            "sample": f"first {wl['sample_bp']} bp of record 0 of each of the {wl['genomes']} genomes ({bp} bp), "
                      f"-k {wl['k']} -f {f} -q {wl['q']} -t {cores}, wall {dt:.2f} s incl. FASTA parsing",
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

    def digest(self):
        """Digest of the whole image = sum over ranks of the digests of the slices they hold (device-side)."""
        torch, api = self.torch, self.api
        d = api.image_digest_device(self.out.ptr, self.last["slice_bytes"], self.last["slice_offset"])
        if self.world > 1:
            t = torch.tensor([x - (1 << 64) if x >= (1 << 63) else x for x in d], dtype=torch.int64, device="cuda")
            torch.distributed.all_reduce(t)                      # int64 sums wrap mod 2^64, as the digest does
            d = tuple(int(x) & (2**64 - 1) for x in t.tolist())
        return [f"{d[0]:016x}", f"{d[1]:016x}"]


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
    ap.add_argument("--no-verify", action="store_true", help="skip the untimed direct-path step whose image digest is compared")
    ap.add_argument("--no-probe", action="store_true", help="skip the roofline probes")
    ap.add_argument("--filter-mode", default=os.environ.get("TPC_FILTER_MODE", "auto"), choices=["auto", "direct", "binned"],
                    help="filter passes of the timed steps (default: the session's own choice)")
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
    from tools import benchutil
    from twopaco_b200 import api
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if wl.get("windowed"):
        windowed_workload(args, wl, rank, world, local_rank)
        if world > 1:
            torch.distributed.destroy_process_group()
        return
    if args.filter_mode != "auto":
        os.environ["TPC_FILTER_MODE"] = args.filter_mode
    else:
        os.environ.pop("TPC_FILTER_MODE", None)

    probe, slice_probe = {}, {}
    if rank == 0 and not args.sim_world and not args.no_probe:   # roofline probes while HBM is still empty
        for mode, name in ((0, "load32B"), (2, "load_condAtomicOr")):
            probe[name] = round(benchutil.random_access_probe(wl["f"], mode, 1 << 30) / 1e9, 2)
        slice_log2 = int(os.environ.get("TPC_SLICE_LOG2", "26"))
        for mode, name in ((1, "query"), (2, "fill")):
            # best over the launch shapes the apply kernels may use; 7-fold duplicated k-mers as in the workload
            slice_probe[name] = round(max(benchutil.slice_probe(slice_log2, 6, 32 << 20, wl["genomes"], mode, U, c)
                                          for U, c in ((4, 4), (8, 2), (4, 8))) / 1e9, 2)
    dg = benchutil.synth_family_device(wl["seed"], wl["genomes"], wl["records"], wl["length"], wl["p"], keep_ascii=False)
    total_bp = dg.total_bp

    if args.sim_world:
        for i in range(args.warmup + args.steps):
            s = api.Session(k=wl["k"], filter_bits=wl["f"], q=wl["q"], shard_index=0, shard_count=args.sim_world)
            dg.attach(s)
            s.find_candidates()
            st = s.stats()
            s.close()
        print(json.dumps({"sim_world": args.sim_world, "workload": args.workload,
                          "stages_ms": {k: round(getattr(st, k), 3) for k in ("ms_bin", "ms_bin_overlapped", "ms_fill", "ms_query", "ms_insert", "ms_classify", "ms_wall_candidates")},
                          "bin_waves": st.bin_waves, "sub_rounds": st.sub_rounds, "marks": st.candidate_marks, "candidate_kmers": st.candidate_kmers}))
        return

    runner = Runner(wl, dg, rank, world)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    stage_keys = ("ms_bin", "ms_bin_overlapped", "ms_fill", "ms_query", "ms_insert", "ms_classify", "ms_index", "ms_emit")
    wall_keys = ("ms_wall_candidates", "ms_wall_index", "ms_wall_emit")
    stage_ms = {k: 0.0 for k in stage_keys + wall_keys}
    launches = 0
    with ClockSampler(local_rank, enabled=(rank == 0)) as clocks:
        for _ in range(max(args.warmup, 0)):
            runner.step()
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        clocks.mark_begin()
        ev0.record()
        for _ in range(args.steps):
            runner.step()
            st = runner.session.stats()
            for k in stage_ms:
                stage_ms[k] += getattr(st, k)
            launches += st.kernel_launches
        ev1.record()
        barrier()
        clocks.mark_end()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    st = runner.session.stats()
    digest = runner.digest()
    timed_result = dict(runner.last)

    # ---- parity of the timed code path (untimed): one more step through the DIRECT filter kernels (k_fill / k_query:
    # no binning, no sub-round pipeline, no L2-resident slices) must produce the same image, byte for byte
    verify = None
    if not args.no_verify and st.bin_waves:
        os.environ["TPC_FILTER_MODE"] = "direct"
        runner.step()
        st_d = runner.session.stats()
        d_direct = runner.digest()
        verify = {"direct_path_digest": d_direct, "binned_equals_direct": d_direct == digest and st_d.bin_waves == 0,
                  "direct_ms": {k: round(getattr(st_d, k), 3) for k in ("ms_fill", "ms_query")}}
        if args.filter_mode != "auto":
            os.environ["TPC_FILTER_MODE"] = args.filter_mode
        else:
            os.environ.pop("TPC_FILTER_MODE", None)
    gold = golden_digest(args.workload) if wl.get("rounds", 1) >= 1 else None

    # ---- roofline of the dominant filter-pass kernel ---------------------------
    # algorithmic HBM bytes, summed over the kernel's launches of one step:
    #   direct : k_fill  = 32 B (one filter sector) x owned k-mers + 0.375 B x positions (packed stream)
    #            k_query = the same + 1 bit/position of candidate mask
    #   binned : binning       = 0.375 B x positions (+ planes) per pass + 12 B x records written
    #            k_apply_fill  = 8 B x records read + the filter read once and written back once per (sub-)round
    #            k_apply_query = 8 B x records + 8 B x marks (position word + mask word) + the filter read once per (sub-)round
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    positions, recs, marks = st.positions, total_bp / world, st.candidate_marks
    filter_bytes = (1 << wl["f"]) / 8
    per = {k: v / args.steps for k, v in stage_ms.items()}
    rounds_local = wl.get("rounds", 1) * max(st.sub_rounds, 1)              # hash sub-ranges this GPU runs in sequence
    if st.bin_waves:
        waves = st.bin_waves
        passes = 1 if waves == 1 else 2
        filter_sweeps = waves * rounds_local                                 # every (sub-)round has the whole filter to itself
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
                   "k_apply_fill": (per["ms_fill"], 8.0 * recs + 2.0 * filter_bytes * filter_sweeps),
                   "k_apply_query": (per["ms_query"], 8.0 * recs + 8.0 * marks + filter_bytes * filter_sweeps)}
    else:
        kernels = {"k_fill": (per["ms_fill"], 32.0 * recs + 0.375 * positions * rounds_local),
                   "k_query": (per["ms_query"], 32.0 * recs + 0.5 * positions * rounds_local)}
    dom = max(kernels, key=lambda k: kernels[k][0])
    dom_ms, dom_bytes = kernels[dom]
    gbps = lambda ms, nbytes: round(nbytes / (ms * 1e-3) / 1e9, 1) if ms > 0 else None
    achieved = gbps(dom_ms, dom_bytes) or 0.0
    # DRAM traffic of that kernel: `ncu --set full` capture of this workload when one is committed
    # (profiles/r*_traffic_<workload>.json: dram__bytes_read.sum + dram__bytes_write.sum over the kernel's launches of a
    # step), else the measured-traffic / algorithmic-bytes ratio of the C2 capture x this run's algorithmic bytes
    traffic, traffic_src = None, None
    for pattern, scaled in ((f"r*_traffic_{args.workload}.json", False), ("r*_traffic_c2.json", True)):
        for f in sorted((ROOT / "profiles").glob(pattern), reverse=True):
            entry = json.loads(f.read_text())["kernels"].get(dom, {})
            if not scaled and entry.get("dram_bytes_per_step") and entry.get("n_gpus", 1) == world:
                traffic, traffic_src = entry["dram_bytes_per_step"], f"{f.name}: ncu dram__bytes_read.sum + dram__bytes_write.sum over the launches of one step"
            elif scaled and entry.get("traffic_over_algorithmic"):
                ratio = entry["traffic_over_algorithmic"]
                traffic, traffic_src = ratio * dom_bytes, f"{f.name}: ncu DRAM bytes / algorithmic bytes = {ratio} (C2 capture), scaled to this workload"
            break
        if traffic:
            break
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "peak_source": "measured" if peaks else "fallback", "traffic": traffic,
                "traffic_source": traffic_src,
                "algorithmic_bytes_per_step": dom_bytes, "ms_per_step": round(dom_ms, 3),
                "filter_path": "binned (L2-resident slices)" if st.bin_waves else "direct (random HBM sectors)",
                "all_filter_kernels": {k: {"ms": round(v[0], 3), "GBps": gbps(*v), "frac_of_hbm_peak": round((gbps(*v) or 0) / peak, 4)}
                                       for k, v in kernels.items()}}
    if per.get("ms_bin_overlapped", 0) > 0:
        roofline["overlap"] = ("pipelined rounds: the binning of round r+1 runs beside the fill of round r on a second stream; "
                               "k_bin's time is ms_bin + ms_bin_overlapped, the fill's time is what it takes while sharing the SMs")
    touches = {k: round(recs / (v[0] * 1e-3) / 1e9, 2) for k, v in kernels.items() if v[0] > 0 and k != "k_bin"}
    if st.bin_waves:
        # the binned apply kernels touch one random 32-byte sector per record INSIDE an L2-resident slice: their bound is
        # the probe of exactly that pattern (tpcb_slice_probe), not a random-HBM probe
        roofline["l2_random"] = {
            "unit": "G sector touches/s", "slice_bytes": 1 << int(os.environ.get("TPC_SLICE_LOG2", "26")),
            "probe": slice_probe, "achieved": {k: touches.get(k) for k in ("k_apply_fill", "k_apply_query")},
            "frac": {k: (round(touches[k] / slice_probe[n], 3) if touches.get(k) and slice_probe.get(n) else None)
                     for k, n in (("k_apply_fill", "fill"), ("k_apply_query", "query"))},
            "note": "probe = random 256-bit sector loads (+ the fill's conditional atomicOr) inside one slice with the 8-byte record "
                    "stream read beside it, best launch shape; pipelined rounds: the fill shares the SMs with the binning kernel"}
        roofline["hbm_random"] = {"probe_Gtouch_s": probe, "note": "bound of the DIRECT kernels only (bench.py --filter-mode direct)"}
    else:
        roofline["hbm_random"] = {
            "unit": "G sector touches/s", "table_bits": wl["f"], "probe_Gtouch_s": probe, "achieved": touches,
            "frac": {"k_fill": round(touches["k_fill"] / probe["load_condAtomicOr"], 3) if probe.get("load_condAtomicOr") and touches.get("k_fill") else None,
                     "k_query": round(touches["k_query"] / probe["load32B"], 3) if probe.get("load32B") and touches.get("k_query") else None},
            "note": "one random 32-byte sector per owned k-mer and pass; probe = k_probe (uniform random sectors of a 2^f-bit table)"}

    timed_stage_sum = sum(v for k, v in stage_ms.items() if k in stage_keys and k != "ms_bin_overlapped") / args.steps
    result = {
        "metric": "input Gbp/s to exact junction set", "value": round(total_bp / (ms_per_step * 1e-3) / 1e9, 4), "unit": "Gbp/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 3),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": wl["name"], "total_bp": total_bp, "k": wl["k"], "filter_bits": wl["f"], "q": wl["q"],
                   "parallelism": f"hash-range shards x{world}" if world > 1 else "single GPU",
                   "l2_hygiene": "inputs (packed genome + 2^f-bit filter) are far larger than the 126 MB L2"},
        "stages_ms": {k: round(v / args.steps, 3) for k, v in stage_ms.items()},
        "untimed_ms": round(ms_per_step - timed_stage_sum, 3),
        "sub_rounds": st.sub_rounds,
        "result": {**timed_result, "candidate_marks": st.candidate_marks, "candidate_kmers": st.candidate_kmers,
                   "image_digest": digest, "verify": verify,
                   "golden": (None if not gold else
                              {"digest_matches": gold.get("digest") == digest,
                               "counts_match": [gold.get("junctions"), gold.get("records"), gold.get("stubs")]
                                               == [timed_result.get("junctions"), timed_result.get("records"), timed_result.get("stubs")],
                               "provenance": gold.get("provenance")})},
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
        out = torch.empty(timed_result["image_bytes"] + 4096, dtype=torch.uint8).pin_memory()
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
        d_host = api.image_digest_host(img)
        result["e2e"] = {"value": round(total_bp / e2e_s / 1e9, 4), "unit": "Gbp/s",
                         # bytes the library actually copied (tpc_stats.h2d_bytes): the packed codes + the non-uniform 64 KiB
                         # blocks of the n-mask; its all-zero / all-one blocks are set on the device
                         "h2d_bytes_per_step": int(st2.h2d_bytes) or int(pinned.codes.nbytes + pinned.n_mask.nbytes),
                         "host_input_bytes": int(pinned.codes.nbytes + pinned.n_mask.nbytes),
                         "d2h_bytes_per_step": int(len(img)), "ms_per_step": round(e2e_s * 1e3, 3),
                         "api": "tpc_junctions_host (pinned host genome -> pinned host de_bruijn.bin image)",
                         "image_digest_equals_device_run": [f"{d_host[0]:016x}", f"{d_host[1]:016x}"] == digest}
    elif not args.no_e2e:
        # N GPUs, host buffers in / host buffers out through the C ABI: rank 0 calls tpc_multi_junctions_host (one process, one
        # host thread per GPU, every GPU uploads 1/N of each chunk of the packed genome from rank 0's pinned host memory over its
        # own PCIe link, NCCL all-gathers the chunks over NVLink while the first pass already runs, every GPU copies its slice
        # of the image back).  The other ranks of the launch free their GPUs and wait on the CPU.  Wall clock on rank 0.
        cpu_group = torch.distributed.new_group(backend="gloo")
        runner.session.close()
        runner.out = None
        e2e = None
        if rank == 0:
            host = dg.to_host()
            codes = torch.empty(len(host.codes), dtype=torch.int64, pin_memory=True)
            nmask = torch.empty(len(host.n_mask), dtype=torch.int64, pin_memory=True)
            codes.numpy().view(np.uint64)[:] = host.codes
            nmask.numpy().view(np.uint64)[:] = host.n_mask
            pinned = api.PackedGenome(codes.numpy().view(np.uint64), nmask.numpy().view(np.uint64), host.n_positions, host.rec_start, host.rec_len)
            del host
        dg.codes.close(); dg.n_mask.close()                     # the e2e region starts from HOST buffers only
        torch.cuda.empty_cache()
        api.release_cached_memory()                             # (this process's pool: rank 0's threads need the GPU's memory)
        torch.distributed.barrier(group=cpu_group)
        if rank == 0:
            out = torch.empty((timed_result["records"] + len(pinned.rec_len)) * 12 + 4096, dtype=torch.uint8, pin_memory=True).numpy()
            mg = api.MultiGpu(world)
            times = []
            for i in range(1 + max(1, min(args.steps, 3))):
                t0 = time.perf_counter()
                img, st2 = mg.junctions_host(pinned, k=wl["k"], filter_bits=wl["f"], q=wl["q"], rounds=wl.get("rounds", 1), out=out)
                if i:
                    times.append(time.perf_counter() - t0)
            mg.close()
            e2e_s = float(np.mean(times))
            d_host = api.image_digest_host(img)
            e2e = {"value": round(total_bp / e2e_s / 1e9, 4), "unit": "Gbp/s",
                   "h2d_bytes_per_step": int(pinned.codes.nbytes + pinned.n_mask.nbytes), "d2h_bytes_per_step": int(len(img)),
                   "ms_per_step": round(e2e_s * 1e3, 3),
                   "api": "tpc_multi_junctions_host (C ABI, one process: pinned host genome -> 1/N upload per GPU + NCCL all-gather -> "
                          "hash-range shards -> de_bruijn.bin image in pinned host memory)",
                   "junctions": st2.junctions, "records": st2.occurrences,
                   "image_digest_equals_device_run": [f"{d_host[0]:016x}", f"{d_host[1]:016x}"] == digest}
        torch.distributed.barrier(group=cpu_group)
        if rank == 0:
            result["e2e"] = e2e

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        result["cpu_baseline"] = cpu_baseline(sample_records(wl), wl)
    if rank == 0:
        print(json.dumps(result))
    if world > 1:
        torch.distributed.destroy_process_group()


def windowed_workload(args, wl, rank, world, local_rank) -> None:
    """Workloads larger than HBM (C5): the packed genome sits in rank 0's pinned host memory and streams through the N
    GPUs window by window (tpc_multi_junctions_digest: one process, one host thread per GPU, NCCL all-gather of every
    window; N = 1: tpc_junctions_host).  The other ranks of the torchrun launch hold no GPU memory and wait.  A step = the
    whole path, host buffers in, image digest out (the image itself stays on the GPUs: `output_policy`)."""
    import torch
    from tools import benchutil
    from twopaco_b200 import api

    # the other ranks wait on the CPU (a gloo group): an NCCL barrier would spin on their GPUs, which rank 0's threads use
    cpu_group = torch.distributed.new_group(backend="gloo") if world > 1 else None

    def barrier():
        if world > 1:
            torch.distributed.barrier(group=cpu_group)
        torch.cuda.synchronize()

    result = None
    if rank == 0:
        t0 = time.perf_counter()
        host = benchutil.synth_family_host(wl["seed"], wl["genomes"], wl["records"], wl["length"], wl["p"], group=wl.get("group", 8))
        gen_s = time.perf_counter() - t0
        total_bp = host.total_bp
        os.environ["TPC_WINDOW_TILES"] = str(wl["windowed"])
        mg = api.MultiGpu(world) if world > 1 else None
        out = None

        def step():
            nonlocal out
            if mg:
                return mg.junctions_digest(host, k=wl["k"], filter_bits=wl["f"], q=wl["q"], rounds=wl["rounds"])
            img, st = api.junctions_host(host, k=wl["k"], filter_bits=wl["f"], q=wl["q"], rounds=wl["rounds"], out=out)
            out = img.base if img.base is not None else img
            return img, len(img), st

        for _ in range(max(args.warmup, 0)):
            step()
        times = []
        with ClockSampler(local_rank) as clocks:
            clocks.mark_begin()
            for _ in range(args.steps):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                digest, nbytes, st = step()
                times.append(time.perf_counter() - t0)
            clocks.mark_end()
        sec = float(np.mean(times))
        if not mg:
            digest = api.image_digest_host(digest)   # (one GPU: the image came back; its digest is computed outside the timed region)
        verify = None
        if not args.no_verify and wl["f"] <= 34:   # (small enough for the resident path: same digest?)
            os.environ["TPC_WINDOW_TILES"] = "0"
            d2, n2, _ = step()
            if not mg:
                d2 = api.image_digest_host(d2)
            verify = {"windowed_equals_resident": d2 == digest and n2 == nbytes}
            os.environ["TPC_WINDOW_TILES"] = str(wl["windowed"])
        if mg:
            mg.close()
        v = total_bp / sec / 1e9
        h2d = int(host.codes.nbytes + host.n_mask.nbytes)
        passes = 2 * wl["rounds"] + (2 if world > 1 else 1)
        result = {
            "metric": "input Gbp/s to exact junction set", "value": round(v, 4), "unit": "Gbp/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(sec * 1e3, 3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": {"workload": wl["name"], "total_bp": total_bp, "k": wl["k"], "filter_bits": wl["f"], "q": wl["q"], "rounds": wl["rounds"],
                       "parallelism": f"hash-range shards x{world}, position-windowed ({wl['windowed']} tiles = {wl['windowed'] * 8192} positions per window)",
                       "output_policy": "the de_bruijn.bin image is produced window by window on the GPUs and reduced to its size and "
                                        "position-keyed digest there (SURVEY appendix F: ~0.7 TB of records at C5); nothing but 16 bytes is read back",
                       "l2_hygiene": "every pass streams the whole packed genome (far larger than L2) from host memory"},
            "stages_ms": {k: round(getattr(st, k), 3) for k in ("ms_fill", "ms_query", "ms_classify", "ms_index", "ms_wall_candidates", "ms_wall_emit")},
            "result": {"junctions": st.junctions, "records": st.occurrences, "image_bytes": nbytes,
                       "image_digest": [f"{digest[0]:016x}", f"{digest[1]:016x}"], "verify": verify},
            "gpu_launches": st.kernel_launches,
            "e2e": {"value": round(v, 4), "unit": "Gbp/s", "h2d_bytes_per_step": h2d * passes,
                    "d2h_bytes_per_step": 16 if mg else int(nbytes), "ms_per_step": round(sec * 1e3, 3),
                    "api": "tpc_multi_junctions_digest (pinned host genome -> windows over PCIe + NCCL all-gather -> digest)" if mg else
                           "tpc_junctions_host (pinned host genome -> windows -> host image)",
                    "note": f"the packed genome crosses PCIe once per pass: {passes} passes (fill + query per round, emit count + write)"},
            "roofline": {"bound": "hbm", "kernel": "k_fill + k_query (direct kernels: a 2^40-bit filter has 2048 L2-sized slices, the binned path does not apply)",
                         "achieved": None, "peak": None, "unit": "GB/s", "frac": None, "traffic": None,
                         "note": "windowed runs are bound by the host link and the random-sector rate of the direct kernels; see DESIGN.md"},
            "input_generation_s": round(gen_s, 1),
            "clocks": clocks.summary(),
        }
    barrier()
    if rank == 0:
        print(json.dumps(result))


def reference_arm(args, wl, world) -> None:
    """--impl reference: the unmodified reference CPU implementation (oracle/_ref/twopaco, all host cores) on a bounded
    sample of the same workload, rank 0 only.  The sample comes from the numpy restatement of the generator: this arm
    needs no GPU and loads none of this repository's libraries."""
    from oracle import oracle as O
    if not O.have_reference():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/twopaco was not built (no /root/reference at build time)"}))
        return
    cores = os.cpu_count() or 1
    recs = sample_records(wl)
    bp = sum(len(r) for r in recs)
    f = reference_filter_bits(wl)
    times = []
    for i in range(args.warmup + args.steps):
        _, dt = run_reference_on(recs, wl, cores, f)
        if i >= args.warmup:
            times.append(dt)
    ms = float(np.mean(times)) * 1e3
    v = bp / (ms * 1e-3) / 1e9
    total_bp = wl["genomes"] * wl["records"] * wl["length"]
    sample = (f"first {wl['sample_bp']} bp of record 0 of each of the {wl['genomes']} genomes ({bp} bp of the workload's ~{total_bp} bp), "
              f"-k {wl['k']} -f {f} -q {wl['q']} -t {cores}, wall incl. FASTA parsing; Gbp/s of the sample, i.e. a linear extrapolation to the workload")
    print(json.dumps({
        "impl": "reference", "metric": "input Gbp/s to exact junction set", "value": round(v, 6), "unit": "Gbp/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": wl["name"], "k": wl["k"], "filter_bits": wl["f"], "q": wl["q"], "reference_filter_bits": f},
        "cpu_baseline": {"value": round(v, 6), "unit": "Gbp/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": round(v, 6), "unit": "Gbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


if __name__ == "__main__":
    main()
