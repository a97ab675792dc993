#!/usr/bin/env python
"""Benchmark of the per-gene-cluster k-mer streaming hot path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A step = one pass of K1..K4 over the whole synthetic pangenome of BASELINE.json
config #2 (500 genomes x 4,000 gene clusters, 1.2 kb cut sequences, k=31,
first pass).  `value` = input bases/s with the batch already resident in HBM;
`e2e` = the same through pf_submit/pf_collect from pinned host buffers
(H2D + kernels + D2H).  N > 1: every rank runs its own 4,000 clusters of a
4,000*N-cluster pangenome (weak scaling) and a step ends with the global
pattern dedup exchange (NCCL all-to-all).

`--impl reference` times the CPU restatement of the reference's own code path
(oracle/ref_port.py, pure Python + numpy like the reference, one process per
host core) on a bounded sample of the same workload; the reference package
itself cannot travel to the GPU box.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 20261018 + 2          # seed = 20261018 + config_id
R_BYTES = 12                 # record: 8-byte key + 4-byte sample rank


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--samples", type=int, default=500)
    p.add_argument("--clusters", type=int, default=4000)
    p.add_argument("--gene-len", type=int, default=1200)
    p.add_argument("-k", type=int, default=31)
    p.add_argument("--maf", type=float, default=0.01)
    p.add_argument("--consider-missing", action="store_true")
    p.add_argument("--targets-all", action="store_true",
                   help="second pass (BASELINE config 3): every sample is a --targets strain, "
                        "positional kmers.tsv records are produced")
    p.add_argument("--sort-bits", type=int, default=0)
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--cpu-seconds", type=float, default=15.0)
    return p.parse_args()


# --------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------
class ClockSampler:
    """SM clock and clock-event reasons of one GPU, sampled every 2 ms through NVML from a
    thread of this process while the timed region runs (the timed region is tens of ms: an
    `nvidia-smi -lms` child would not have printed its first line by the time it ends).
    Falls back to `nvidia-smi -lms 20` when NVML cannot be loaded."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.p = None
        self.f = None
        self.thread = None
        self.stop_flag = False
        self.samples = []        # (sm_mhz, reasons bitmask)
        self.max_mhz = None
        self.nvml = None
        self.handle = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(device).uuid)
                if not uuid.startswith("GPU-"):
                    uuid = "GPU-" + uuid
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
                idx = device
                if vis:
                    ent = vis.split(",")[device].strip()
                    idx = int(ent) if ent.isdigit() else device
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def _loop(self):
        nv = self.nvml
        while not self.stop_flag:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)
                try:
                    why = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    why = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.samples.append((float(mhz), int(why)))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nvml is not None:
            import threading
            self.stop_flag = False
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
            return
        try:
            self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.p = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "20", "-i", str(self.device)], stdout=self.f,
                stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            nv = self.nvml
            if not self.samples:
                return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["no samples"]}
            names = (("hw_slowdown", nv.nvmlClocksThrottleReasonHwSlowdown),
                     ("hw_thermal_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                     ("sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwThermalSlowdown),
                     ("sw_power_cap", nv.nvmlClocksThrottleReasonSwPowerCap),
                     ("hw_power_brake_slowdown", nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown))
            reasons = sorted({n for _, why in self.samples for n, bit in names if why & bit})
            sm = [m for m, _ in self.samples]
            return {"sm_mhz": float(np.median(sm)), "sm_min_mhz": float(min(sm)), "sm_max_mhz": self.max_mhz,
                    "reasons": reasons, "samples": len(sm), "source": "nvml, 2 ms period, timed region only"}
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().split("\n") if r]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)),
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 20"}


# --------------------------------------------------------------------------
# CPU legs (oracle = test infrastructure; allowed here as the measured baseline)
# --------------------------------------------------------------------------
def unpack_ascii(hb, seq_slice):
    """ASCII bases + oracle seq array for a slice of a HostBatch's sequences."""
    from oracle import oracle_c
    seqs = hb.seqs[seq_slice]
    w0 = int(seqs["base_off"][0]) // 32
    w1 = int(seqs["base_off"][-1] + (seqs["len"][-1] + 63) // 64 * 64) // 32
    sh = (62 - 2 * np.arange(32)).astype(np.uint64)
    codes = ((hb.packed[w0:w1, None] >> sh[None, :]) & np.uint64(3)).astype(np.uint8).ravel()
    ascii_plane = np.frombuffer(b"ACGT", np.uint8)[codes]
    out = np.zeros(len(seqs), oracle_c.SEQ_DTYPE)
    for f in ("len", "sample", "start", "end", "offset", "strand"):
        out[f] = seqs[f]
    out["cluster"] = seqs["cluster"] - seqs["cluster"][0]
    out["off"] = seqs["base_off"] - np.uint64(w0 * 32)
    return ascii_plane, out


def cpu_baseline_c(hb, S, k, maf, consider_missing, budget_s):
    """C oracle, all host cores, on the first n clusters of the workload."""
    from oracle import oracle_c
    oracle_c.build()
    cores = os.cpu_count() or 1
    first = np.searchsorted(hb.seqs["cluster"], np.arange(len(hb.clusters) + 1))
    idx = np.arange(S)
    n = min(len(hb.clusters), max(cores, 8))
    best = None
    while True:
        sl = slice(int(first[0]), int(first[n]))
        ascii_plane, seqs = unpack_ascii(hb, sl)
        presab = ((hb.presence[:n, idx >> 5] >> (idx & 31)) & 1).astype(np.uint8)
        t0 = time.perf_counter()
        res = oracle_c.run_arrays(ascii_plane, seqs, presab, k, True, consider_missing,
                                  False, maf, n_threads=cores)
        dt = time.perf_counter() - t0
        bases = int(seqs["len"].sum())
        best = {"value": bases / dt, "unit": "bases/s", "cores": cores, "kind": "port",
                "sample": f"oracle/oracle.c (C restatement, {cores} threads) on the first {n} "
                          f"clusters of the workload: {bases} bases in {dt:.2f} s",
                "unique_kmers_per_s": res["n_unique"] / dt}
        if dt >= budget_s / 3 or n >= len(hb.clusters):
            return best
        n = min(len(hb.clusters), max(n + 1, int(n * min(8.0, budget_s / max(dt, 1e-3)) * 0.7)))


def numpy_cluster(rng, S, L, cluster_index, total_clusters):
    """One synthetic cluster as reference-style Seqinfo lists (model of SURVEY §8(d))."""
    from oracle.ref_port import CutSeq
    comp = str.maketrans("ACGT", "TGCA")
    anc = rng.integers(0, 4, L)
    founders = []
    for _ in range(8):
        f = anc.copy()
        m = rng.random(L) < 0.01
        f[m] = (f[m] + rng.integers(1, 4, int(m.sum()))) & 3
        founders.append(f)
    core = cluster_index < int(0.6 * total_clusters)
    p = 0.99 if core else rng.uniform(0.05, 0.95)
    names = [f"g{i:05d}" for i in range(S)]
    cluster, presab, absent = {}, np.zeros(S, dtype=int), []
    lut = np.frombuffer(b"ACGT", np.uint8)
    for i, s in enumerate(names):
        if rng.random() >= p:
            absent.append(s)
            continue
        presab[i] = 1
        lst = []
        for _ in range(2 if rng.random() < 0.01 else 1):
            q = founders[int(rng.integers(8))].copy()
            m = rng.random(L) < 0.001
            q[m] = (q[m] + rng.integers(1, 4, int(m.sum()))) & 3
            seq = lut[q].tobytes().decode()
            lst.append(CutSeq(seq, seq.translate(comp), s + "_1", "ctg", 1001, 1000 + L,
                              int(rng.choice([1, -1])), 100))
        cluster[s] = lst
    for s in absent:
        cluster[s] = []
    return cluster, f"group_{cluster_index}", presab


def _ref_worker(job):
    from oracle import ref_port
    item, k, maf, cm = job
    res = ref_port.kmer_stage(item, k, "", True, cm)
    pats = set()
    a, b, c = ref_port.pattern_stage((res,), True, maf, cm, pats)
    bases = sum(len(q.sequence) for v in item[0].values() for q in v)
    return bases, len(res[1]), len(b) + len(c)


def run_reference(args):
    """`--impl reference`: the reference's algorithm as restated in oracle/ref_port.py
    (pure Python + numpy, like the reference), one worker process per host core, on a
    bounded sample of the same workload per step."""
    import multiprocessing as mp
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    rng = np.random.default_rng(SEED)
    per_step = cores                      # one cluster per core per step
    ctx = mp.get_context("fork")
    steps = args.warmup + args.steps
    items = [[numpy_cluster(rng, args.samples, args.gene_len, (s * per_step + i) % args.clusters,
                            args.clusters) for i in range(per_step)] for s in range(min(steps, 2))]
    times, bases_l, uniq_l = [], [], []
    with ctx.Pool(cores) as pool:
        for s in range(steps):
            jobs = [(it, args.k, args.maf, args.consider_missing) for it in items[s % len(items)]]
            t0 = time.perf_counter()
            out = pool.map(_ref_worker, jobs, chunksize=1)
            dt = time.perf_counter() - t0
            if s >= args.warmup:
                times.append(dt)
                bases_l.append(sum(o[0] for o in out))
                uniq_l.append(sum(o[1] for o in out))
    total_t = sum(times)
    value = sum(bases_l) / total_t
    line = {
        "impl": "reference", "metric": "input_bases_per_s", "value": value, "unit": "bases/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total_t / max(1, len(times)), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": workload_config(args),
        "unique_kmers_per_s": sum(uniq_l) / total_t,
        "cpu_baseline": {"value": value, "unit": "bases/s", "cores": cores, "kind": "port",
                         "sample": f"oracle/ref_port.py (Python restatement of panfeed.py:23-235; the "
                                   f"reference package cannot travel to the GPU box), {cores} worker "
                                   f"processes, {per_step} clusters of the workload per step"},
        "e2e": {"value": value, "unit": "bases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args):
    return {"workload": f"BASELINE.json configs[1]: synthetic {args.samples} genomes x {args.clusters} "
                        f"gene clusters (~1 kb + 100 bp flanks = {args.gene_len} bp), first pass, k={args.k}",
            "samples": args.samples, "clusters_per_gpu": args.clusters, "gene_len": args.gene_len,
            "k": args.k, "maf": args.maf, "consider_missing": bool(args.consider_missing),
            "l2": "inputs larger than L2: every step reads the whole packed plane (0.6 GB at 500 x 4000) and "
                  "writes / re-reads GBs of partial rows, far beyond the 126 MB L2; nothing survives between steps"}


# --------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    import torch
    import torch.distributed as dist
    from panfeed_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1:      # the ranks share the host: split its cores between their planning threads
        os.environ.setdefault("PF_HOST_THREADS", str(max(2, (os.cpu_count() or 16) // world)))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        # rank 0 prints ONE JSON line on stdout: NCCL's version banner / warnings go to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    S, C, L, k = args.samples, args.clusters, args.gene_len, args.k
    total_clusters = C * world
    hb = capi.synth_batch(local, SEED, S, C, first_cluster=rank * C, total_clusters=total_clusters,
                          gene_len=L, pinned=True, all_targets=args.targets_all)
    n_bases = hb.n_bases
    ctx = capi.Context(k, S, canonical=True, consider_missing=args.consider_missing,
                       cluster_equal_filter=False, emit_positions=args.targets_all, maf=args.maf,
                       sort_bits=args.sort_bits, device=local)
    stream = torch.cuda.ExternalStream(ctx.stream_handle(), device=dev)
    exch = None
    if world > 1:
        from panfeed_b200 import dist as pfdist
        exch = pfdist.PatternExchange(ctx, dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    exch_ms = {}

    def step_resident():
        ctx.reset_patterns()
        ctx.execute()
        if exch is not None:
            out = exch.run()
            for ns in ("cluster", "kmer"):
                for k_, v_ in out[ns]["ms"].items():
                    exch_ms[ns + "_" + k_] = v_

    # ---- value: batch resident in HBM -------------------------------------
    ctx.upload(hb)
    sampler = ClockSampler(local)
    for _ in range(max(3, args.warmup)):
        step_resident()
    launches0 = ctx.stats()["total_launches"]
    barrier()
    sampler.start()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    stage_ms = {}
    for _ in range(args.steps):
        step_resident()
        st = ctx.stats()
        for key in ("ms_extract", "ms_hist", "ms_sort", "ms_mark", "ms_count", "ms_reduce",
                    "ms_dedup", "ms_total"):
            stage_ms[key] = stage_ms.get(key, 0.0) + st[key]
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms_total = e0.elapsed_time(e1)
    st = ctx.stats()
    launches = st["total_launches"] - launches0
    if world > 1:
        t = torch.tensor([ms_total], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    total_bases = n_bases
    if world > 1:
        t = torch.tensor([n_bases], device=dev, dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        total_bases = int(t.item())
    value = total_bases / (ms_step * 1e-3)
    # sizes of one step: read from a collect of the last execution
    ctx.collect(copy=False)
    st = ctx.stats()
    M = st["instances"]             # exactly one collected batch so far
    U = st["unique_kmers"]
    U_total = U
    if world > 1:
        t = torch.tensor([U], device=dev, dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        U_total = int(t.item())
    rows = st["rows"]
    passes = st["sort_passes"]
    for key in stage_ms:
        stage_ms[key] /= args.steps

    # ---- e2e: pinned host buffers -> pf_submit -> pf_collect -----------------
    e2e = None
    if not args.no_e2e:
        for _ in range(2):
            ctx.reset_patterns()
            ctx.submit(hb)
            ctx.collect(copy=False)
        barrier()
        t0 = time.perf_counter()
        d2h = 0
        for _ in range(args.steps):
            ctx.reset_patterns()
            ctx.submit(hb)
            r = ctx.collect(copy=False)
            if exch is not None:
                exch.run()
            d2h = r["d2h_bytes"]
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        h2d = hb.packed.nbytes + len(hb.seqs) * 64 + len(hb.clusters) * 32 + hb.presence.nbytes
        e2e = {"value": total_bases * args.steps / dt, "unit": "bases/s",
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": 1e3 * dt / args.steps, "sub_batches": ctx.stats()["sub_batches"]}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        fused = stage_ms["ms_extract"] < 0.2          # K1 ran inside the histogram / first pass
        block = st["engine"] == 2                     # block aggregation: no records at all
        pass_ms = stage_ms["ms_sort"] / max(1, passes)
        k1_bytes = n_bases / 4 + 32 * len(hb.seqs) + R_BYTES * M
        pass_bytes = 2 * R_BYTES * M                                       # SURVEY 8(d) one K2 pass
        W4 = 4 * ((S + 31) // 32)
        tr = {}
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        except Exception:
            pass
        n_launch = passes
        if block:
            kernel = ("kA_block_aggregate<canonical> (K1 extraction + all of K2 + the grouping of K3 in one "
                      "kernel: sequence chunks and k-mers are grouped in shared memory, no record reaches HBM)")
            pass_ms = stage_ms["ms_sort"]
            n_launch = 1
            alg_bytes = k1_bytes
            wp = 16 if S > 1024 else (W4 // 4 + 3) // 4 * 4      # S > 1024: 512-sample slices, 16-word rows
            partial_bytes = st["partial_rows"] * (8 + 4 * wp)
            compulsory = n_bases / 4 + 16 * len(hb.seqs) + partial_bytes
            per_rec = tr.get("kA_block_aggregate_dram_bytes_per_window")
            note = ("algorithmic = SURVEY 8(d) K1 only (N/4 + 32*seqs + R*M): the figure of the cheapest stage this "
                    "kernel replaces; it also does the declared K2 (195 B/base) and the record read of K3 without "
                    "touching HBM, so its real DRAM traffic (traffic / compulsory: packed bases in, partial "
                    "(k-mer, bitset) rows out) is ~1/5 of that and it is bound by instruction issue and "
                    "shared-memory latency, not by HBM; launch_ms includes the rescue launches and their host syncs")
        elif fused:
            kernel = "k2_extract_scatter<canonical> (K1 extraction fused into the first radix pass)"
            alg_bytes = k1_bytes + pass_bytes
            compulsory = n_bases / 4 + R_BYTES * M
            per_rec = tr.get("k2_extract_scatter_dram_bytes_per_record")
            note = ("algorithmic = SURVEY 8(d) K1 (N/4 + 32*seqs + R*M) + one declared K2 pass (2*R*M); "
                    "the fused kernel never writes then re-reads the unsorted records, so its compulsory "
                    "DRAM traffic is only N/4 + R*M (see traffic / frac_compulsory); it is instruction-bound")
        else:
            kernel = "k2_onesweep_pass<u64>"
            alg_bytes = pass_bytes
            compulsory = pass_bytes
            per_rec = tr.get("k2_onesweep_pass_dram_bytes_per_record")
            note = "one radix pass reads and writes every 12-byte record once (2*R*M)"
        achieved = alg_bytes / (pass_ms * 1e-3) / 1e9
        traffic = per_rec * M if per_rec else None     # dram read+write per launch, scaled from the ncu capture
        # whole-path algorithmic bytes per SURVEY.md 8(d), declared 8-pass model
        alg = {
            "k1_extract": k1_bytes,
            "k2_sort_declared_8_passes": 8 * M + 8 * 2 * R_BYTES * M,
            "k3_reduce": R_BYTES * M + 12 * U + rows * 4 * ((S + 31) // 32),
        }
        k3_ms = stage_ms["ms_mark"] + stage_ms["ms_count"] + stage_ms["ms_reduce"]
        k4_stage = {"ms": stage_ms["ms_dedup"],
                    "alg_GBps": (rows * W4 + st["kmer_patterns"] * W4 + 4 * rows) /
                    max(stage_ms["ms_dedup"], 1e-6) / 1e6}
        if block:
            stages = {
                "kA_block_aggregate(+rescue launches, host syncs)": {
                    "ms": stage_ms["ms_sort"], "reads_GB": n_bases / 4 / 1e9,
                    "partial_rows": st["partial_rows"], "writes_GB": partial_bytes / 1e9,
                    "alg_GBps_k1_k2_k3read_declared": (k1_bytes + alg["k2_sort_declared_8_passes"] + R_BYTES * M) /
                    stage_ms["ms_sort"] / 1e6},
                "kB_merge(kB1 insert/fold + kB3 emit, incl. host sync)": {
                    "ms": k3_ms, "alg_bytes": 2 * partial_bytes + 12 * U + rows * W4,
                    "alg_GBps": (2 * partial_bytes + 12 * U + rows * W4) / k3_ms / 1e6},
                "k4_dedup": k4_stage,
                "raw_ms": {k_: round(v_, 3) for k_, v_ in stage_ms.items()},
            }
        else:
            stages = {
                "k1_histogram(+extract)": {"ms": stage_ms["ms_extract"] + stage_ms["ms_hist"],
                                           "reads_GB": n_bases / 4 / 1e9},
                "k2_radix_passes": {"ms": stage_ms["ms_sort"], "passes": passes, "first_pass_fused_with_k1": fused,
                                    "alg_GBps": (passes * pass_bytes + (k1_bytes if fused else 0)) /
                                    stage_ms["ms_sort"] / 1e6},
                "k3_reduce(mark+local+rescue, incl. host sync)": {
                    "ms": k3_ms, "alg_GBps": alg["k3_reduce"] / k3_ms / 1e6,
                    "alg_bytes": alg["k3_reduce"]},
                "k4_dedup": k4_stage,
                "raw_ms": {k_: round(v_, 3) for k_, v_ in stage_ms.items()},
            }
        line = {
            "metric": "input_bases_per_s", "value": value, "unit": "bases/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic", "config": workload_config(args),
            "unique_kmers_per_s": U_total / (ms_step * 1e-3),
            "bases_per_step_per_gpu": n_bases, "kmer_instances_per_step_per_gpu": M,
            "unique_kmers_per_step_per_gpu": U, "rows_per_step_per_gpu": rows,
            "patterns_per_gpu": st["kmer_patterns"],
            "engine": {0: "records (partition mode)", 1: "records (full sort)", 2: "block aggregation"}[st["engine"]],
            "roofline": {"bound": "hbm", "kernel": kernel,
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "compulsory_bytes_per_launch": compulsory,
                         "frac_compulsory": compulsory / (pass_ms * 1e-3) / 1e9 / peak,
                         "launch_ms": pass_ms, "launches_per_step": n_launch, "note": note},
            "whole_step_alg_GBps_declared_model": sum(alg.values()) / (ms_step * 1e-3) / 1e9,
            "stages": stages,
            "end_to_end_alg_bytes_per_base_declared": sum(alg.values()) / n_bases,
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
        }
        if args.targets_all:
            line["positional_records_per_s"] = world * M / (ms_step * 1e-3)
            line["config"]["workload"] = line["config"]["workload"].replace("first pass", "second pass, all samples --targets")
        if exch_ms:
            line["exchange_ms_last_step_rank0"] = {k_: round(v_, 3) for k_, v_ in exch_ms.items()}
        if not args.no_cpu_baseline and world == 1:      # reported at N = 1 only
            line["cpu_baseline"] = cpu_baseline_c(hb, S, k, args.maf, args.consider_missing,
                                                  args.cpu_seconds)
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
