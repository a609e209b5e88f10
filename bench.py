#!/usr/bin/env python
"""bench.py -- Mreads/s classified (k=31, 150 bp) on N B200s, with the lookup kernel's roofline and the
reference CPU path timed beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--reads R] [--workload NAME]

A "step" is one pass of the hot path (`classify_seqs`: reads -> canonical 31-mers -> k-mer->taxid lookup -> per-read
resolve_tree) over one batch of R synthetic 150 bp reads per GPU. Workloads (BASELINE.json `configs`):
  config2 (headline) R = 10 M reads, DB = entropy-minimised (w=50) set of the 4 test genomes, classified the way
                     `bonsai classify` does (Lex, w=k=31, canonical; 120 lookups/read, mostly misses)
  config1db          same reads against the full canonical 31-mer DB of the 4 genomes (10.5 M keys, ~70 % hits)
  config4            spaced seed (k=31, 6 gaps, comb 40), for_each_uncanon_spaced semantics, spaced DB
  stress             BASELINE configs[4]: synthetic random-stream DB of --stress-keys keys (2^30 = the 34 GB table), half of
                     the reads hit on every k-mer: the HBM-bound lookup regime
The default invocation times config2 as the headline and appends `stress`, `config4` and `config1db` sub-records to the
same JSON line (each with its own device-timed value, roofline against the random-gather ceiling of ITS table, clock
samples taken inside its timed region and a 1 M-read prefix checked against the CPU reference); --no-sub skips them.
Under torchrun (N > 1) every rank classifies its own R reads (read-sharded, weak scaling); rank 0 builds the DB and
it is replicated with ONE NCCL broadcast at load, after which every rank's replica is checked against rank 0's
(`replicas_match`). No per-step collective.

Prints one JSON line (see README / DESIGN.md for the keys).
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

L_READ = 150
K = 31
METRIC = "Mreads/s classified (k=31,150bp)"


def rank_info():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


class ClockSampler:
    """SM clock, power and throttle reasons of one GPU, sampled every few ms through NVML from a thread (nvidia-smi -lms as
    the fallback): `window(t0, t1)` summarises the samples whose host time lies inside a timed region."""
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, cuda_index, period_s=0.004):
        self.rows, self.stop_flag, self.period, self.how = [], False, period_s, None
        self.h = self.nv = self.proc = None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            h = None
            try:
                uuid = str(torch.cuda.get_device_properties(cuda_index).uuid)
                h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                vis = os.environ.get("CUDA_VISIBLE_DEVICES")
                idx = int(vis.split(",")[cuda_index]) if vis and vis.split(",")[cuda_index].isdigit() else cuda_index
                h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.h, self.nv, self.how = h, pynvml, "nvml"
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.h = None
            self.idx = cuda_index
        self.th = threading.Thread(target=self._run_nvml if self.h is not None else self._run_smi, daemon=True)
        self.th.start()

    def _run_nvml(self):
        nv, h = self.nv, self.h
        reasons_fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag:
            try:
                clk = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    pw = nv.nvmlDeviceGetPowerUsage(h) / 1e3
                except Exception:
                    pw = None
                try:
                    rs = int(reasons_fn(h))
                except Exception:
                    rs = 0
                self.rows.append((time.perf_counter(), clk, self.sm_max, pw, rs))
            except Exception:
                pass
            time.sleep(self.period)

    def _run_smi(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return
        self.how = "nvidia-smi"
        for line in self.proc.stdout:
            r = [x.strip() for x in line.split(",")]
            try:
                clk, mx = float(r[0]), float(r[1])
            except (ValueError, IndexError):
                continue
            try:
                pw = float(r[2])
            except ValueError:
                pw = None
            rs = 0
            for (bit, _), v in zip(((0x8, 0), (0x40, 0), (0x20, 0), (0x4, 0)), r[3:7]):
                if v.lower().startswith("active"):
                    rs |= bit
            self.rows.append((time.perf_counter(), clk, mx, pw, rs))
            if self.stop_flag:
                break

    def window(self, t0, t1):
        rows = [r for r in list(self.rows) if t0 <= r[0] <= t1]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": [], "how": self.how or "unavailable"}
        bits = 0
        for r in rows:
            bits |= r[4]
        pw = [r[3] for r in rows if r[3] is not None]
        return {"sm_mhz": float(np.median([r[1] for r in rows])), "sm_max_mhz": rows[-1][2], "samples": len(rows),
                "power_w_max": max(pw) if pw else None, "reasons": [n for b, n in self.REASONS if bits & b],
                "how": self.how, "inside_timed_region": True}

    def stop(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()


def workload_spec(name):
    from bonsai_b200 import capi, workload as W
    if name == "config2":
        return dict(db=dict(k=K, w=50, gaps=None, score=capi.SCORE_ENTROPY, canon=True),
                    cls=dict(k=K, w=K, gaps=None, canon=True, api=capi.API_STRING), n_lookup=L_READ - K + 1,
                    label="10M synthetic 150bp reads, k=31, DB = entropy-min (w=50) 4-genome set, classify Lex w=k canonical")
    if name == "config1db":
        return dict(db=dict(k=K, w=K, gaps=None, score=capi.SCORE_LEX, canon=True),
                    cls=dict(k=K, w=K, gaps=None, canon=True, api=capi.API_STRING), n_lookup=L_READ - K + 1,
                    label="synthetic 150bp reads, k=31 w=31 Lex canonical, full 4-genome DB (10.5M keys)")
    if name == "config4":
        return dict(db=dict(k=K, w=K, gaps=W.SPACED_GAPS, score=capi.SCORE_LEX, canon=False),
                    cls=dict(k=K, w=K, gaps=W.SPACED_GAPS, canon=False, api=capi.API_PATH), n_lookup=L_READ - 40 + 1,
                    label="synthetic 150bp reads, spaced seed k=31 comb=40 (for_each_uncanon_spaced), spaced 4-genome DB")
    if name == "win50lex":       # Encoder-API-level windowed minimizers on the reads (for_each_canon_windowed, W = 20)
        return dict(db=dict(k=K, w=K, gaps=None, score=capi.SCORE_LEX, canon=True),
                    cls=dict(k=K, w=50, gaps=None, canon=True, api=capi.API_STRING), n_lookup=L_READ - 50 + 1,
                    label="synthetic 150bp reads, Lex minimizers w=50 canonical on the reads (101 lookups/read), full 4-genome DB")
    if name == "win50ent":       # rolling entropy minimizers (for_each_canon_unspaced_windowed_entropy_), tail flush
        return dict(db=dict(k=K, w=50, gaps=None, score=capi.SCORE_ENTROPY, canon=True),
                    cls=dict(k=K, w=50, gaps=None, canon=True, api=capi.API_STRING, score=capi.SCORE_ENTROPY), n_lookup=L_READ - 50 + 1,
                    label="synthetic 150bp reads, entropy minimizers w=50 canonical on the reads, entropy-min 4-genome DB")
    if name == "stress":
        return dict(db=None, cls=dict(k=K, w=K, gaps=None, canon=True, api=capi.API_STRING), n_lookup=L_READ - K + 1,
                    label="synthetic 150bp reads vs synthetic random-stream DB (HBM-bound lookup stress, BASELINE configs[4] at "
                          "--stress-keys keys), 50% of reads hit on every k-mer")
    raise SystemExit("unknown workload " + name)


def build_database(ctx, spec, g):
    """rank 0: `bonsai build` on the device (encode -> insert-or-LCA-merge kernel), straight into ctx's table"""
    from bonsai_b200 import capi, dbbuild, workload as W
    tc, tp = W.toy_tax_arrays()
    d, c = spec["db"], spec["cls"]
    genomes = [W.genome_records(g, gi) for gi in range(4)]
    dbbuild.build_on_device(ctx, genomes, W.GENOME_TAXIDS, tc, tp, d["k"], d["w"], d["gaps"], d["score"], d["canon"])
    ctx.reconfigure(c["k"], c["w"], c["gaps"], c.get("score", capi.SCORE_LEX), c["canon"], c["api"])


def run_reference(args, spec):
    """--impl reference: the reference's own CPU implementation of the path (oracle/_ref when the reference compiled
    here, else the C port) on all host threads, on a bounded sample of the same workload per step."""
    rank, _, world = rank_info()
    if rank != 0:
        return
    from bonsai_b200 import workload as W
    from oracle import pyoracle as po
    cpu = po.load_ref() or po.load_oracle()
    g = W.load_genomes()
    # the DB for the CPU arm is built by the CPU checker itself (it must not depend on the GPU library)
    o = po.load_oracle()
    tc, tp = W.toy_tax_arrays()
    To = o.tax_from_pairs(tc, tp)
    d = spec["db"]
    dbo = o.db_new()
    for gi, taxid in enumerate(W.GENOME_TAXIDS):
        o.db_add_genome(dbo, To, W.genome_records(g, gi), taxid, d["k"], d["w"], d["gaps"], d["score"], d["canon"])
    keys, vals = o.db_pairs(dbo)
    T = cpu.tax_from_pairs(tc, tp)
    db = cpu.db_from_pairs(keys, vals)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers as H
    nthreads = os.cpu_count() or 1
    c = spec["cls"]
    # size the per-step sample for ~5 s of CPU work
    probe_b, probe_o, _ = H.make_reads(20000, seed=99, genomes=g)
    for _ in range(2):                                # the first call pays for thread start-up and cold caches
        t0 = time.perf_counter()
        cpu.classify(db, T, probe_b, probe_o, c["k"], c["w"], c["gaps"], 0, c["canon"], c["api"], nthreads=nthreads)
        rate = 20000 / (time.perf_counter() - t0)
    # per-step sample: the whole --steps K --warmup W run should end within a few minutes whatever K is (~2 min of CPU work)
    per_step_s = min(3.0, max(0.2, 120.0 / max(args.steps + args.warmup, 1)))
    n = int(max(20000, min(args.reads, rate * per_step_s, 4_000_000)))
    parts, done = [], 0
    while done < n:                                   # generated in slices: the numpy generator is memory-hungry
        m = min(500_000, n - done)
        parts.append(H.make_reads(m, seed=1234 + len(parts), genomes=g)[0])
        done += m
    bases = np.concatenate(parts)
    offs = np.arange(n + 1, dtype=np.uint64) * np.uint64(L_READ)
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        taxon, _, _ = cpu.classify(db, T, bases, offs, c["k"], c["w"], c["gaps"], 0, c["canon"], c["api"], nthreads=nthreads)
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    dt = sum(times)
    v = n * len(times) / dt / 1e6
    sample = "%d reads/step of the same generator and DB, %d threads" % (n, nthreads)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "Mreads/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": spec["label"], "reads_per_step": n},
        "cpu_baseline": {"value": v, "unit": "Mreads/s", "cores": nthreads, "kind": cpu.kind, "sample": sample},
        "e2e": {"value": v, "unit": "Mreads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "n_unclassified": int((taxon == 0).sum())}))


# ----------------------------------------------------------------------------------------------------------------------
class Env:
    """per-process state shared by the workloads of one bench run"""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.rank, self.local_rank, self.world = rank_info()
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.numa = None
        if self.world > 1:
            from bonsai_b200 import sharding
            self.numa = sharding.bind_to_gpu_numa(self.local_rank)      # before any pinned buffer exists
            dist.init_process_group("nccl", device_id=self.dev)
        from bonsai_b200 import build, capi, workload as W
        if self.rank == 0:
            build.build()
        if self.world > 1:
            dist.barrier()
        capi.load_library()
        self.capi, self.W = capi, W
        self.g = W.load_genomes()
        self.clocks = ClockSampler(self.local_rank)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            pass
        self.stream_peak = float(peaks.get("hbm_gbs", 6650.0))
        self.stream_peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
        self._cpu = None

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([x], device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def cpu_ref(self):
        """the checker: oracle/_ref (unmodified reference headers) when built, else the C restatement"""
        if self._cpu is None:
            from oracle import pyoracle as po
            self._cpu = po.load_ref() or po.load_oracle()
        return self._cpu


def segment_digest(env, ctx):
    """two wrapping 64-bit sums over the raw device segments of a replica (slots, value dictionary, val_info, node_info)"""
    torch = env.torch
    out = []
    for ptr, nbytes in ctx.db_segments():
        n8 = nbytes // 8
        if not n8:
            out += [0, 0]
            continue
        t = torch.as_tensor(env.capi.DevMem(ptr, n8 * 8), device=env.dev).view(torch.int64)
        s1 = s2 = 0
        CH = 1 << 27
        for lo in range(0, n8, CH):
            w = t[lo:lo + CH]
            s1 = (s1 + int(w.sum().item())) & (2**64 - 1)
            idx = torch.arange(lo, lo + w.numel(), device=env.dev, dtype=torch.int64) * 2 + 1
            s2 = (s2 + int((w * idx).sum().item())) & (2**64 - 1)
        out += [s1, s2]
    return out


def check_replicas(env, ctx, spec, probe):
    """config 3's rule: after the broadcast every rank's replica must BE rank 0's database. Every rank computes
    digests of its device segments (+ md5 of the sorted table dump for tables up to 32 M keys) and classifies one common
    probe batch; the digests are all-gathered and compared with rank 0's."""
    torch, dist = env.torch, env.dist
    d_bases, d_offs, n = probe
    d_tax = torch.zeros(n, dtype=torch.int32, device=env.dev)
    ctx.classify_device(d_bases.data_ptr(), d_offs.data_ptr(), n, d_tax.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    words = [int(x) for x in ctx.db_export_header()] + segment_digest(env, ctx)
    words.append(int.from_bytes(hashlib.md5(d_tax.cpu().numpy().tobytes()).digest()[:8], "little"))
    tinfo = ctx.table_info()
    if tinfo["n_keys"] <= (32 << 20):
        keys, vals = ctx.table_dump()
        words.append(int.from_bytes(hashlib.md5(keys.tobytes() + vals.tobytes()).digest()[:8], "little"))
    mine = torch.tensor([w - (1 << 64) if w >= (1 << 63) else w for w in words], dtype=torch.int64, device=env.dev)
    allw = [torch.zeros_like(mine) for _ in range(env.world)]
    dist.all_gather(allw, mine)
    return bool(all(torch.equal(a, allw[0]) for a in allw)), d_tax


def oracle_prefix_check(env, spec, pairs, d_bases, d_offs, taxon_dev, n_check):
    """classify the first n_check reads of this rank's batch with the CPU reference on (keys, vals) and compare taxids"""
    cpu = env.cpu_ref()
    W = env.W
    c = spec["cls"]
    tc, tp = W.toy_tax_arrays()
    T = cpu.tax_from_pairs(tc, tp)
    db = cpu.db_from_pairs(pairs[0], pairs[1])
    sb = d_bases[: n_check * L_READ].cpu().numpy()
    so = d_offs[: n_check + 1].cpu().numpy().astype(np.uint64)
    nthreads = os.cpu_count() or 1
    t0 = time.perf_counter()
    ct, _, _ = cpu.classify(db, T, sb, so, c["k"], c["w"], c["gaps"], c.get("score", 0), c["canon"], c["api"], nthreads=nthreads)
    dt = time.perf_counter() - t0
    cpu.db_free(db)
    return {"reads": int(n_check), "taxids_match": bool(np.array_equal(ct, taxon_dev[:n_check])), "kind": cpu.kind,
            "cpu_mreads_s": n_check / dt / 1e6, "cores": nthreads, "db_keys_on_cpu": int(pairs[0].size)}


def stress_restricted_pairs(env, d_keys, d_bases, n_check):
    """The DB keys that any k-mer of the first n_check reads can hit, found with torch (sort + searchsorted -- none of this
    library's kernels): classifying those reads against this subset is classifying them against the whole DB, and it is a
    set the CPU reference can hold (the full 2^30-key khash would take 25 GB and minutes to build)."""
    torch, W = env.torch, env.W
    lut = torch.full((256,), 0, dtype=torch.uint8, device=env.dev)
    for i, ch in enumerate(b"ACGT"):
        lut[ch] = i
    found = []
    sk, _ = torch.sort(d_keys)
    CH = 1 << 17
    for lo in range(0, n_check, CH):
        m = min(CH, n_check - lo)
        codes = lut[d_bases[lo * L_READ:(lo + m) * L_READ].long()].view(m, L_READ)      # stress reads hold no invalid base
        npos = L_READ - K + 1
        c64 = codes.to(torch.int64)
        f = torch.zeros((m, npos), dtype=torch.int64, device=env.dev)
        r = torch.zeros((m, npos), dtype=torch.int64, device=env.dev)
        for j in range(K):
            f = (f << 2) | c64[:, j:j + npos]
            r = r | ((3 - c64[:, j:j + npos]) << (2 * j))
        km = torch.unique(torch.minimum(f, r).reshape(-1))
        pos = torch.searchsorted(sk, km).clamp_(max=sk.numel() - 1)
        found.append(km[sk[pos] == km])
    del sk
    keys = torch.unique(torch.cat(found))
    vals_tab = torch.from_numpy(W.STRESS_VALUES.astype(np.int64)).to(env.dev)
    vals = vals_tab[(keys % len(W.STRESS_VALUES)).long()]
    return keys.cpu().numpy().astype(np.uint64), vals.cpu().numpy().astype(np.uint32)


def h2d_ceiling(env, nbytes=1 << 30, reps=6):
    """aggregate pinned host -> device copy rate with every rank copying at once: what `e2e` can be read against"""
    torch = env.torch
    h = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h.fill_(65)
    d = torch.empty(nbytes, dtype=torch.uint8, device=env.dev)
    d.copy_(h, non_blocking=True)
    env.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    ms = env.max_over_ranks(e0.elapsed_time(e1))
    del h, d
    return env.world * nbytes * reps / (ms * 1e-3) / 1e9


def run_workload(env, name, n, steps, warmup, e2e_steps, n_check, with_cpu_baseline, with_runs=False):
    """One workload end to end: database (built on rank 0, replicated), reads, device-timed steps, e2e steps, roofline, oracle
    check. Returns the record (meaningful on rank 0; collectives are entered by every rank)."""
    torch, dist, capi, W, args = env.torch, env.dist, env.capi, env.W, env.args
    rank, world, dev = env.rank, env.world, env.dev
    spec = workload_spec(name)
    c = spec["cls"]
    ctx = capi.Context(c["k"], c["w"], c["gaps"], c.get("score", capi.SCORE_LEX), c["canon"], c["api"], device=env.local_rank)
    # ---- database: built on rank 0, replicated with one broadcast ------------------------------------
    t_db0 = time.perf_counter()
    stress = name == "stress"
    stream = d_keys = None
    if stress:
        # every rank regenerates the same seeded stream (reads are sampled from it); only rank 0 builds the table
        stream, d_keys, d_vals = W.make_stress_db(args.stress_keys, seed=77, device=dev)
        if rank == 0:
            tc, tp = W.toy_tax_arrays()
            ctx.load_pairs_device(d_keys.data_ptr(), d_vals.data_ptr(), args.stress_keys, W.STRESS_VALUES)
            ctx.load_taxonomy(tc, tp)
        else:
            d_keys = None
        del d_vals
        torch.cuda.empty_cache()
    elif rank == 0:
        build_database(ctx, spec, env.g)
    bcast_ms = None
    if world > 1:
        from bonsai_b200 import sharding
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sharding.replicate_db(ctx, dist, rank, root=0, device=dev)       # ONE broadcast of the DB at load
        e1.record()
        torch.cuda.synchronize()
        bcast_ms = e0.elapsed_time(e1)
    tinfo = ctx.table_info()
    t_db = time.perf_counter() - t_db0

    # ---- reads: generated on the device, different per rank -----------------------------------------
    from_db = None
    if stress:
        d_bases, d_offs, from_db = W.make_stress_reads(stream, n, seed=1234 + rank, device=dev)
    else:
        d_bases, d_offs = W.make_reads_torch(env.g, n, seed=1234 + rank, device=dev)
    replicas_match = None
    if world > 1:
        # one common probe batch (the same seed on every rank), classified by every replica
        npb = min(200_000, n)
        if stress:
            pb, po_, _ = W.make_stress_reads(stream, npb, seed=4242, device=dev)
        else:
            pb, po_ = W.make_reads_torch(env.g, npb, seed=4242, device=dev)
        replicas_match, _ = check_replicas(env, ctx, spec, (pb, po_, npb))
        del pb, po_
    del stream
    torch.cuda.empty_cache()
    d_taxon = torch.zeros(n, dtype=torch.int32, device=dev)
    cs = torch.cuda.current_stream()
    torch.cuda.synchronize()

    def step_device():
        ctx.classify_device(d_bases.data_ptr(), d_offs.data_ptr(), n, d_taxon.data_ptr(), stream=cs.cuda_stream)

    for _ in range(max(warmup, 3)):
        step_device()
    env.barrier()
    launches0 = ctx.stats()["kernel_launches"]
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    host_t0 = time.perf_counter()
    t_start.record(cs)
    for a, b in ev:
        a.record(cs)
        step_device()
        b.record(cs)
    t_end.record(cs)
    torch.cuda.synchronize()
    host_t1 = time.perf_counter()
    env.barrier()
    elapsed_ms = env.max_over_ranks(t_start.elapsed_time(t_end))
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    gpu_launches = ctx.stats()["kernel_launches"] - launches0
    clk = env.clocks.window(host_t0, host_t1)
    value = world * n * steps / (elapsed_ms * 1e-3) / 1e6

    taxon_dev = d_taxon.cpu().numpy().astype(np.uint32)
    stress_ok = None
    if stress:
        # every read cut from the stream is classified (all 120 k-mers hit), no random read is
        fd = from_db.cpu().numpy()
        stress_ok = bool((taxon_dev[fd] != 0).all() and (taxon_dev[~fd] == 0).all())

    # ---- end to end: host (pinned) buffers through the C ABI, H2D + D2H inside the timed region -------------
    e2e = None
    if e2e_steps > 0:
        h_bases = torch.empty(n * L_READ, dtype=torch.uint8, pin_memory=True)
        h_offs = torch.empty(n + 1, dtype=torch.int64, pin_memory=True)
        h_taxon = torch.empty(n, dtype=torch.int32, pin_memory=True)
        h_bases.copy_(d_bases)
        h_offs.copy_(d_offs)
        torch.cuda.synchronize()

        def step_host():
            ctx.classify_into(h_bases.data_ptr(), h_offs.data_ptr(), n, h_taxon.data_ptr())

        step_host()
        env.barrier()
        st0 = ctx.stats()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step_host()
        torch.cuda.synchronize()
        e2e_s = env.max_over_ranks(time.perf_counter() - t0)
        st1 = ctx.stats()
        # bytes the library actually moved per step (it does not ship the offsets of fixed-length batches)
        h2d_step = (st1["h2d_bytes"] - st0["h2d_bytes"]) // e2e_steps
        d2h_step = (st1["d2h_bytes"] - st0["d2h_bytes"]) // e2e_steps
        e2e = {"value": world * n * e2e_steps / e2e_s / 1e6, "unit": "Mreads/s", "h2d_bytes_per_step": int(h2d_step),
               "d2h_bytes_per_step": int(d2h_step), "steps": e2e_steps,
               "h2d_gbs": world * h2d_step * e2e_steps / e2e_s / 1e9,
               "taxids_match_device_path": bool(np.array_equal(h_taxon.numpy().astype(np.uint32), taxon_dev))}
        # The call packs chunks of the batch to 2 bits per base on the host's cores (inside the timed region) next to the chunks
        # that cross as ASCII; the same call with the packing threads switched off is measured beside it.
        hp = ctx.host_pack_threads()
        e2e["host_pack_threads"] = hp
        e2e["input"] = "ASCII bases in pinned host memory; the library packs part of them to 2-bit units on %d host threads inside the call" % hp \
            if hp > 0 else "ASCII bases in pinned host memory, copied as they are"
        if hp > 0:
            ctx.set_host_pack_threads(0)
            h_taxon.zero_()
            step_host()
            env.barrier()
            sa = ctx.stats()
            t0 = time.perf_counter()
            for _ in range(3):
                step_host()
            torch.cuda.synchronize()
            dt = env.max_over_ranks(time.perf_counter() - t0)
            sb_ = ctx.stats()
            e2e["ascii_only"] = {"value": world * n * 3 / dt / 1e6, "unit": "Mreads/s", "h2d_bytes_per_step": int((sb_["h2d_bytes"] - sa["h2d_bytes"]) // 3),
                                 "taxids_match_device_path": bool(np.array_equal(h_taxon.numpy().astype(np.uint32), taxon_dev))}
            ctx.set_host_pack_threads(hp)
        # the same call with Kraken run lists (bns_b200_classify_batch_runs: runs produced by the lean kernel itself)
        if with_runs:
            import ctypes as C
            nr = min(n, 4_000_000)
            h_hit = torch.empty(nr, dtype=torch.int32, pin_memory=True)
            h_pos = torch.empty(nr, dtype=torch.int64, pin_memory=True)
            h_nruns = torch.empty(nr, dtype=torch.int32, pin_memory=True)
            cap = nr * 24 + (1 << 21)
            h_runs = torch.empty(cap, dtype=torch.int64, pin_memory=True)
            total = C.c_uint64(0)

            def step_runs():
                ctx._ck(ctx.lib.bns_b200_classify_batch_runs(ctx.h, h_bases.data_ptr(), h_offs.data_ptr(), nr, 0, h_taxon.data_ptr(), h_hit.data_ptr(),
                                                             None, None, h_runs.data_ptr(), cap, h_pos.data_ptr(), h_nruns.data_ptr(), C.byref(total)))
            try:
                step_runs()
                env.barrier()
                s0 = ctx.stats()
                t0 = time.perf_counter()
                for _ in range(3):
                    step_runs()
                dt = env.max_over_ranks(time.perf_counter() - t0)
                s1 = ctx.stats()
                pos, cnt, rr = h_pos.numpy(), h_nruns.numpy(), h_runs.numpy()
                # spot check: the taxids agree with the device path and sampled records' run lengths add up to their hit counts
                ok = bool(np.array_equal(h_taxon.numpy()[:nr].astype(np.uint32), taxon_dev[:nr]))
                sample = range(0, nr, max(1, nr // 5000))
                ok = ok and all(int((rr[pos[i]:pos[i] + cnt[i]] & 0xffffffff).sum()) == int(h_hit[i]) for i in sample)
                # ... and the kernel alone on device-resident buffers (bns_b200_classify_device_runs), CUDA events on the launching stream
                cap_d = nr * (L_READ + 2) + (1 << 21)
                d_hit = torch.empty(nr, dtype=torch.int32, device=dev); d_pos = torch.empty(nr, dtype=torch.int64, device=dev)
                d_nr = torch.empty(nr, dtype=torch.int32, device=dev); d_runs = torch.empty(cap_d, dtype=torch.int64, device=dev)
                d_tot = torch.zeros(1, dtype=torch.int64, device=dev)

                def step_runs_dev():
                    ctx.classify_device_runs(d_bases.data_ptr(), d_offs.data_ptr(), nr, d_taxon.data_ptr(), d_hit.data_ptr(), d_runs.data_ptr(), cap_d,
                                             d_pos.data_ptr(), d_nr.data_ptr(), d_tot.data_ptr(), stream=cs.cuda_stream)
                for _ in range(3):
                    step_runs_dev()
                torch.cuda.synchronize()
                ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ea.record(cs)
                for _ in range(10):
                    step_runs_dev()
                eb.record(cs)
                torch.cuda.synchronize()
                dev_ms = env.max_over_ranks(ea.elapsed_time(eb)) / 10
                same_runs = bool(torch.equal(d_nr.cpu(), h_nruns[:nr]) and torch.equal(d_hit.cpu(), h_hit[:nr]))
                ok = ok and same_runs
                del d_hit, d_pos, d_nr, d_runs, d_tot
                e2e["runs"] = {"e2e_mreads_s": world * nr * 3 / dt / 1e6, "reads_per_step": nr,
                               "device_mreads_s": world * nr / (dev_ms * 1e-3) / 1e6, "device_kernel_ms": dev_ms,
                               "runs_per_read": float(cnt.sum()) / nr, "d2h_bytes_per_step": int((s1["d2h_bytes"] - s0["d2h_bytes"]) // 3),
                               "consistent_with_taxon_path": ok}
            except capi.BnsError as ex:
                e2e["runs"] = {"error": str(ex)}
            del h_hit, h_pos, h_nruns, h_runs
        del h_bases, h_offs, h_taxon

    # ---- roofline of the dominant kernel -----------------------------------------------------------------
    # p-bar: buckets (32 B sectors) touched per lookup, measured on a sample of this rank's reads
    ns = min(n, 50_000)
    samp_b = d_bases[: ns * L_READ].cpu().numpy()
    samp_o = d_offs[: ns + 1].cpu().numpy().astype(np.uint64)
    with capi.Context(c["k"], c["w"], c["gaps"], c.get("score", capi.SCORE_LEX), c["canon"], c["api"], device=env.local_rank) as ectx:
        km, oo, cnt = ectx.encode(samp_b, samp_o)
        idx = np.repeat(oo[:-1], cnt) + (np.arange(int(cnt.sum()), dtype=np.uint64) - np.repeat(np.cumsum(cnt, dtype=np.uint64) - cnt, cnt))
        sample_kmers = km[idx.astype(np.int64)]
    lookups_per_read = sample_kmers.size / ns
    sectors = ctx.lookup_sectors(sample_kmers) / max(sample_kmers.size, 1)
    # one 32-byte sector per probe. A minimizer-layout probe is a 64-byte unit (two adjacent sectors of one DRAM
    # line): it still counts as ONE algorithmic sector, so that `achieved` does not grow with the layout's own choice of probe
    # width (with it, algorithmic bytes/read ~ the DRAM bytes/read ncu measures on the big tables)
    pbar = sectors / (2.0 if tinfo.get("layout") else 1.0)
    bytes_per_read = L_READ + lookups_per_read * pbar * 32 + 4
    achieved = n * bytes_per_read / (kernel_ms * 1e-3) / 1e9
    gather_ms = ctx.bench_gather(1 << 28)
    gather_gbs = (1 << 28) * 32 / (gather_ms * 1e-3) / 1e9
    lean = not c["gaps"] or c["w"] == c["k"]         # what launch_classify picks for single-end records without a hit list
    in_l2 = tinfo["bytes"] < 120e6
    minimizer = bool(tinfo.get("layout"))
    roofline = {
        "bound": "l2/issue" if in_l2 else "hbm", "kernel": "bns_classify_u_kernel" if lean else "bns_classify_kernel",
        "achieved": achieved, "peak": gather_gbs, "unit": "GB/s", "frac": achieved / gather_gbs,
        "peak_source": "measured in this run: independent random 32-byte sector loads over the SAME table (bns_gather_kernel), SURVEY 8(d)",
        "frac_of_stream": achieved / env.stream_peak, "stream_peak": env.stream_peak, "stream_peak_source": env.stream_peak_src,
        "traffic": None,
        "bytes_per_read": bytes_per_read, "lookups_per_read": lookups_per_read, "sectors_per_lookup": pbar, "sectors_touched_per_lookup": sectors,
        "kernel_ms": kernel_ms, "table_bytes": tinfo["bytes"], "random_gather_gbs": gather_gbs,
        "frac_of_random_gather": achieved / gather_gbs,
        "note": ("table of %.0f MB is L2-resident: the kernel is bound by SM issue / L1TEX, DRAM traffic << algorithmic bytes; "
                 "traffic is not measured in-run (ncu captures under profiles/)" % (tinfo["bytes"] / 1e6)) if in_l2 else
                ("table exceeds L2" + ("; minimizer layout: consecutive k-mers of a read share 128-byte lines, so the kernel can exceed "
                                       "the independent-random-sector ceiling of the hash layout" if minimizer else ""))}

    # ---- the CPU reference beside it -------------------------------------------------------------------
    check = cpu_baseline = None
    if rank == 0:
        if stress:
            pairs = stress_restricted_pairs(env, d_keys, d_bases, min(n_check, n)) if n_check else None
        else:
            pairs = ctx.table_dump() if (n_check or with_cpu_baseline) else None
        if pairs is not None and n_check:
            check = oracle_prefix_check(env, spec, pairs, d_bases, d_offs, taxon_dev, min(n_check, n))
        if with_cpu_baseline and world == 1 and not stress:
            cpu = env.cpu_ref()
            tc, tp = W.toy_tax_arrays()
            T = cpu.tax_from_pairs(tc, tp)
            db = cpu.db_from_pairs(pairs[0], pairs[1])
            nthreads = os.cpu_count() or 1
            nb = 20000
            sb, so = d_bases[: nb * L_READ].cpu().numpy(), d_offs[: nb + 1].cpu().numpy().astype(np.uint64)
            t0 = time.perf_counter()
            cpu.classify(db, T, sb, so, c["k"], c["w"], c["gaps"], c.get("score", 0), c["canon"], c["api"], nthreads=nthreads)
            rate = nb / (time.perf_counter() - t0)
            nb = int(max(20000, min(n, rate * 10.0)))
            sb, so = d_bases[: nb * L_READ].cpu().numpy(), d_offs[: nb + 1].cpu().numpy().astype(np.uint64)
            t0 = time.perf_counter()
            ct, _, _ = cpu.classify(db, T, sb, so, c["k"], c["w"], c["gaps"], c.get("score", 0), c["canon"], c["api"], nthreads=nthreads)
            dt = time.perf_counter() - t0
            cpu_baseline = {"value": nb / dt / 1e6, "unit": "Mreads/s", "cores": nthreads, "kind": cpu.kind,
                            "sample": "first %d reads of this run's batch, same DB and taxonomy" % nb,
                            "taxids_match_gpu": bool(np.array_equal(ct, taxon_dev[:nb]))}
    del d_keys
    rec = {
        "value": value, "unit": "Mreads/s", "steps": steps, "ms_per_step": elapsed_ms / steps, "kernel_ms": kernel_ms,
        "config": {"workload": spec["label"], "name": name, "reads_per_gpu_per_step": n, "read_len": L_READ, "k": K,
                   "db_keys": tinfo["n_keys"], "db_table_mb": tinfo["bytes"] / 1e6, "db_layout": "minimizer" if minimizer else "hash",
                   "db_displaced": tinfo["n_displaced"],
                   "parallelism": "reads sharded x%d, DB replicated" % world,
                   "l2": "inputs (%.1f GB/step) exceed the 126 MB L2; no explicit flush" % (n * L_READ / 1e9),
                   "synthetic_genomes": bool(env.g["synthetic_genomes"]), "db_build_s": t_db, "db_broadcast_ms": bcast_ms,
                   "numa_node_rank0": env.numa},
        "e2e": e2e, "gpu_launches": int(gpu_launches), "roofline": roofline, "clocks": clk,
        "sectors_per_lookup": pbar, "random_gather_gbs": gather_gbs, "frac_of_random_gather": achieved / gather_gbs,
        "oracle_check": check, "taxids_match": check["taxids_match"] if check else None,
        "n_unclassified": int((taxon_dev == 0).sum()),
    }
    if cpu_baseline is not None:
        rec["cpu_baseline"] = cpu_baseline
    if world > 1:
        rec["replicas_match"] = replicas_match
    if stress:
        rec["config"]["stress_keys"] = args.stress_keys
        rec["config"]["stress_reads_classified_as_expected"] = stress_ok
    ctx.close()
    del d_bases, d_offs, d_taxon
    torch.cuda.empty_cache()
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--reads", type=int, default=10_000_000, help="reads per GPU per step")
    ap.add_argument("--workload", default="config2")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sub", action="store_true", help="skip the stress / config4 / config1db sub-records")
    ap.add_argument("--sub", default="stress,config4,config1db", help="sub-records appended to a config2 run")
    ap.add_argument("--sub-steps", type=int, default=40, help="timed steps of the config4 / config1db sub-records")
    ap.add_argument("--check-reads", type=int, default=1_000_000, help="prefix of each workload's reads checked against the CPU reference")
    ap.add_argument("--stress-keys", type=int, default=1 << 30, help="DB keys of the stress workload (2^30 = the 34 GB table of BASELINE configs[4])")
    ap.add_argument("--stress-steps", type=int, default=13, help="13 steps of 10 M reads per GPU: 1.04 B reads at 8 GPUs")
    args = ap.parse_args()
    spec = workload_spec(args.workload)
    if args.impl == "reference":
        return run_reference(args, spec)

    env = Env(args)
    main_rec = run_workload(env, args.workload, args.reads, args.steps, args.warmup, args.e2e_steps,
                            0 if not args.no_cpu_baseline and args.workload != "stress" and env.world == 1 else args.check_reads,
                            with_cpu_baseline=not args.no_cpu_baseline, with_runs=args.workload != "stress")
    ceiling = h2d_ceiling(env)
    subs = {}
    if args.workload == "config2" and not args.no_sub:
        for name in [s for s in args.sub.split(",") if s]:
            try:
                if name == "stress":
                    free, _ = env.torch.cuda.mem_get_info()
                    # table (32 B per key at one key per bucket) + keys/values/sort scratch of the generator and the checker
                    while args.stress_keys > (1 << 24) and args.stress_keys * 80 > free:
                        args.stress_keys >>= 1
                # timed regions long enough for the clock sampler; the same step count on every rank
                steps = args.stress_steps if name == "stress" else args.sub_steps
                subs[name] = run_workload(env, name, args.reads, steps, 3, min(args.e2e_steps, 3), args.check_reads, with_cpu_baseline=False)
            except Exception as e:                      # a sub-record must not take the headline down with it
                if env.world > 1:
                    raise                               # ... unless other ranks would be left waiting in a collective
                subs[name] = {"error": "%s: %s" % (type(e).__name__, e)}
                env.torch.cuda.empty_cache()
    env.clocks.stop()
    if env.rank == 0:
        out = {"metric": METRIC, "value": main_rec["value"], "unit": "Mreads/s", "n_gpus": env.world,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": main_rec["ms_per_step"],
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
               "config": main_rec["config"], "e2e": main_rec["e2e"], "gpu_launches": main_rec["gpu_launches"],
               "roofline": main_rec["roofline"], "cpu_baseline": main_rec.get("cpu_baseline"), "clocks": main_rec["clocks"],
               "n_unclassified": main_rec["n_unclassified"]}
        if out["e2e"] is not None:
            out["e2e"]["h2d_ceiling_gbs"] = ceiling
            out["e2e"]["h2d_ceiling_note"] = "aggregate pinned H2D rate with all %d ranks copying 1 GiB buffers at once" % env.world
        if main_rec.get("oracle_check"):
            out["oracle_check"] = main_rec["oracle_check"]
        if env.world > 1:
            out["replicas_match"] = main_rec.get("replicas_match")
        for name, rec in subs.items():
            out[name] = rec
        print(json.dumps(out))
    if env.world > 1:
        env.dist.destroy_process_group()


if __name__ == "__main__":
    main()
