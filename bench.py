#!/usr/bin/env python
"""bench.py -- Mreads/s classified (k=31, 150 bp) on N B200s, with the lookup kernel's roofline and the
reference CPU path timed beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--reads R] [--workload NAME]

A "step" is one pass of the hot path (`classify_seqs`: reads -> canonical 31-mers -> k-mer->taxid lookup -> per-read
resolve_tree) over one batch of R synthetic 150 bp reads per GPU. Workloads (BASELINE.json `configs`):
  config2 (default)  R = 10 M reads, DB = entropy-minimised (w=50) set of the 4 test genomes, classified the way
                     `bonsai classify` does (Lex, w=k=31, canonical; 120 lookups/read, mostly misses)
  config1db          same reads against the full canonical 31-mer DB of the 4 genomes (10.5 M keys, ~70 % hits)
  config4            spaced seed (k=31, 6 gaps, comb 40), for_each_uncanon_spaced semantics, spaced DB
Under torchrun (N > 1) every rank classifies its own R reads (read-sharded, weak scaling); rank 0 builds the DB and
it is replicated with ONE NCCL broadcast at load. No per-step collective.

Prints one JSON line (see README / DESIGN.md for the keys).
"""
import argparse
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


def rank_info():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], None, [], set()
        for r in self.rows:
            if len(r) < 8:
                continue
            try:
                clk, clk_max = float(r[1]), float(r[2])
            except ValueError:
                continue
            sm.append(clk); mx = clk_max
            try:
                pw.append(float(r[3]))
            except ValueError:                       # power.draw can read [N/A]: the clocks of the row still count
                pass
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "samples": len(sm),
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons)}


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
                    label="synthetic 150bp reads vs synthetic random-stream DB (HBM-bound lookup stress, BASELINE config 5 scaled "
                          "by --stress-keys), 50% of reads hit on every k-mer")
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
        "impl": "reference", "metric": "Mreads/s classified (k=31,150bp)", "value": v, "unit": "Mreads/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": spec["label"], "reads_per_step": n},
        "cpu_baseline": {"value": v, "unit": "Mreads/s", "cores": nthreads, "kind": cpu.kind, "sample": sample},
        "e2e": {"value": v, "unit": "Mreads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "n_unclassified": int((taxon == 0).sum())}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--reads", type=int, default=10_000_000, help="reads per GPU per step")
    ap.add_argument("--workload", default="config2")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--stress-keys", type=int, default=1 << 28, help="DB keys of the stress workload (2^30 = the 16 GB table)")
    args = ap.parse_args()
    spec = workload_spec(args.workload)
    if args.impl == "reference":
        return run_reference(args, spec)

    import torch
    import torch.distributed as dist
    from bonsai_b200 import build, capi, workload as W
    rank, local_rank, world = rank_info()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = None
    if world > 1:
        from bonsai_b200 import sharding
        numa = sharding.bind_to_gpu_numa(local_rank)      # before any pinned buffer exists
        dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        build.build()
    if world > 1:
        dist.barrier()
    capi.load_library()

    g = W.load_genomes()
    c = spec["cls"]
    ctx = capi.Context(c["k"], c["w"], c["gaps"], c.get("score", capi.SCORE_LEX), c["canon"], c["api"], device=local_rank)
    # ---- database: built on rank 0, replicated with one broadcast ------------------------------------
    t_db0 = time.perf_counter()
    keys = vals = None
    stress = args.workload == "stress"
    stream = None
    if stress:
        # every rank regenerates the same seeded stream (reads are sampled from it); only rank 0 builds the table
        stream, d_keys, d_vals = W.make_stress_db(args.stress_keys, seed=77, device=dev)
        if rank == 0:
            tc, tp = W.toy_tax_arrays()
            ctx.load_pairs_device(d_keys.data_ptr(), d_vals.data_ptr(), args.stress_keys, W.STRESS_VALUES)
            ctx.load_taxonomy(tc, tp)
        del d_keys, d_vals
        torch.cuda.empty_cache()
    elif rank == 0:
        build_database(ctx, spec, g)
    bcast_ms = None
    if world > 1:
        from bonsai_b200 import sharding
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        moved = sharding.replicate_db(ctx, dist, rank, root=0, device=dev)       # ONE broadcast of the DB at load
        e1.record()
        torch.cuda.synchronize()
        bcast_ms = e0.elapsed_time(e1)
    tinfo = ctx.table_info()
    t_db = time.perf_counter() - t_db0

    # ---- reads: generated on the device, different per rank -----------------------------------------
    n = args.reads
    from_db = None
    if stress:
        d_bases, d_offs, from_db = W.make_stress_reads(stream, n, seed=1234 + rank, device=dev)
        del stream
        torch.cuda.empty_cache()
    else:
        d_bases, d_offs = W.make_reads_torch(g, n, seed=1234 + rank, device=dev)
    d_taxon = torch.zeros(n, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream()
    torch.cuda.synchronize()

    def step_device():
        ctx.classify_device(d_bases.data_ptr(), d_offs.data_ptr(), n, d_taxon.data_ptr(), stream=stream.cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local_rank)
    clocks.start()                      # nvidia-smi needs ~0.3 s to produce its first row: start it before the warm-up
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    time.sleep(0.3)
    clocks.rows.clear()                 # keep only samples taken during the timed region
    launches0 = ctx.stats()["kernel_launches"]
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record(stream)
    for a, b in ev:
        a.record(stream)
        step_device()
        b.record(stream)
    t_end.record(stream)
    barrier()
    elapsed_ms = t_start.elapsed_time(t_end)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    gpu_launches = ctx.stats()["kernel_launches"] - launches0
    # a timed region shorter than two nvidia-smi periods (small --steps): keep the same kernel running, untimed, until two
    # clock samples under this load exist (at most 1 s), and say so
    clocks_extended = False
    if clocks.proc is not None and len(clocks.rows) < 2:
        clocks_extended = True
        t_ext = time.perf_counter()
        while len(clocks.rows) < 2 and time.perf_counter() - t_ext < 1.0:
            step_device()
            torch.cuda.synchronize()
    clk = clocks.stop()
    clk["sampled_past_timed_region"] = clocks_extended
    if world > 1:
        t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    value = world * n * args.steps / (elapsed_ms * 1e-3) / 1e6

    taxon_dev = d_taxon.cpu().numpy().astype(np.uint32)
    stress_ok = None
    if stress:
        # every read cut from the stream is classified (all 120 k-mers hit), no random read is
        fd = from_db.cpu().numpy()
        stress_ok = bool((taxon_dev[fd] != 0).all() and (taxon_dev[~fd] == 0).all())

    # ---- end to end: host (pinned) buffers through the C ABI, H2D + D2H inside the timed region -------------
    h_bases = torch.empty(n * L_READ, dtype=torch.uint8, pin_memory=True)
    h_offs = torch.empty(n + 1, dtype=torch.int64, pin_memory=True)
    h_taxon = torch.empty(n, dtype=torch.int32, pin_memory=True)
    h_bases.copy_(d_bases)
    h_offs.copy_(d_offs)
    torch.cuda.synchronize()

    def step_host():
        ctx.classify_into(h_bases.data_ptr(), h_offs.data_ptr(), n, h_taxon.data_ptr())

    step_host()
    barrier()
    st0 = ctx.stats()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        step_host()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    st1 = ctx.stats()
    # bytes the library actually moved per step (it does not ship the offsets of fixed-length batches)
    h2d_step = (st1["h2d_bytes"] - st0["h2d_bytes"]) // args.e2e_steps
    d2h_step = (st1["d2h_bytes"] - st0["d2h_bytes"]) // args.e2e_steps
    if world > 1:
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * n * args.e2e_steps / e2e_s / 1e6
    same = bool(np.array_equal(h_taxon.numpy().astype(np.uint32), taxon_dev))

    # ---- roofline of the dominant kernel (bns_classify_kernel) -----------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
    # p-bar: buckets (32 B sectors) touched per lookup, measured on a sample of this rank's reads
    ns = min(n, 50_000)
    samp_b = d_bases[: ns * L_READ].cpu().numpy()
    samp_o = d_offs[: ns + 1].cpu().numpy().astype(np.uint64)
    with capi.Context(c["k"], c["w"], c["gaps"], c.get("score", capi.SCORE_LEX), c["canon"], c["api"], device=local_rank) as ectx:
        km, oo, cnt = ectx.encode(samp_b, samp_o)
        idx = np.repeat(oo[:-1], cnt) + (np.arange(int(cnt.sum()), dtype=np.uint64) - np.repeat(np.cumsum(cnt, dtype=np.uint64) - cnt, cnt))
        sample_kmers = km[idx.astype(np.int64)]
    lookups_per_read = sample_kmers.size / ns
    pbar = ctx.lookup_sectors(sample_kmers) / max(sample_kmers.size, 1)
    bytes_per_read = L_READ + lookups_per_read * pbar * 32 + 4
    achieved = n * bytes_per_read / (kernel_ms * 1e-3) / 1e9
    gather_ms = ctx.bench_gather(1 << 28)
    gather_gbs = (1 << 28) * 32 / (gather_ms * 1e-3) / 1e9
    traffic = None
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "ncu_classify_latest.json")))
        if prof.get("workload") == args.workload:
            traffic = prof["dram_bytes_per_read"] * n
    except (OSError, KeyError, ValueError):
        pass
    lean = c["w"] == c["k"] and not c["gaps"]       # what launch_classify picks for FAM_U, single-end, no hit list
    roofline = {"bound": "hbm", "kernel": "bns_classify_u_kernel" if lean else "bns_classify_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "bytes_per_read": bytes_per_read, "lookups_per_read": lookups_per_read, "sectors_per_lookup": pbar,
                "kernel_ms": kernel_ms, "table_bytes": tinfo["bytes"],
                "random_gather_gbs": gather_gbs, "frac_of_random_gather": achieved / gather_gbs,
                "note": "table of %.0f MB is L2-resident in this config; DRAM traffic << algorithmic bytes" % (tinfo["bytes"] / 1e6)
                if tinfo["bytes"] < 120e6 else "table exceeds L2"}

    # ---- CPU baseline beside it (rank 0, N = 1 only; bounded sample) -------------------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not stress:
        from oracle import pyoracle as po          # the checker, timed as the reported baseline only
        cpu = po.load_ref() or po.load_oracle()
        tc, tp = W.toy_tax_arrays()
        T = cpu.tax_from_pairs(tc, tp)
        keys, vals = ctx.table_dump()               # the CPU arm probes its own khash built from the same pairs
        db = cpu.db_from_pairs(keys, vals)
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

    if rank == 0:
        out = {
            "metric": "Mreads/s classified (k=31,150bp)", "value": value, "unit": "Mreads/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": spec["label"], "name": args.workload, "reads_per_gpu_per_step": n, "read_len": L_READ, "k": K,
                       "db_keys": tinfo["n_keys"], "db_table_mb": tinfo["bytes"] / 1e6,
                       "db_layout": "minimizer" if tinfo.get("layout") else "hash", "parallelism": "reads sharded x%d, DB replicated" % world,
                       "l2": "inputs (%.1f GB/step) exceed the 126 MB L2; no explicit flush" % (n * L_READ / 1e9),
                       "synthetic_genomes": bool(g["synthetic_genomes"]), "db_build_s": t_db, "db_broadcast_ms": bcast_ms,
                       "numa_node_rank0": numa},
            "e2e": {"value": e2e_value, "unit": "Mreads/s", "h2d_bytes_per_step": int(h2d_step),
                    "d2h_bytes_per_step": int(d2h_step), "steps": args.e2e_steps, "taxids_match_device_path": same},
            "gpu_launches": int(gpu_launches),
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "clocks": clk,
            "n_unclassified": int((taxon_dev == 0).sum()),
        }
        if stress:
            out["config"]["stress_keys"] = args.stress_keys
            out["config"]["stress_reads_classified_as_expected"] = stress_ok
        print(json.dumps(out))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
