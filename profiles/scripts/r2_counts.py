#!/usr/bin/env python
"""What the per-record counts and the run lists cost on top of the taxon-only lean kernel (BASELINE configs[1], device buffers,
CUDA events on the launching stream): taxon only / + hit & missing counts / + run lists."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import bench
from bonsai_b200 import capi, workload as W

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
spec = bench.workload_spec("config2")
c = spec["cls"]
g = W.load_genomes()
ctx = capi.Context(c["k"], c["w"], c["gaps"], capi.SCORE_LEX, c["canon"], c["api"], device=0)
bench.build_database(ctx, spec, g)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
d_bases, d_offs = W.make_reads_torch(g, n, seed=1234, device=dev)
d_tax = torch.zeros(n, dtype=torch.int32, device=dev)
d_hit = torch.zeros(n, dtype=torch.int32, device=dev)
d_miss = torch.zeros(n, dtype=torch.int32, device=dev)
cap = n * 152 + (1 << 21)
d_runs = torch.empty(cap, dtype=torch.int64, device=dev)
d_pos = torch.empty(n, dtype=torch.int64, device=dev)
d_nr = torch.empty(n, dtype=torch.int32, device=dev)
d_tot = torch.zeros(1, dtype=torch.int64, device=dev)
cs = torch.cuda.current_stream()

def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(cs)
    for _ in range(reps):
        fn()
    b.record(cs)
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

out = {}
ms = timed(lambda: ctx.classify_device(d_bases.data_ptr(), d_offs.data_ptr(), n, d_tax.data_ptr(), stream=cs.cuda_stream))
out["taxon_only"] = {"ms": ms, "mreads_s": n / ms / 1e3}
ms = timed(lambda: ctx.classify_device(d_bases.data_ptr(), d_offs.data_ptr(), n, d_tax.data_ptr(), d_hit.data_ptr(), d_miss.data_ptr(), stream=cs.cuda_stream))
out["with_counts"] = {"ms": ms, "mreads_s": n / ms / 1e3}
ms = timed(lambda: ctx.classify_device_runs(d_bases.data_ptr(), d_offs.data_ptr(), n, d_tax.data_ptr(), d_hit.data_ptr(), d_runs.data_ptr(), cap,
                                            d_pos.data_ptr(), d_nr.data_ptr(), d_tot.data_ptr(), stream=cs.cuda_stream))
out["with_runs"] = {"ms": ms, "mreads_s": n / ms / 1e3, "runs_per_read": float(d_nr.sum().item()) / n, "entries_used": int(d_tot.item())}
print(json.dumps(out))
