import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from bonsai_b200 import capi, workload as W
import bench
class A: pass
spec = bench.workload_spec("config2")
g = W.load_genomes()
c = spec["cls"]
dev = torch.device("cuda", 0)
ctx = capi.Context(c["k"], c["w"], c["gaps"], capi.SCORE_LEX, c["canon"], c["api"], device=0, host_pack_threads=0)
bench.build_database(ctx, spec, g)
n = 10_000_000
d_bases, d_offs = W.make_reads_torch(g, n, seed=1234, device=dev)
h_bases = torch.empty(n * 150, dtype=torch.uint8, pin_memory=True); h_offs = torch.empty(n + 1, dtype=torch.int64, pin_memory=True); h_taxon = torch.empty(n, dtype=torch.int32, pin_memory=True)
h_bases.copy_(d_bases); h_offs.copy_(d_offs); torch.cuda.synchronize()
d_taxon = torch.zeros(n, dtype=torch.int32, device=dev)
cs = torch.cuda.current_stream()
for nn in (10_000_000, 1_022_976, 2_045_952, 113_664 * 20):
    for _ in range(3): ctx.classify_device(d_bases.data_ptr(), d_offs.data_ptr(), nn, d_taxon.data_ptr(), stream=cs.cuda_stream)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(cs)
    for _ in range(10): ctx.classify_device(d_bases.data_ptr(), d_offs.data_ptr(), nn, d_taxon.data_ptr(), stream=cs.cuda_stream)
    b.record(cs); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print("device-resident %9d reads: %.3f ms  %.1f Mreads/s" % (nn, ms, nn / ms / 1e3))
ctx.classify_into(h_bases.data_ptr(), h_offs.data_ptr(), n, h_taxon.data_ptr())
s0 = ctx.stats()
for _ in range(3): ctx.classify_into(h_bases.data_ptr(), h_offs.data_ptr(), n, h_taxon.data_ptr())
s1 = ctx.stats()
print("host path (ASCII only), kernels of the chunks per 10 M reads: %.3f ms, launches %d" % ((s1["kernel_ms_total"] - s0["kernel_ms_total"]) / 3, (s1["kernel_launches"] - s0["kernel_launches"]) / 3))
