"""Multi-GPU plumbing: reads shard embarrassingly, the database is replicated once at load.

One process per GPU. Rank `root` holds the loaded table + taxonomy; `replicate_db` ships the 128-byte header and the
device segments (slots, value dictionary, val_info, node_info) with one broadcast each over the process group (NCCL
on GPUs; the CPU tests drive the same protocol over gloo with host buffers). No per-step collective exists.
"""
import numpy as np


def shard_range(n_records, rank, world):
    """Contiguous, balanced [lo, hi) of the records rank `rank` classifies; concatenating the ranks' outputs in rank
    order restores the input order."""
    base, rem = divmod(int(n_records), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_reads(offsets, rank, world, mates=1):
    """(record range, read range, byte range) of this rank's shard for a concatenated-bases + offsets batch."""
    n_rec = (len(offsets) - 1) // mates
    lo, hi = shard_range(n_rec, rank, world)
    return (lo, hi), (lo * mates, hi * mates), (int(offsets[lo * mates]), int(offsets[hi * mates]))


def _cuda_view(ptr, nbytes, device):
    import torch
    from . import capi
    return torch.as_tensor(capi.DevMem(ptr, nbytes), device=device)


def replicate_db(ctx, dist, rank, root=0, device=None, as_tensor=None, header_device=None):
    """Broadcast the database of `root`'s context into every other rank's context. Returns bytes moved per rank."""
    import torch
    as_tensor = as_tensor or (lambda ptr, nbytes: _cuda_view(ptr, nbytes, device))
    hdr = torch.zeros(16, dtype=torch.int64, device=header_device if header_device is not None else device)
    if rank == root:
        hdr.copy_(torch.from_numpy(ctx.db_export_header().astype(np.int64)))
    dist.broadcast(hdr, root)
    if rank != root:
        ctx.db_alloc_from_header(hdr.cpu().numpy().astype(np.uint64))
    moved = 0
    for ptr, nbytes in ctx.db_segments():
        if nbytes:
            dist.broadcast(as_tensor(ptr, nbytes), root)
            moved += nbytes
    if rank != root:
        ctx.db_commit()
    return moved


def bind_to_gpu_numa(device_index):
    """Pin this process to the CPUs of the NUMA node the GPU hangs off, BEFORE any pinned host buffer is allocated, so that
    the ring buffers the H2D copies read from are local to the GPU's PCIe root (first-touch placement). With one process
    per GPU on a two-socket box this keeps every rank's staging traffic off the inter-socket link. Returns the node or
    None when the topology cannot be read (then nothing is changed)."""
    import os
    try:
        import torch
        p = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bdf) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except (OSError, ValueError, AttributeError, RuntimeError):
        return None
