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
