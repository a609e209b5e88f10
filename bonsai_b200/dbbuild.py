"""Host-side database construction over the GPU encoder (minimal form of SURVEY 8(f)-1).

Reproduces what `bonsai build` does with the reference's own pieces:
  fill_set_genome   (include/bonsai/feature_min.h:68-83)   per-genome k-mer / minimizer SET through the PATH
                    overload of Encoder::for_each -- here one bns_b200_encode_batch call per genome
  update_lca_map    (include/bonsai/feature_min.h:205-228) first genome's taxid, then lca(tax, taxid, old)
  lca               (include/bonsai/util.h:634-663)
The result is a (keys, vals) list for bns_b200_load_pairs; no khash is involved.
"""
import numpy as np

from . import capi


def lca(parent, a, b):
    """util.h:634-663 over a dict child -> parent (taxid 1 -> 0)."""
    if a == b:
        return a
    if b == 0:
        return a
    if a == 0:
        return b
    nodes = []
    while a:
        nodes.append(a)
        if a not in parent:
            return 0xFFFFFFFF
        a = parent[a]
    while b:
        if b in nodes:
            return b
        if b not in parent:
            return 0xFFFFFFFF
        b = parent[b]
    return 1


def parent_map(child, parent):
    """build_parent_map, util.h:766-785"""
    m = {int(c): int(p) for c, p in zip(child, parent)}
    m[1] = 0
    return m


def genome_kmer_sets(genomes, k, w, gaps=None, score=capi.SCORE_LEX, canonicalize=True,
                     entropy_cast=capi.CAST_SATURATE, device=-1):
    """genomes: list of (bases uint8, offsets uint64[n+1]) -- the records (contigs) of each genome.
    -> list of sorted unique uint64 arrays (khash_t(all) contents)."""
    out = []
    with capi.Context(k, w, gaps, score, canonicalize, capi.API_PATH, entropy_cast, device) as ctx:
        for bases, offsets in genomes:
            kmers, oo, cnt = ctx.encode(bases, offsets)
            if cnt.size:
                idx = np.repeat(oo[:-1], cnt) + (np.arange(int(cnt.sum()), dtype=np.uint64) -
                                                 np.repeat(np.cumsum(cnt, dtype=np.uint64) - cnt, cnt))
                out.append(np.unique(kmers[idx.astype(np.int64)]))
            else:
                out.append(np.zeros(0, np.uint64))
    return out


def merge_lca(sets, taxids, tax_child, tax_parent):
    """update_lca_map over the genomes in order -> (keys sorted, vals)."""
    pm = parent_map(tax_child, tax_parent)
    assert len(sets) <= 62
    keys = np.concatenate(sets) if sets else np.zeros(0, np.uint64)
    gidx = np.concatenate([np.full(s.size, i, np.uint64) for i, s in enumerate(sets)]) if sets else np.zeros(0, np.uint64)
    order = np.argsort(keys, kind="stable")
    keys, gidx = keys[order], gidx[order]
    if keys.size == 0:
        return keys, np.zeros(0, np.uint32)
    starts = np.concatenate([[0], np.nonzero(keys[1:] != keys[:-1])[0] + 1])
    masks = np.bitwise_or.reduceat(np.uint64(1) << gidx, starts)
    ukeys = keys[starts]
    vals = np.zeros(ukeys.size, np.uint32)
    for m in np.unique(masks):
        v = None
        for i, t in enumerate(taxids):
            if (int(m) >> i) & 1:
                v = int(t) if v is None else (v if v == int(t) else lca(pm, int(t), v))
        vals[masks == m] = v
    return ukeys, vals


def build_db(genomes, taxids, tax_child, tax_parent, k, w, gaps=None, score=capi.SCORE_LEX, canonicalize=True,
             entropy_cast=capi.CAST_SATURATE, device=-1):
    sets = genome_kmer_sets(genomes, k, w, gaps, score, canonicalize, entropy_cast, device)
    return merge_lca(sets, taxids, tax_child, tax_parent)


def build_on_device(ctx, genomes, taxids, tax_child, tax_parent, k, w, gaps=None, score=capi.SCORE_LEX, canonicalize=True,
                    entropy_cast=capi.CAST_SATURATE, max_kmers=None):
    """`bonsai build` entirely on the GPU, into `ctx`'s own table: the context is switched to the database's encoder
    (record overloads), every genome is streamed through bns_b200_build_add_genome (encode -> insert-or-LCA-merge in one
    kernel), and the caller's classification encoder is restored by the caller with ctx.reconfigure(...)."""
    ctx.load_taxonomy(tax_child, tax_parent)
    ctx.reconfigure(k, w, gaps, score, canonicalize, capi.API_PATH, entropy_cast)
    bound = max_kmers or max(1024, int(sum(int(o[-1] - o[0]) for _, o in genomes)))
    while True:
        ctx.build_begin(bound, taxids)
        for (bases, offsets), t in zip(genomes, taxids):
            ctx.build_add_genome(bases, offsets, t)
        try:
            ctx.build_finish()
            return ctx.table_info()
        except capi.BnsError as e:
            if e.code != -6:
                raise
            bound *= 2
