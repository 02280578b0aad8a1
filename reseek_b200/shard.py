"""Multi-GPU sharding of a search: every pair is independent, so the streamed side (the -db set) is
block-partitioned across ranks by residue count, queries and parameters are replicated, and there is no
data-path collective.  torch.distributed is used only to gather the per-rank hit tables at the end."""
import numpy as np


def partition_by_residues(lens, world):
    """Contiguous chain ranges [lo, hi) per rank with (nearly) equal residue totals.  Returns a list of (lo, hi)."""
    lens = np.asarray(lens, np.int64)
    n = len(lens)
    cum = np.concatenate([[0], np.cumsum(lens)])
    total = int(cum[-1])
    bounds = [0]
    for r in range(1, world):
        target = total * r / world
        k = int(np.searchsorted(cum, target, side="left"))
        k = min(max(k, bounds[-1]), n)
        bounds.append(k)
    bounds.append(n)
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


def all_gather_varlen(arr, dist=None):
    """All-gather of one 1-D numpy array per rank (lengths differ); returns the list of arrays in rank order.
    gloo on CPU in the tests, NCCL on GPUs (the payload travels as a uint8 tensor over NVLink)."""
    import torch
    arr = np.ascontiguousarray(arr)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [arr]
    world = dist.get_world_size()
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    count = torch.tensor([arr.nbytes], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(count) for _ in range(world)]
    dist.all_gather(counts, count)
    counts = [int(c.item()) for c in counts]
    buf = torch.zeros(max(counts + [1]), dtype=torch.uint8, device=dev)
    if arr.nbytes:
        buf[:arr.nbytes] = torch.from_numpy(arr.view(np.uint8).reshape(-1)).to(dev)
    bufs = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(bufs, buf)
    return [bufs[r][:counts[r]].cpu().numpy().view(arr.dtype).copy() for r in range(world)]


def merge_prefilter_triples(nq, raw, t_lo, dist=None, rsb_size=0):
    """The one real exchange step of `-search Q -db DB -fast` on a sharded DB: every rank contributes the
    (target, query, score) triples of its own target block (`raw` = Context.prefilter(..., raw_only=True), targets local),
    the blocks are concatenated in rank order - which is the stream order of the unsharded DB because the shards are
    contiguous - and every rank applies RankedScoresBag to the same merged stream.  The result is therefore identical,
    ties at the cut-off included, to the single-GPU candidate list (the reference at -threads 1)."""
    from . import lib
    ts = all_gather_varlen(raw.targets.astype(np.uint32) + np.uint32(t_lo), dist)
    qs = all_gather_varlen(raw.queries.astype(np.uint32), dist)
    ss = all_gather_varlen(raw.scores.astype(np.uint16), dist)
    return lib.prefilter_bag(nq, np.concatenate(ts), np.concatenate(qs), np.concatenate(ss), rsb_size)


def search_fast_db_sharded(ctx, Q, T_local, t_lo, dist=None, index_mode=0, rsb_size=0, kl_swap=True, want_paths=True, dst=0):
    """`reseek -search Q -db DB -fast` with the DB block-partitioned across ranks: local prefilter, merged bag, local
    post-filter of the candidates that fall into this rank's block, hit gather on rank `dst` (targets re-based to the
    unsharded DB).  Returns (merged candidate list, gathered hits or None, local Results)."""
    from . import lib
    empty = T_local is None or T_local.n == 0  # more ranks than chains, or one very long chain: this rank has no block
    if empty:
        # an empty block still takes part in every collective, with empty arrays (a rank that raised here would leave the
        # others waiting in all_gather)
        class _NoTriples:
            targets = np.zeros(0, np.uint32); queries = np.zeros(0, np.uint32); scores = np.zeros(0, np.uint16)
        raw = _NoTriples()
    else:
        raw = ctx.prefilter(Q, T_local, index_mode=index_mode, rsb_size=rsb_size, kl_swap=kl_swap, raw_only=True)
    merged = merge_prefilter_triples(Q.n, raw, t_lo, dist, rsb_size)
    if empty:
        return merged, gather_hits(np.zeros(0, lib.HIT_DTYPE), t_lo, dist, dst=dst, field="b"), None
    local = merged.select(t_lo, t_lo + T_local.n)
    res = ctx.postfilter(Q, T_local, local, keep=lib.KEEP_HITS, want_paths=want_paths)
    hits = gather_hits(res.hits, t_lo, dist, dst=dst, field="b")
    return merged, hits, res


def gather_hits(local_hits, offset, dist=None, dst=0, field="a"):
    """Gather the per-rank hit records (numpy structured arrays, HIT_DTYPE) on rank `dst`.
    `offset` is added to the local indices of the sharded side (`field`: "a" for RunQuery-style searches where the
    streamed -db side is A, "b" for the -fast -db post-filter) so that they refer to the unsharded DB.  Works with any
    backend (gloo on CPU in the tests, NCCL on GPUs: the payload travels as a uint8 tensor).  The gathered records carry NO
    paths: `path_off` still points into the path pool of the rank that produced the record, so alignments have to be printed
    from the local Results (or use Context.search_*_sharded, whose gather ships the path bytes and re-bases the offsets)."""
    import torch
    hits = np.array(local_hits, copy=True)
    if len(hits):
        hits[field] += np.uint32(offset)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return hits
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    count = torch.tensor([len(hits)], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(count) for _ in range(world)]
    dist.all_gather(counts, count)
    counts = [int(c.item()) for c in counts]
    item = hits.dtype.itemsize
    mx = max(counts + [1])
    buf = torch.zeros(mx * item, dtype=torch.uint8, device=dev)
    if len(hits):
        buf[:len(hits) * item] = torch.from_numpy(hits.view(np.uint8).reshape(-1)).to(dev)
    bufs = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(bufs, buf)
    if rank != dst:
        return None
    parts = [bufs[r][:counts[r] * item].cpu().numpy().view(hits.dtype) for r in range(world)]
    return np.concatenate(parts) if parts else hits
