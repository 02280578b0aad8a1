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


def gather_hits(local_hits, a_offset, dist=None, dst=0):
    """Gather the per-rank hit records (numpy structured arrays, HIT_DTYPE) on rank `dst`.
    `a_offset` is added to the local A indices so that they refer to the unsharded DB.  Works with any backend
    (gloo on CPU in the tests, NCCL on GPUs: the payload travels as a uint8 tensor)."""
    import torch
    hits = np.array(local_hits, copy=True)
    if len(hits):
        hits["a"] += np.uint32(a_offset)
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
