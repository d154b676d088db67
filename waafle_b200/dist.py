"""Multi-GPU sharding: contigs are independent (waafle/waafle_orgscorer.py:943-960), so a batch
splits into contiguous contig ranges balanced by hit count; every rank scores its own range on its
own GPU with no data-path collective, and only the compacted results are gathered:

    engine.score_batch_device(shard)          H2D + kernels + compaction, results stay in HBM
    engine.pack_results()                     ONE kernel packs the 18 result arrays into one device buffer
    gather_blobs(...)                         sizes all-gather, then ONE padded NCCL gather to rank 0, issued on the
                                              engine's own stream (no host sync between compaction and the collective)
    merge_blobs(...)                          rank 0: per-rank buffers -> whole-batch result arrays

The helpers take an initialised `torch.distributed` process group (NCCL on the GPU box, gloo in
the CPU tests) -- torch is plumbing here, the engine itself never sees a torch type.
"""

import numpy as np


def shard_bounds(hit_off, world_size):
    """Contig cut points [c_0=0, ..., c_world=n] balancing sum(hits + const) per shard."""
    hit_off = np.asarray(hit_off, dtype=np.int64)
    n = len(hit_off) - 1
    cost = hit_off + 8 * np.arange(n + 1)          # 8 "hit-equivalents" of fixed work per contig
    targets = cost[-1] * np.arange(1, world_size) / float(world_size)
    cuts = np.searchsorted(cost, targets, side="left")
    bounds = np.concatenate([[0], cuts, [n]]).astype(np.int64)
    return np.maximum.accumulate(np.minimum(bounds, n))


def local_shard(batch, rank, world_size):
    b = shard_bounds(batch.hit_off, world_size)
    return batch.slice(int(b[rank]), int(b[rank + 1])), int(b[rank]), int(b[rank + 1])


def bind_rank_to_cores(local_rank, local_world):
    """Give every rank of a node its own slice of the host cores (call BEFORE allocating pinned memory, so that the
    first-touch pages of its staging buffers and its copy threads stay on those cores).  Returns the core list."""
    import os
    try:
        cores = sorted(os.sched_getaffinity(0))
        per = max(1, len(cores) // max(1, local_world))
        mine = cores[local_rank * per:(local_rank + 1) * per] or cores
        os.sched_setaffinity(0, mine)
        return mine
    except (AttributeError, OSError):
        return []


# ---------------------------------------------------------------------------------------------------------------
# packed result buffers (layout: include/waafle_b200.h, wfl_pack_results)
# ---------------------------------------------------------------------------------------------------------------

class _DeviceBytes:
    """`__cuda_array_interface__` view of a raw device pointer, so torch can wrap it without a copy."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False),
                                         "version": 2}


def device_blob(engine):
    """The engine's packed results as a torch uint8 CUDA tensor (no copy) and the torch view of the engine's stream."""
    import torch
    ptr, nbytes, stream = engine.pack_results()
    t = torch.as_tensor(_DeviceBytes(ptr, nbytes), device=torch.device("cuda", engine.device))
    return t, torch.cuda.ExternalStream(stream, device=torch.device("cuda", engine.device))


def pack_results_host(res):
    """numpy twin of wfl_pack_results (same layout), for the gloo / CPU path."""
    from .engine import _SECTIONS, results_layout
    n, nl = len(res["call"]), len(res["synteny"])
    S = res["ann_winner"].shape[1] if res["ann_winner"].ndim == 2 else 0
    members = len(res["members"])
    off, total = results_layout(n, nl, S, members)
    blob = np.zeros(total, dtype=np.uint8)
    blob[:32].view(np.int64)[:] = (n, nl, S, members)
    for (name, dt), o in zip(_SECTIONS, off):
        a = np.ascontiguousarray(res[name], dtype=dt).reshape(-1).view(np.uint8)
        blob[o:o + len(a)] = a
    return blob


def gather_blobs(blob, dist, dst=0):
    """`blob`: 1-D uint8 torch tensor (CUDA under NCCL, CPU under gloo) holding this rank's packed results.
    Returns (gathered, sizes): on `dst` a [world, max_size] uint8 tensor on the same device, elsewhere None.
    Two collectives: an all-gather of the byte counts (the only host read) and one padded gather."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    size = torch.tensor([blob.numel()], dtype=torch.int64, device=blob.device)
    sizes = torch.empty(world, dtype=torch.int64, device=blob.device)
    dist.all_gather_into_tensor(sizes, size)
    sizes = [int(s) for s in sizes.tolist()]
    width = max(sizes)
    if blob.numel() < width:
        pad = torch.empty(width, dtype=torch.uint8, device=blob.device)
        pad[:blob.numel()] = blob
    else:
        pad = blob
    out = torch.empty((world, width), dtype=torch.uint8, device=blob.device) if rank == dst else None
    dist.gather(pad, list(out.unbind(0)) if rank == dst else None, dst=dst)
    return out, sizes


def merge_results(parts, hit_bases):
    """Result dicts of consecutive contig ranges -> the result arrays of the whole batch.
    `hit_bases[r]` = index of part r's first hit in the whole batch (annotation winners are batch hit indices)."""
    out = {}
    for k in ("call", "direction", "lifts", "clade1", "clade2", "lca", "best1", "best2", "crit", "rank",
              "n_members_a", "members", "synteny", "locus_flags"):
        out[k] = np.concatenate([p[k] for p in parts])
    S = parts[0]["ann_winner"].shape[1]
    out["ann_winner"] = np.concatenate(
        [np.where(p["ann_winner"] >= 0, p["ann_winner"] + int(hb), -1).astype(np.int32).reshape(-1)
         for p, hb in zip(parts, hit_bases)]).reshape(len(out["synteny"]), S)
    moffs, base = [np.zeros(1, np.int64)], 0
    for p in parts:
        moffs.append(p["member_off"][1:] + base)
        base += int(p["member_off"][-1])
    out["member_off"] = np.concatenate(moffs)
    order = [np.nonzero(out["call"] == c)[0] for c in (2, 1, 0)]
    out["call_counts"] = np.array([len(o) for o in order], dtype=np.int64)
    out["call_index"] = np.concatenate(order).astype(np.int64)
    return out


def merge_blobs(blobs, hit_bases):
    """Per-rank packed results (uint8 numpy arrays, rank order) -> the result arrays of the whole batch."""
    from .engine import unpack_results
    return merge_results([unpack_results(b) for b in blobs], hit_bases)


def gather_results(res, hit_base, dist, device=None):
    """Host-array convenience path (gloo tests, CPU post-processing): pack, gather to rank 0, merge.
    Returns the whole-batch results on rank 0, None elsewhere."""
    import torch
    blob = torch.from_numpy(pack_results_host(res))
    if device is not None:
        blob = blob.to(device)
    world, rank = dist.get_world_size(), dist.get_rank()
    hb = torch.tensor([int(hit_base)], dtype=torch.int64, device=blob.device)
    hbs = torch.empty(world, dtype=torch.int64, device=blob.device)
    dist.all_gather_into_tensor(hbs, hb)
    gathered, sizes = gather_blobs(blob, dist)
    if rank != 0:
        return None
    g = gathered.cpu().numpy()
    return merge_blobs([g[r, :sizes[r]] for r in range(world)], hbs.tolist())
