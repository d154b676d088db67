"""Multi-GPU sharding: contigs are independent (waafle/waafle_orgscorer.py:943-960), so a batch
splits into contiguous contig ranges balanced by hit count; every rank scores its own range on its
own GPU with no data-path collective, and only the compacted results are gathered.

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


RECORD_FIELDS = ["call", "direction", "lifts", "clade1", "clade2", "lca", "best1", "best2", "crit", "rank",
                 "n_members_a"]
LOCUS_FIELDS = ["synteny", "locus_flags"]


def gather_results(res, batch_shard, hit_base, dist, device=None):
    """All-gather per-shard result arrays into whole-batch arrays (every rank gets the result).

    Variable-length pieces travel as (sizes all-gather) + padded all-gather, which is the
    NCCL-friendly form of a gatherv; volumes are ~50-150 B per contig.
    """
    import torch
    world = dist.get_world_size()
    dev = device or "cpu"

    def allgather_var(a):
        a = np.ascontiguousarray(a)
        t = torch.from_numpy(a.view(np.uint8).reshape(-1)).to(dev)
        size = torch.tensor([t.numel()], dtype=torch.int64, device=dev)
        sizes = [torch.zeros_like(size) for _ in range(world)]
        dist.all_gather(sizes, size)
        sizes = [int(s.item()) for s in sizes]
        pad = torch.zeros(max(max(sizes), 1), dtype=torch.uint8, device=dev)
        pad[:t.numel()] = t
        outs = [torch.zeros_like(pad) for _ in range(world)]
        dist.all_gather(outs, pad)
        parts = [o[:s].cpu().numpy().view(a.dtype) for o, s in zip(outs, sizes)]
        return parts

    out = {}
    for k in RECORD_FIELDS + LOCUS_FIELDS:
        out[k] = np.concatenate(allgather_var(res[k]))
    S = res["ann_winner"].shape[1]
    aw = np.where(res["ann_winner"] >= 0, res["ann_winner"] + hit_base, -1).astype(np.int32)
    out["ann_winner"] = np.concatenate(allgather_var(aw)).reshape(-1, S) if S else \
        np.zeros((len(out["synteny"]), 0), np.int32)
    out["members"] = np.concatenate(allgather_var(res["members"]))
    moffs, base = [np.zeros(1, np.int64)], 0
    for p in allgather_var(res["member_off"]):
        moffs.append(p[1:] + base)
        base += int(p[-1])
    out["member_off"] = np.concatenate(moffs)
    order = [np.nonzero(out["call"] == c)[0] for c in (2, 1, 0)]
    out["call_counts"] = np.array([len(o) for o in order], dtype=np.int64)
    out["call_index"] = np.concatenate(order).astype(np.int64)
    return out
