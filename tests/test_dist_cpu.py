"""CPU: the N>1 host logic (sharding + result gather) over gloo, world_size 2.

The per-shard scoring is stood in for by the oracle here (no GPU in this container); what is under
test is the shard arithmetic and the gather, which must reproduce the single-process results.
"""
import os
import socket
import sys

import numpy as np
import pytest

import helpers


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist
    from oracle import orgscorer_oracle as oracle
    from waafle_b200 import dist as wdist, synth
    from waafle_b200.params import OrgscorerParams
    dist.init_process_group("gloo", rank=rank, world_size=world)
    data = synth.generate_config("cfg2", n_contigs=60, seed=77, annotations=True)
    tax = data.taxonomy()
    batch = data.to_batch(tax)
    P = OrgscorerParams(n_systems=1)
    shard, c0, c1 = wdist.local_shard(batch, rank, world)
    res = oracle.score_batch(P.as_dict(), tax.tables(), shard.arrays())
    full = wdist.gather_results(res, int(batch.hit_off[c0]), dist)
    if rank == 0:
        ref = oracle.score_batch(P.as_dict(), tax.tables(), batch.arrays())
        q.put(helpers.compare_results(ref, full))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds_balance():
    from waafle_b200 import dist as wdist
    rng = np.random.default_rng(0)
    hit_off = np.concatenate([[0], np.cumsum(rng.integers(0, 500, size=1000))])
    for world in (1, 2, 3, 8):
        b = wdist.shard_bounds(hit_off, world)
        assert b[0] == 0 and b[-1] == 1000 and np.all(np.diff(b) >= 0) and len(b) == world + 1
        loads = np.diff(hit_off[b])
        assert loads.max() <= hit_off[-1] / world + 600


def test_two_rank_gather_matches_single_process():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    diffs = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert not diffs, diffs[:4]
