"""World-size-2 test of the N>1 path on CPU (gloo): each rank runs the forward pass for its
shard of the batch (the oracle stands in for the GPU kernels, which need a B200) and one
all-gather reassembles the logits exactly as a single process would have produced them."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_global, q, weights=None):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["OMP_NUM_THREADS"] = "2"
    import mnv1_b200  # noqa: F401
    from mnv1_b200 import shard, synth
    import oracle
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    a, b = shard.weighted_range(n_global, rank, world, weights)
    w = synth.weights()
    sc, sh = synth.batchnorm()
    local, _ = oracle.forward(synth.images(b - a, first=a), w, sc, sh)  # images indexed globally
    full = shard.gather_logits(torch.from_numpy(local), n_global, weights=weights)
    if rank == 0:
        q.put(full.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_global,weights", [(4, None), (3, None), (4, [22.7, 37.9])])
def test_two_rank_shard_and_gather(n_global, weights, oracle_mod, synth_net):
    """equal shards, a ragged batch, and shards in proportion to the ranks' host links (1 + 3 images)"""
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_global, q, weights)) for r in range(2)]
    for p in procs:
        p.start()
    full = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from mnv1_b200 import synth
    w, sc, sh = synth_net
    want, _ = oracle_mod.forward(synth.images(n_global), w, sc, sh)
    assert full.shape == (n_global, 1000)
    assert np.array_equal(full, want)
