"""CPU test of the data-parallel host logic (world_size 2, gloo): contiguous batch shards + ONE in-place
all-gather of logits reproduce the single-process result, in rank order.  The oracle stands in for the engine
(no GPU here); the GPU path uses the same shard_bounds / all_gather_logits code with NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from oracle import ref_torch
from cases import cfg_of


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, global_batch, q):
    from vit_tensorflow_b200.runtime import shard_bounds, all_gather_logits
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfg = cfg_of("vit_small")
        w = oracle.init_weights(cfg, 0)
        img = oracle.make_image(cfg, global_batch, 1)            # every rank can build the global batch (seeded)
        lo, hi = shard_bounds(global_batch, world, rank)
        per = hi - lo
        gathered = torch.zeros((global_batch, cfg["num_classes"]), dtype=torch.float32)
        local = gathered[lo:hi]                                  # in-place: this rank's slice of the gather buffer
        local.copy_(torch.from_numpy(ref_torch.forward(img[lo:hi], w, cfg)))
        assert local.data_ptr() == gathered.data_ptr() + lo * cfg["num_classes"] * 4 and per * world == global_batch
        all_gather_logits(gathered, local, world)
        q.put((rank, gathered.numpy().copy()))
    finally:
        dist.destroy_process_group()


def test_dp_shard_allgather_matches_single_process():
    world, global_batch = 2, 6
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, global_batch, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cfg = cfg_of("vit_small")
    full = ref_torch.forward(oracle.make_image(cfg, global_batch, 1), oracle.init_weights(cfg, 0), cfg)
    for r in range(world):
        np.testing.assert_allclose(results[r], full, rtol=1e-5, atol=1e-6)   # same order on every rank


def test_shard_bounds():
    from vit_tensorflow_b200.runtime import shard_bounds
    assert [shard_bounds(1024, 8, r) for r in (0, 3, 7)] == [(0, 128), (384, 512), (896, 1024)]
    with pytest.raises(ValueError):
        shard_bounds(10, 4, 0)
