"""World-size-2 gloo test of the data-parallel launcher (CPU): contiguous sharding, no
collective on the hot path, optional final all_gather reproduces the single-process result."""
import os
import socket

import torch
import torch.multiprocessing as mp

from sam_quantization_b200.launcher import ShardedEncoder, init_distributed, shard_bounds


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 5, 8, 17):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, results):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    r, w, _ = init_distributed("gloo")
    torch.manual_seed(0)
    enc = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3, padding=1), torch.nn.Tanh()).eval()  # stand-in per-image op
    images = torch.arange(5 * 3 * 8 * 8, dtype=torch.float32).view(5, 3, 8, 8) / 100.0
    sharded = ShardedEncoder(enc, r, w, micro_batch=2)
    local = sharded(images)
    lo, hi = shard_bounds(5, w, r)
    full = sharded(images, gather=True)
    with torch.no_grad():
        ref = enc(images)
    ok = local.shape[0] == hi - lo and torch.allclose(local, ref[lo:hi]) and torch.allclose(full, ref)
    results[rank] = bool(ok)
    torch.distributed.destroy_process_group()


def test_two_rank_gloo_shard_and_gather():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, port, results), nprocs=world, join=True)
    assert dict(results) == {0: True, 1: True}
