"""world_size-2 gloo test of the data-parallel plumbing (frame sharding + max-over-ranks timing)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nsc_b200.sharding import frame_shard, max_over_ranks, sum_over_ranks


def test_frame_shard_partitions():
    for n in (0, 1, 7, 128, 120000, 12000001):
        for world in (1, 2, 4, 8):
            spans = [frame_shard(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and a <= b and c <= d
            assert max(b - a for a, b in spans) <= -(-n // world) if n else True


def _worker(rank, world, port, n_frames, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    a, b = frame_shard(n_frames, rank, world)
    # every rank "codes" its shard: here a checksum of the frame ids it owns
    local = float(sum(range(a, b)))
    total = sum_over_ranks(local)
    slowest = max_over_ranks(1.0 + rank)
    got = [None] * world
    dist.all_gather_object(got, (a, b))
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, total, slowest, got))


def test_two_rank_gloo_sharding():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    n = 1001
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, total, slowest, got in res:
        assert total == float(sum(range(n)))
        assert slowest == 2.0
        assert got == [(0, 501), (501, 1001)]
